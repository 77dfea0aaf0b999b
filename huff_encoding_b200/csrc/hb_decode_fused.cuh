// hb_decode_fused.cuh -- K3f: decompress (comp.rs:487-519) in ONE pass over the stream: every code word is decoded once.
//
// The two-pass decoder (hb_decode.cuh) decodes everything twice -- a count pass to learn where each thread's letters go,
// then a write pass -- and reads the stream twice.  Here a thread decodes its subsequence ONCE, with a multi-letter
// table (up to three letters per lookup), into a private slot in shared memory; where the letters belong in the output
// is settled afterwards:
//
//   team      256 threads that own one CHUNK of the stream (256 subsequences of kFSubWords 32-bit words).  A CTA holds
//             1..4 teams that share the lookup tables and synchronise on their own named barriers; chunks are handed out
//             by an atomic ticket, so a chunk only ever waits for chunks that started before it.
//   stage     the chunk (+ 16-word halos) goes to shared memory DENSE: a subsequence is an odd number of words long, so
//             the 32 lanes of a warp walking their own subsequences in lock step hit 32 different banks without padding.
//             The window of the NEXT chunk is fetched by one 1-D bulk asynchronous copy (cp.async.bulk + mbarrier
//             complete_tx, SASS UBLKCP) issued as soon as the current window is dead, so it lands while the team waits
//             for its look-back and copies its rows out; a byte-swap pass in shared memory makes it MSB-first.
//   phase A   entry candidate by self-synchronisation from a look-back window (as in hb_decode.cuh).
//   phase B   decode entry .. end of the subsequence with the EMIT table: entry = 3 letters | count | bits consumed.
//             Letters are appended to a 32-bit accumulator and every completed word is stored to the thread's slot (a
//             plain store; slots are an odd number of words apart).  The last < kEmitBits bits before the end of the
//             subsequence are decoded letter by letter (first letter of the entry + a 256-byte length table).
//   verify    entry[t] == exit[t-1] inside the team, iterated to a fixed point (a refuted thread decodes again).
//   offsets   team scan of the letter counts; the team's first output position comes from a decoupled look-back over
//             per-chunk descriptors (status | exit | count), which also checks entry == predecessor's exit ACROSS chunks.
//             A mismatch only raises a flag: the host then falls back to the two-pass decoder (exact, with serial repair).
//   compact   every thread owns the 32-byte output rows that start inside its letters; it first pulls the < 32 letters
//             that complete its last row from its successors' slots (or, at the end of the chunk, decodes them from the
//             halo), then copies its slot to global memory with one 256-bit store per row.
//
// Eligible trees: every code fits the emit table (max_len <= kEmitBits), >= 2 leaves, no duplicate letters; the entry of the
// first code word must be known.  Everything else takes the two-pass kernels.
//
// Algorithmic HBM bytes: C + N, and that is what this kernel moves (+ 8 bytes of descriptor per 33 KiB of stream).
#pragma once

#include "hb_common.cuh"
#include "hb_decode.cuh"

namespace hb {

#ifndef HB_FUSED_SUB_WORDS
#define HB_FUSED_SUB_WORDS 33
#endif
#ifndef HB_EMIT_BITS
#define HB_EMIT_BITS 12
#endif
constexpr int kFSubWords = HB_FUSED_SUB_WORDS;                 // words per subsequence: ODD (bank-conflict-free dense layout)
static_assert(kFSubWords % 2 == 1, "subsequence length must be an odd number of words");
constexpr int kFSubBits = kFSubWords * 32;
#ifndef HB_FUSED_TEAM
#define HB_FUSED_TEAM 256
#endif
constexpr int kFTeam = HB_FUSED_TEAM;                          // threads per team
constexpr int kFMaxTeams = 1024 / kFTeam;
static_assert(kFTeam % 32 == 0 && kFTeam >= 64 && kFMaxTeams <= 15, "teams synchronise on named barriers 1..15");
constexpr int kFChunkWords = kFTeam * kFSubWords;
static_assert(kFChunkWords % 4 == 0, "chunks must keep 16-byte alignment");
constexpr int kFHalo = 16;                                     // words staged before and after the chunk
constexpr int kFWinWords = kFHalo + kFChunkWords + kFHalo;
constexpr int kFWinAlloc = kFWinWords + 4;                     // + look-ahead slack
constexpr uint32_t kFWinBits = kFWinWords * 32u;
constexpr int kEmitBits = HB_EMIT_BITS;
#ifndef HB_FUSED_LOOKBACK_BITS
#define HB_FUSED_LOOKBACK_BITS 320       // in-team look-back: a refuted thread costs its whole team (and, through the scan,
#endif                                   // every later chunk) a second decode, so the window is longer than the two-pass one
constexpr int kFLookbackBits = HB_FUSED_LOOKBACK_BITS;
static_assert(kEmitBits <= 13 && kEmitBits >= 8, "emit table index width");
static_assert(kFHalo * 32 >= HB_LEAD_LOOKBACK_BITS, "the leading look-back must fit the halo");
static_assert(kFHalo * 32 >= 32 * kEmitBits + kEmitBits + 64, "31 extra letters + one code word + look-ahead must fit the halo");

constexpr uint64_t kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62;
constexpr int kDescExitShift = 42;
constexpr uint64_t kDescValueMask = (1ull << kDescExitShift) - 1;
constexpr uint32_t kDescExitEnd = 0xFFFFFu;

struct FusedResult {
    unsigned long long total_letters;
    unsigned long long entry0;         // absolute buffer bit, kEnd64 = none
    unsigned long long exit_last;
    uint32_t error;                    // bit 0: a chunk's entry != its predecessor's exit (speculation refuted)
    uint32_t slow_chunks;              // chunks that overflowed their slots and were written letter by letter
    unsigned long long phase_cycles[8]; // SM clock cycles summed over all chunks (first thread of each team): stage, decode
                                       // (phase A + B), verify rounds, scan, look-back, compaction, [6] = chunks
};

struct FusedParams {
    const uint32_t *words;
    uint64_t n_words_readable;
    uint64_t avail_bits;
    uint64_t own_begin, own_end;
    uint64_t entry_bit;                // known first code-word start (>= own_begin)
    uint64_t stream_bit0;
    uint32_t len_gcd, fixed_len;
    uint32_t first_chunk, n_chunks;
    uint32_t slot_words;               // per-thread slot, in words (odd)
    uint32_t spoil_speculation;
    uint32_t max_decoders;             // teams of a CTA that may be in their decode phase at the same time (0 = all)
    const uint32_t *emit;              // 1 << kEmitBits entries
    const uint8_t *code_len;           // 256 bytes
    unsigned long long *desc;          // n_chunks, zeroed before the launch
    uint32_t *ticket;                  // zeroed before the launch
    FusedResult *result;               // zeroed before the launch
    uint8_t *out;
    uint64_t out_cap;
};

// per-team shared memory (bytes) for a slot of `slot_words` words
__host__ __device__ constexpr size_t fused_team_bytes(uint32_t slot_words) {
    return static_cast<size_t>(kFWinAlloc) * 4 + static_cast<size_t>(kFTeam) * slot_words * 4 + kFTeam * 4 + 128;
}
__host__ __device__ constexpr size_t fused_shared_bytes() { return (static_cast<size_t>(1) << kEmitBits) * 4 + 256 + 16; }

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(kFTeam) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
// Descriptor traffic of the look-back: a descriptor carries its whole payload in one 64-bit word, so relaxed GPU-scope
// accesses are enough (an acquire load would invalidate L1 on every poll, a release store would first drain the
// thread's output stores).
__device__ __forceinline__ unsigned long long ld_relaxed_ull(const unsigned long long *p) {
    unsigned long long r;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_relaxed_ull(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Bit reader over the DENSE staged window: three words + position.  The funnel shift takes the position modulo 32 and a
// refill is due exactly when bit 5 of the position flips (a step is < 32 bits).  The third word is loaded one refill
// ahead, so the shared-memory latency of a refill is off the position -> peek -> lookup -> position chain.  Refills are
// predicated, never branches.
struct FReader {
    uint32_t w0, w1, w2, q, wa;        // wa: shared address of the next word to load
    __device__ __forceinline__ void init(uint32_t win, uint32_t q0) {
        q = q0;
        wa = win + ((q0 >> 5) << 2);
        w0 = lds32(wa);
        w1 = lds32(wa + 4);
        w2 = lds32(wa + 8);
        wa += 12;
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(w1, w0, q); }
    __device__ __forceinline__ void step(uint32_t bits) {
        const uint32_t qn = q + bits;
        const uint32_t r = (qn ^ q) & 32u;
        asm volatile("{\n\t.reg .pred f;\n\tsetp.ne.u32 f, %5, 0;\n\t@f mov.u32 %0, %1;\n\t@f mov.u32 %1, %2;\n\t"
                     "@f ld.shared.u32 %2, [%3];\n\t@f add.u32 %3, %3, 4;\n\t}"
                     : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(wa) : "r"(0u), "r"(r) : "memory");
        q = qn;
    }
};

// byte offset of the emit-table entry for the next kEmitBits bits
__device__ __forceinline__ uint32_t emit_off(uint32_t x) {
    uint32_t y;
    asm("and.b32 %0, %1, %2;" : "=r"(y) : "r"(x), "n"(~((1u << (32 - kEmitBits)) - 1u)));
    return y >> (30 - kEmitBits);
}

// Advance from q over whole code words, no output: first code-word start >= q_stop, kEnd32 if a code word does not end
// at or before q_avail.  `letters` counts them.
__device__ __forceinline__ uint32_t fused_run(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_stop,
                                              uint32_t q_avail, uint32_t &letters) {
    letters = 0;
    if (q >= q_stop) return q;
    FReader rd;
    rd.init(win, q);
    uint32_t acc = 0;                                      // sum of the entries' top bytes: bits << 4 | count
    const uint32_t q_begin = q;
    const uint32_t lim = min(q_stop, q_avail);
    if (lim >= static_cast<uint32_t>(kEmitBits)) {
        const uint32_t last = lim - kEmitBits;
        while (rd.q <= last) {
            const uint32_t e = lds32(lut + emit_off(rd.peek()));
            rd.step(e >> 28);
            acc += e >> 24;
        }
    }
    uint32_t n1 = 0, ret = 0;
    for (;;) {
        if (rd.q >= q_stop) { ret = rd.q; break; }
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        const uint32_t len = lds8(lens + (e & 0xFFu));
        if (rd.q + len > q_avail) { ret = kEnd32; break; }
        rd.step(len);
        acc += len << 4;
        n1++;
    }
    letters = acc - ((rd.q - q_begin) << 4) + n1;
    return ret;
}

// Letter sink of one thread: accumulator word + letter position + slot address.  A completed word goes to the slot
// with a predicated store (no branch).  `p` counts letters modulo 16 only (the whole top byte of an entry is added);
// the letter count is recovered from the slot address.
struct FSink {
    uint32_t acc, p, wp;
    __device__ __forceinline__ void init(uint32_t slot) { acc = 0; p = 0; wp = slot; }
    // e = emit-table entry: <= 3 letters in the low bytes, count in bits 24..25, anything above
    __device__ __forceinline__ void append(uint32_t e) {
        const uint32_t L = e & 0xFFFFFFu;
        const uint32_t s = p << 3;                         // funnel shifts take the amount modulo 32
        acc |= __funnelshift_l(0u, L, s);                  // L << (8 * (p & 3))
        const uint32_t hi = __funnelshift_l(L, 0u, s);     // the letters that fall past the accumulator word
        const uint32_t pn = p + (e >> 24);
        const uint32_t fl = (pn ^ p) & 4u;                 // <= 3 letters per append: at most one word completes
        asm volatile("{\n\t.reg .pred f;\n\tsetp.ne.u32 f, %3, 0;\n\t@f st.shared.u32 [%2], %0;\n\t@f mov.u32 %0, %1;\n\t}"
                     : "+r"(acc) : "r"(hi), "r"(wp), "r"(fl) : "memory");
        wp += fl;
        p = pn;
    }
    __device__ __forceinline__ uint32_t stored(uint32_t slot) const { return (wp - slot) + (p & 3u); }
    __device__ __forceinline__ void finish() { if (p & 3u) sts32(wp, acc); }
};

// Decode [entry, q_hi) into the thread's slot.  Returns the exit (first code-word start >= q_hi, kEnd32 at the end of
// the stream); count = letters whose code word starts in [entry, q_hi) and ends <= q_avail.  When the slot fills up the
// remaining letters are only counted (count > capacity tells the caller).
__device__ __forceinline__ uint32_t fused_emit(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_hi,
                                               uint32_t q_avail, uint32_t slot, uint32_t slot_end, uint32_t &count) {
    count = 0;
    if (q == kEnd32) return kEnd32;
    if (q >= q_hi) return q;
    FReader rd;
    rd.init(win, q);
    FSink sk;
    sk.init(slot);
    const uint32_t wsafe = slot_end - 12;                  // a trip of two lookups completes at most two words
    const uint32_t lim = min(q_hi, q_avail);
    if (lim >= 2u * kEmitBits) {
        const uint32_t last2 = lim - 2 * kEmitBits;        // both lookups of a trip start at or before lim - kEmitBits
#pragma unroll 1
        while (rd.q <= last2 && sk.wp <= wsafe) {
            const uint32_t e1 = lds32(lut + emit_off(rd.peek()));
            rd.step(e1 >> 28);
            sk.append(e1);
            const uint32_t e2 = lds32(lut + emit_off(rd.peek()));
            rd.step(e2 >> 28);
            sk.append(e2);
        }
    }
    uint32_t exitq = 0;
#pragma unroll 1
    for (;;) {                                             // the last few letters before q_hi, one at a time
        if (rd.q >= q_hi || sk.wp > wsafe) { exitq = rd.q; break; }
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        const uint32_t letter = e & 0xFFu;
        const uint32_t len = lds8(lens + letter);
        if (rd.q + len > q_avail) { exitq = kEnd32; break; }
        rd.step(len);
        sk.append(letter | (1u << 24));
    }
    sk.finish();
    count = sk.stored(slot);
    if (exitq != kEnd32 && exitq < q_hi) {                 // slot full: count the rest without storing
        uint32_t more;
        exitq = fused_run(win, lut, lens, exitq, q_hi, q_avail, more);
        count += more;
    }
    return exitq;
}

// Generic pull (rare: the successor holds fewer letters than a row, or the thread is the team's last): up to `need`
// letters that follow the thread's own, from the successors' slots, then -- past the team's last letter -- decoded from the
// halo behind the chunk.  Appended bytewise to the thread's slot at offset `cnt`.  Returns how many were appended.
__device__ __noinline__ uint32_t fused_pull(uint32_t slots, uint32_t slot_bytes, const uint32_t *s_cnt, uint32_t tt,
                                            uint32_t cnt, uint32_t need, uint32_t win, uint32_t lut, uint32_t lens,
                                            uint32_t q_chunk_exit, uint32_t q_own_end, uint32_t q_avail) {
    uint32_t got = 0;
    const uint32_t dst = slots + tt * slot_bytes + cnt;
    uint32_t u = tt + 1;
    while (got < need && u < static_cast<uint32_t>(kFTeam)) {
        const uint32_t c = s_cnt[u];
        const uint32_t src = slots + u * slot_bytes;
        for (uint32_t o = 0; o < c && got < need; o++, got++) sts8(dst + got, lds8(src + o));
        u++;
    }
    if (got < need && q_chunk_exit != kEnd32) {
        FReader rd;
        rd.init(win, q_chunk_exit);
        while (got < need && rd.q < q_own_end) {
            const uint32_t e = lds32(lut + emit_off(rd.peek()));
            const uint32_t letter = e & 0xFFu;
            const uint32_t len = lds8(lens + letter);
            if (rd.q + len > q_avail) break;
            rd.step(len);
            sts8(dst + got, letter);
            got++;
        }
    }
    return got;
}

// Slow path of a chunk in which some slot overflowed (a run of short codes: more letters than the slots were sized for).
// Positions and counts are known by now, so every thread decodes its letters a SECOND time, letter by letter, straight
// into registers -- 32 letters, one 256-bit store per output row it owns, like the two-pass write kernel -- and runs on
// past its subsequence to complete its last row (the window and the halo hold those bits).
__device__ __noinline__ void fused_slow_rows(uint32_t win, uint32_t lut, uint32_t lens, uint32_t entry, uint32_t q_own_end,
                                             uint32_t q_avail, uint32_t count, uint8_t *out, uintptr_t out_addr, uint64_t D,
                                             uint64_t out_cap) {
    if (entry == kEnd32 || count == 0) return;
    FReader rd;
    rd.init(win, entry);
    auto next = [&](uint32_t &letter) -> bool {                // false: no further letter of the owned range exists
        if (rd.q >= q_own_end) return false;
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        letter = e & 0xFFu;
        const uint32_t len = lds8(lens + letter);
        if (rd.q + len > q_avail) return false;
        rd.step(len);
        return true;
    };
    const uint64_t end = D + count;
    uint64_t pos = D;
    uint32_t letter = 0;
    if (D == 0) {                                               // ragged head of the whole output
        for (; pos < end && ((out_addr + pos) & 31); pos++) {
            if (!next(letter)) return;
            if (pos < out_cap) out[pos] = static_cast<uint8_t>(letter);
        }
    } else {
        for (; pos < end && ((out_addr + pos) & 31); pos++)    // these sit in a row my predecessor writes
            if (!next(letter)) return;
    }
    while (pos < end) {                                         // rows that start inside my letters
        uint32_t v[8];
        uint32_t got = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t w = 0;
#pragma unroll
            for (int b8 = 0; b8 < 4; b8++) {
                if (got == static_cast<uint32_t>(4 * k + b8) && next(letter)) {
                    w |= letter << (8 * b8);
                    got++;
                }
            }
            v[k] = w;
        }
        if (got == 32 && pos + 32 <= out_cap) {
            stg256(out + pos, v);
        } else {
            for (uint32_t i = 0; i < got; i++)
                if (pos + i < out_cap) out[pos + i] = static_cast<uint8_t>(v[i >> 2] >> (8 * (i & 3)));
        }
        if (got < 32) return;
        pos += 32;
    }
}

// ---- 1-D bulk asynchronous copy global -> shared with mbarrier completion (TMA engine; SASS: UBLKCP + SYNCS)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\tDONE_%=:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

extern __shared__ __align__(16) uint8_t fused_smem[];

// -DHB_FUSED_TIMING: per-phase SM clock stamps (hb_ctx_fused_phase_cycles); off in the product build (register pressure)
#ifdef HB_FUSED_TIMING
#define HB_TK(i) tk[i] = clock()
#else
#define HB_TK(i) ((void)0)
#endif

// s_misc words
enum { kMWarp = 0, kMFlag = 8, kMChunk = 9, kMBaseLo = 10, kMBaseHi = 11, kMSlow = 12, kMExit = 13, kMEntry = 14,
       kMNext = 15, kMBar = 16 /* 8 bytes */, kMPrefetched = 18 };

__global__ void __launch_bounds__(kFTeam * kFMaxTeams, 1)
dec_fused_kernel(const FusedParams p) {
    const int team = threadIdx.x / kFTeam, tt = threadIdx.x % kFTeam;
    const int lane = tt & 31;
    // shared by the CTA: emit table + code lengths; then one block per team
    uint32_t *s_lut = reinterpret_cast<uint32_t *>(fused_smem);
    uint8_t *s_lens = fused_smem + (static_cast<size_t>(1) << kEmitBits) * 4;
    uint32_t *s_token = reinterpret_cast<uint32_t *>(s_lens + 256);   // teams currently decoding
    const size_t team_bytes = fused_team_bytes(p.slot_words);
    uint8_t *tb = fused_smem + fused_shared_bytes() + team * team_bytes;
    uint32_t *s_win = reinterpret_cast<uint32_t *>(tb);
    uint32_t *s_slots = s_win + kFWinAlloc;
    uint32_t *s_exit = s_slots + static_cast<size_t>(kFTeam) * p.slot_words;
    uint32_t *s_cnt = s_exit;                               // the exits are dead when the counts are written
    uint32_t *s_misc = s_exit + kFTeam;

    for (int i = threadIdx.x; i < (1 << kEmitBits); i += blockDim.x) s_lut[i] = p.emit[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_lens)[i] = reinterpret_cast<const uint32_t *>(p.code_len)[i];

    uint32_t b = smem_addr(fused_smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(b));             // one opaque base register (see hb_decode.cuh)
    const uint32_t a_lut = b;
    const uint32_t a_lens = b + (1u << kEmitBits) * 4u;
    const uint32_t a_win = b + static_cast<uint32_t>(fused_shared_bytes() + team * team_bytes);
    const uint32_t a_slots = a_win + kFWinAlloc * 4u;
    const uint32_t slot_bytes = p.slot_words * 4u;
    const uint32_t a_slot = a_slots + tt * slot_bytes;
    const uint32_t a_bar = smem_addr(s_misc + kMBar);
    const uintptr_t out_addr = reinterpret_cast<uintptr_t>(p.out);
    if (threadIdx.x == 0) *s_token = 0;
    if (tt == 0) {
        mbar_init(a_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_misc[kMNext] = atomicAdd(p.ticket, 1u);
        s_misc[kMPrefetched] = 0;
    }
    __syncthreads();
    uint32_t bar_parity = 0;
    // chunks whose whole window lies inside the 16-byte-aligned readable part of the stream can be fetched in bulk
    const uint64_t bulk_words = p.n_words_readable & ~3ull;
    auto bulk_ok = [&](uint32_t jj) -> bool {
        const long long wb = static_cast<long long>(p.first_chunk + jj) * kFChunkWords - kFHalo;
        return wb >= 0 && static_cast<uint64_t>(wb) + kFWinWords <= bulk_words;
    };

#ifdef HB_FUSED_TIMING
    uint32_t tk[7], tk_lb = 0;
#endif
    for (;;) {
        const uint32_t j = s_misc[kMNext];                 // chunk, relative to first_chunk (ticket order)
        const bool prefetched = s_misc[kMPrefetched] != 0;
        if (j >= p.n_chunks) break;
        HB_TK(0);
        const uint32_t chunk = p.first_chunk + j;

        // ---- stage [chunk * kFChunkWords - kFHalo, + kFWinWords) MSB-first
        if (prefetched) {
            // the bulk copy was issued during the previous chunk: wait for its bytes, then byte-swap in place
            mbar_wait(a_bar, bar_parity);
            bar_parity ^= 1u;
            constexpr int kVecs = kFWinWords / 4;
#pragma unroll
            for (int k = 0; k < (kVecs + kFTeam - 1) / kFTeam; k++) {
                const int i4 = tt + k * kFTeam;
                if (i4 < kVecs) {
                    uint4 v = reinterpret_cast<uint4 *>(s_win)[i4];
                    reinterpret_cast<uint4 *>(s_win)[i4] = make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
                }
            }
        } else {
            // all loads of a thread in flight together
            const long long w_begin = static_cast<long long>(chunk) * kFChunkWords - kFHalo;
            constexpr int kVecs = kFWinWords / 4;
            constexpr int kPer = (kVecs + kFTeam - 1) / kFTeam;
            uint4 v[kPer];
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int i4 = tt + k * kFTeam;
                const long long gw = w_begin + 4ll * i4;
                v[k] = make_uint4(0, 0, 0, 0);
                if (i4 < kVecs && gw >= 0 && static_cast<uint64_t>(gw) < p.n_words_readable) {
                    if (static_cast<uint64_t>(gw) + 4 <= p.n_words_readable) {
                        v[k] = ld_stream_u4(reinterpret_cast<const uint4 *>(p.words + gw));
                    } else {
                        const uint64_t left = p.n_words_readable - static_cast<uint64_t>(gw);
                        v[k].x = ld_stream_u32(p.words + gw);
                        if (left > 1) v[k].y = ld_stream_u32(p.words + gw + 1);
                        if (left > 2) v[k].z = ld_stream_u32(p.words + gw + 2);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int i4 = tt + k * kFTeam;
                if (i4 < kVecs)
                    reinterpret_cast<uint4 *>(s_win)[i4] = make_uint4(bswap32(v[k].x), bswap32(v[k].y), bswap32(v[k].z), bswap32(v[k].w));
            }
        }
        if (tt < 4) s_win[kFWinWords + tt] = 0;
        // ---- decode token: the teams of a CTA take turns in the decode phase, so that one team's look-back wait, row
        //      copies and window fetch run under another team's decode instead of all teams idling in the same phase
        if (p.max_decoders && tt == 0) {
            for (;;) {
                const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(s_token);
                if (cur < p.max_decoders && atomicCAS(s_token, cur, cur + 1) == cur) break;
                __nanosleep(64);
            }
        }
        team_sync(team);
        HB_TK(1);

        // ---- window coordinates: window bit q <-> buffer bit win_bit0 + q
        const long long win_bit0 = (static_cast<long long>(chunk) * kFChunkWords - kFHalo) * 32;
        auto to_win = [&](uint64_t abs_bit) -> uint32_t {
            const long long q = static_cast<long long>(abs_bit) - win_bit0;
            return q < 0 ? 0u : (q > static_cast<long long>(kFWinBits) ? kFWinBits : static_cast<uint32_t>(q));
        };
        const uint32_t q_avail = to_win(p.avail_bits);
        const uint32_t q_own_begin = to_win(p.own_begin);
        const uint32_t q_own_end = to_win(p.own_end);
        const uint32_t q_buf0 = to_win(0);
        const uint32_t q_sub = (kFHalo + tt * kFSubWords) * 32u;
        const uint32_t q_lo = max(q_sub, q_own_begin);
        const uint32_t q_hi = min(q_sub + kFSubBits, q_own_end);
        const bool active = q_lo < q_sub + kFSubBits && q_sub < q_own_end;
        const bool is_first = active && q_own_begin >= q_sub;
        const bool has_pred = active && !is_first && tt > 0;

        // ---- phase A: entry candidate
        uint32_t entry = kEnd32;
        if (active) {
            if (is_first) {
                entry = to_win(p.entry_bit);
            } else {
                uint32_t window = tt == 0 ? static_cast<uint32_t>(kLeadLookbackBits) : static_cast<uint32_t>(kFLookbackBits);
                if (p.fixed_len) window = 0;
                if (p.spoil_speculation && tt == 0) window = 0;
                uint32_t q0 = q_lo > window ? q_lo - window : 0;
                if (q0 < q_buf0) q0 = q_buf0;
                if (p.len_gcd > 1) {
                    const unsigned long long abs0 = static_cast<unsigned long long>(win_bit0 + q0) + p.stream_bit0;
                    const uint32_t rem = static_cast<uint32_t>(abs0 % p.len_gcd);
                    if (rem) q0 += p.len_gcd - rem;
                }
                uint32_t dummy;
                entry = fused_run(a_win, a_lut, a_lens, q0, q_lo, q_avail, dummy);
            }
        }

        // ---- phase B + in-team verification to a fixed point
        uint32_t count = 0, exitq = kEnd32;
        bool redo = active;
        for (int round = 0;; round++) {
            if (round > kFTeam + 1) asm volatile("trap;");
            if (redo) exitq = fused_emit(a_win, a_lut, a_lens, entry, q_hi, q_avail, a_slot, a_slot + slot_bytes, count);
            s_exit[tt] = exitq;
            if (tt == 0) s_misc[kMFlag] = 0;
            team_sync(team);
            if (round == 0) { HB_TK(2); }
            redo = false;
            if (has_pred) {
                const uint32_t want = s_exit[tt - 1];
                if (want != entry) { entry = want; redo = true; s_misc[kMFlag] = 1; }
            }
            team_sync(team);
            const uint32_t any = s_misc[kMFlag];
            team_sync(team);
            if (!any) break;
        }
        if (!active) count = 0;
        if (p.max_decoders && tt == 0) atomicSub(s_token, 1u);
        HB_TK(3);

        // ---- team scan of the letter counts; overflow flag; the ticket of the next chunk
        const bool overflow = count + 40u > slot_bytes;        // own letters + a pulled row (36 bytes from a word boundary) must fit
        const uint32_t incl = warp_incl_scan(count);
        if (lane == 31) s_misc[kMWarp + (tt >> 5)] = incl;
        s_cnt[tt] = count;
        const bool is_last_active = active && (tt == kFTeam - 1 || q_sub + kFSubBits >= q_own_end);
        if (is_last_active) s_misc[kMExit] = exitq;
        if (active && !has_pred) s_misc[kMEntry] = entry;
        if (tt == 0) {
            s_misc[kMSlow] = 0;
            s_misc[kMNext] = atomicAdd(p.ticket, 1u);
        }
        team_sync(team);
        if (overflow) s_misc[kMSlow] = 1;
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kFTeam / 32; k++) { const uint32_t w = s_misc[kMWarp + k]; if (k < (tt >> 5)) before += w; total += w; }
        const uint32_t off = before + incl - count;
        const uint32_t q_chunk_exit = s_misc[kMExit];
        const uint32_t next_j = s_misc[kMNext];
        HB_TK(4);

        // ---- decoupled look-back (first warp of the team): exclusive letter offset of this chunk
        if (tt < 32) {
            const uint32_t q_chunk_end = (kFHalo + kFChunkWords) * 32u;
            const uint32_t exit_rel = q_chunk_exit == kEnd32 ? kDescExitEnd : (q_chunk_exit - q_chunk_end) & 0xFFFFFu;
            const uint32_t q_entry = s_misc[kMEntry];
            const unsigned long long mine = (static_cast<unsigned long long>(exit_rel) << kDescExitShift) | total;
            if (lane == 0 && j + 1 < p.n_chunks) st_relaxed_ull(p.desc + j, kDescAgg | mine);
            unsigned long long excl = 0;
            if (j > 0) {
                // 128 descriptors per trip (4 per lane, all loads in flight together): when the teams run in waves the
                // nearest inclusive prefix is up to one wave (~300 chunks) back
                long long look = static_cast<long long>(j) - 1;
                bool first_window = true;
                for (;;) {
                    unsigned long long d[4];
                    bool empty;
                    do {
                        empty = false;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const long long idx = look - lane - 32 * k;
                            d[k] = idx >= 0 ? ld_relaxed_ull(p.desc + idx) : kDescPrefix;
                            empty = empty || (d[k] >> 62) == 0;
                        }
                    } while (__any_sync(0xFFFFFFFFu, empty));
                    if (first_window && lane == 0) {
                        // my entry must be where the chunk before me stopped
                        const uint32_t pred_exit = static_cast<uint32_t>(d[0] >> kDescExitShift) & 0xFFFFFu;
                        const uint32_t entry_rel = q_entry == kEnd32 ? kDescExitEnd : (q_entry - kFHalo * 32u) & 0xFFFFFu;
                        if (pred_exit != entry_rel) atomicOr(&p.result->error, 1u);
                    }
                    first_window = false;
                    bool done = false;
                    unsigned long long v = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t pref_mask = __ballot_sync(0xFFFFFFFFu, (d[k] >> 62) == 2);
                        const int stop = pref_mask ? __ffs(pref_mask) - 1 : 31;
                        if (!done && lane <= stop) v += d[k] & kDescValueMask;
                        done = done || pref_mask != 0;
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, sft);
                    excl += v;
                    if (done) break;
                    look -= 128;
                }
            }
            if (lane == 0) {
                if (j + 1 < p.n_chunks) {
                    st_relaxed_ull(p.desc + j, kDescPrefix | (static_cast<unsigned long long>(exit_rel) << kDescExitShift) |
                                               ((excl + total) & kDescValueMask));
                } else {
                    p.result->total_letters = excl + total;
                    p.result->exit_last = q_chunk_exit == kEnd32 ? kEnd64 : static_cast<unsigned long long>(win_bit0 + q_chunk_exit);
                }
                if (j == 0) p.result->entry0 = q_entry == kEnd32 ? kEnd64 : static_cast<unsigned long long>(win_bit0 + q_entry);
                s_misc[kMBaseLo] = static_cast<uint32_t>(excl);
                s_misc[kMBaseHi] = static_cast<uint32_t>(excl >> 32);
            }
        }

#ifdef HB_FUSED_TIMING
        tk_lb = clock();
#endif
        // ---- pull: the letters that may complete my last output row (up to 31, whatever the row phase turns out to be)
        //      are appended to my slot NOW, while the first warp waits for the look-back: 9 words of the successor's slot,
        //      shifted to my letter count.  Needs nothing but the team's counts.
        uint32_t avail_after = 0;                               // letters after my own that my slot now holds
        if (count && !overflow) {
            const uint32_t nxt = tt + 1 < static_cast<uint32_t>(kFTeam) ? s_cnt[tt + 1] : 0u;
            if (nxt >= 36u) {
                const uint32_t src = a_slots + (tt + 1) * slot_bytes;
                const uint32_t dstw = a_slot + (count & ~3u);
                const uint32_t sh = (count & 3u) << 3;
                uint32_t prev = sh ? (lds32(dstw) << (32u - sh)) : 0u;    // my partial word, moved to the top bytes
                uint32_t sv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) sv[k] = lds32(src + 4 * k);
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    sts32(dstw + 4 * k, __funnelshift_l(prev, sv[k], sh));
                    prev = sv[k];
                }
                avail_after = 32;
            } else {
                avail_after = fused_pull(a_slots, slot_bytes, s_cnt, tt, count, 31u, a_win, a_lut, a_lens, q_chunk_exit,
                                         q_own_end, q_avail);
            }
        }
        team_sync(team);
        HB_TK(5);
        const bool slow_chunk = s_misc[kMSlow] != 0;
        const uint64_t base = (static_cast<uint64_t>(s_misc[kMBaseHi]) << 32) | s_misc[kMBaseLo];
        const uint64_t D = base + off;                          // output position of my first letter

        // ---- the window is dead (unless the chunk takes the slow path): fetch the next chunk's window in bulk
        bool will_prefetch = !slow_chunk && next_j < p.n_chunks && bulk_ok(next_j);
        if (tt == 0) {
            s_misc[kMPrefetched] = will_prefetch ? 1u : 0u;
            if (slow_chunk) atomicAdd(&p.result->slow_chunks, 1u);
            if (will_prefetch) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses before async writes
                const long long wb = static_cast<long long>(p.first_chunk + next_j) * kFChunkWords - kFHalo;
                mbar_expect_tx(a_bar, kFWinWords * 4u);
                bulk_g2s(a_win, p.words + wb, kFWinWords * 4u, a_bar);
            }
        }

        // ---- compaction
        if (slow_chunk) {
            // a slot overflowed somewhere in the team: every thread decodes its letters again, straight to its rows
            if (active) fused_slow_rows(a_win, a_lut, a_lens, entry, q_own_end, q_avail, count, p.out, out_addr, D, p.out_cap);
        } else if (count) {
            const uint64_t end = D + count;
            const uint32_t head = D == 0 ? 0u : static_cast<uint32_t>((0 - (out_addr + D)) & 31);   // letters my predecessor's row holds
            const uint32_t need = static_cast<uint32_t>((0 - (out_addr + end)) & 31);
            const uint32_t cnt_eff = head < count + need ? count + min(need, avail_after) : count;
            uint32_t o = head;
            if (D == 0) {                                       // nobody precedes the first letter: its ragged head is mine
                const uint32_t rag = static_cast<uint32_t>((0 - out_addr) & 31);
                for (; o < rag && o < cnt_eff; o++)
                    if (o < p.out_cap) p.out[o] = static_cast<uint8_t>(lds8(a_slot + o));
            }
#pragma unroll 1
            for (; o + 32 <= cnt_eff && D + o + 32 <= p.out_cap; o += 32) {
                const uint32_t wa = a_slot + (o & ~3u);
                const uint32_t sh = (o & 3u) << 3;
                uint32_t w[9];
#pragma unroll
                for (int k = 0; k < 9; k++) w[k] = lds32(wa + 4 * k);
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = __funnelshift_r(w[k], w[k + 1], sh);
                stg256(p.out + D + o, v);
            }
            for (; o < cnt_eff; o++)                            // end of the output (or of the caller's buffer)
                if (D + o < p.out_cap) p.out[D + o] = static_cast<uint8_t>(lds8(a_slot + o));
        }
        team_sync(team);                                        // slots (and, without prefetch, the window) are reused
#ifdef HB_FUSED_TIMING
        if (tt == 0) {
            HB_TK(6);
#pragma unroll
            for (int k = 0; k < 6; k++) atomicAdd(&p.result->phase_cycles[k], static_cast<unsigned long long>(tk[k + 1] - tk[k]));
            atomicAdd(&p.result->phase_cycles[6], 1ull);
            atomicAdd(&p.result->phase_cycles[7], static_cast<unsigned long long>(tk_lb - tk[4]));   // look-back alone
        }
#endif
    }
}

}  // namespace hb
