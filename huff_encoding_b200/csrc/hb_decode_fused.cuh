// hb_decode_fused.cuh -- K3f: decompress (comp.rs:487-519) in ONE kernel that reads the stream ONCE.
//
// The two-pass decoder (hb_decode.cuh) is two kernels -- a count pass to learn where each thread's letters go, then a write
// pass -- with the stream read from HBM by both and five tiny kernels + a host sync in between.  Here one persistent
// kernel stages a chunk of the stream in shared memory once and does both passes over it:
//
//   team      256 threads that own one CHUNK of the stream (256 subsequences of kFSubWords 32-bit words).  A CTA holds
//             up to 4 teams (32 warps per SM) that share the lookup table and synchronise on their own named barriers;
//             chunks are handed out by an atomic ticket, so a chunk only ever waits for chunks that started before it.
//   stage     the chunk (+ halos) goes to shared memory DENSE: a subsequence is an odd number of words long, so the 32
//             lanes of a warp walking their own subsequences in lock step hit 32 different banks without padding.  The
//             window of the team's NEXT chunk is fetched by one 1-D bulk asynchronous copy (cp.async.bulk + mbarrier
//             complete_tx; SASS UBLKCP / SYNCS) issued when the current window is dead; a byte-swap pass in shared
//             memory makes it MSB-first.
//   phase A   entry candidate by self-synchronisation from a look-back window (as in hb_decode.cuh).
//   phase B   COUNT: walk entry .. end of the subsequence with the multi-letter EMIT table (entry = up to 3 letters |
//             count | bits consumed; the count pass only uses the top byte): letters and exit.
//   verify    entry[t] == exit[t-1] inside the team, iterated to a fixed point (a refuted thread counts again).
//   offsets   team scan of the letter counts; the team's first output position comes from a decoupled look-back over
//             per-chunk descriptors (status | exit | count), which also checks entry == predecessor's exit ACROSS chunks.
//             A mismatch only raises a flag: the host then falls back to the two-pass decoder (exact, with serial repair).
//   phase C   EMIT: every thread owns the 32-byte output rows that START inside its letters.  It walks its subsequence a
//             second time (shared memory only), appends up to three letters per lookup to an accumulator word, stores
//             completed words to a private 64-byte ring in shared memory with predicated stores, and after every ten
//             lookups copies a finished row from the ring to global memory with one 256-bit store.  It runs on past its
//             subsequence to complete its last row (its successor skips those letters).
//
// No per-thread output slots: shared memory per thread is 132 B of stream + 68 B of ring, whatever the code lengths, so
// any tree whose codes fit the emit table is eligible and four teams fit an SM.
// Eligible: max_len <= kEmitBits, >= 2 leaves, no duplicate letters, known entry of the first code word.
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
#ifndef HB_FUSED_TEAM
#define HB_FUSED_TEAM 256
#endif
constexpr int kFSubWords = HB_FUSED_SUB_WORDS;                 // words per subsequence: ODD (bank-conflict-free dense layout)
static_assert(kFSubWords % 2 == 1, "subsequence length must be an odd number of words");
constexpr int kFSubBits = kFSubWords * 32;
constexpr int kFTeam = HB_FUSED_TEAM;                          // threads per team
constexpr int kFMaxTeams = 1024 / kFTeam;
static_assert(kFTeam % 32 == 0 && kFTeam >= 64 && kFMaxTeams <= 15, "teams synchronise on named barriers 1..15");
constexpr int kFChunkWords = kFTeam * kFSubWords;
static_assert(kFChunkWords % 4 == 0, "chunks must keep 16-byte alignment");
constexpr int kFHalo = 16;                                     // words staged before the chunk (leading look-back)
constexpr int kFHaloAfter = 32;                                // words staged after it (the last thread completes its row)
constexpr int kFWinWords = kFHalo + kFChunkWords + kFHaloAfter;
constexpr int kFWinAlloc = kFWinWords + 4;                     // + look-ahead slack
constexpr uint32_t kFWinBits = kFWinWords * 32u;
constexpr int kEmitBits = HB_EMIT_BITS;                         // index width of the emit table: trees whose codes fit it ...
constexpr int kEmitBitsWide = 14;                              // ... and a second kernel instance for codes of up to 14 bits
                                                               // (64 KiB table, three teams per SM instead of four)
#ifndef HB_FUSED_LOOKBACK_BITS
#define HB_FUSED_LOOKBACK_BITS 320       // in-team look-back: a refuted thread costs its whole team (and, through the
#endif                                   // look-back, every later chunk) a second count, so it is longer than the two-pass one
constexpr int kFLookbackBits = HB_FUSED_LOOKBACK_BITS;
#ifndef HB_FUSED_EMIT_TRIPS
#define HB_FUSED_EMIT_TRIPS 4
#endif
constexpr int kFEmitTrips = HB_FUSED_EMIT_TRIPS;               // two lookups per trip; a row check after every 10 lookups
constexpr int kFRingWords = 16, kFRingStride = 16;             // 64-byte ring per thread, stored WORD-MAJOR: word k of thread t at
                                                               // [k][t], so a lane always hits bank t % 32 wherever it is in its ring
// letters a thread may decode beyond the end of its last row before it notices: one block of lookups
constexpr int kFOverrunLetters = 31 + 6 * kFEmitTrips;
static_assert(kEmitBits <= 13 && kEmitBits >= 8 && kEmitBitsWide <= 15, "emit table index width (4-bit length field)");
static_assert(kFHalo * 32 >= HB_LEAD_LOOKBACK_BITS, "the leading look-back must fit the halo");
static_assert(kFHaloAfter * 32 >= kFOverrunLetters * kEmitBitsWide + 2 * kEmitBitsWide + 96, "the last thread's overrun must fit the halo");
static_assert(31 + 6 * kFEmitTrips <= 4 * kFRingWords - 3, "pending letters + one block must fit the ring");

constexpr uint64_t kDescAgg = 1ull << 62, kDescPrefix = 2ull << 62;
constexpr int kDescExitShift = 42;
constexpr uint64_t kDescValueMask = (1ull << kDescExitShift) - 1;
constexpr uint32_t kDescExitEnd = 0xFFFFFu;

struct FusedResult {
    unsigned long long total_letters;
    unsigned long long entry0;         // absolute buffer bit, kEnd64 = none
    unsigned long long exit_last;
    uint32_t error;                    // bit 0: a chunk's entry != its predecessor's exit (speculation refuted)
    uint32_t careful_threads;          // threads that wrote their rows letter by letter (stream ends, ragged head)
    unsigned long long phase_cycles[8]; // -DHB_FUSED_TIMING: SM clock cycles summed over all chunks (first thread of each
                                       // team): stage, phase A + B, verify, scan, look-back, emit; [6] = chunks
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
    uint32_t spoil_speculation;
    uint32_t k1, k4;                   // the constants 1 and 4 as PARAMETERS: register moves and small adds written as
                                       // multiply-adds by them stay on the FMA pipe (the compiler cannot fold them back
                                       // into ALU-pipe SEL / IADD; the decode loops are ALU-pipe bound: ncu pipe_alu 72 %)
    const uint32_t *emit;              // 1 << kEmitBits entries
    const uint8_t *code_len;           // 256 bytes
    unsigned long long *desc;          // n_chunks, zeroed before the launch
    uint32_t *ticket;                  // zeroed before the launch
    FusedResult *result;               // zeroed before the launch
    uint8_t *out;
    uint64_t out_cap;
};

// shared memory: per CTA the emit table + code lengths; per team window + rings + exits + misc
__host__ __device__ constexpr size_t fused_team_bytes() {
    return static_cast<size_t>(kFWinAlloc) * 4 + 48 + static_cast<size_t>(kFTeam) * kFRingStride * 4 + kFTeam * 4 + 128;
}
__host__ __device__ constexpr size_t fused_shared_bytes(int emit_bits) { return (static_cast<size_t>(1) << emit_bits) * 4 + 256; }
static_assert(fused_team_bytes() % 64 == 0 && fused_shared_bytes(kEmitBits) % 64 == 0 && (kFWinAlloc * 4 + 48) % 64 == 0,
              "64-byte alignment of the team blocks and of the rings");

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(kFTeam) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
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
    uint32_t k1, k4;                   // FusedParams::k1 / k4
    __device__ __forceinline__ void init(uint32_t win, uint32_t q0, uint32_t one = 1u, uint32_t four = 4u) {
        k1 = one;
        k4 = four;
        q = q0;
        wa = win + ((q0 >> 5) << 2);
        w0 = lds32(wa);
        w1 = lds32(wa + 4);
        w2 = lds32(wa + 8);
        wa += 12;
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(w1, w0, q); }
    __device__ __forceinline__ void refill(uint32_t qn) {
        const uint32_t r = (qn ^ q) & 32u;
        asm volatile("{\n\t.reg .pred f;\n\tsetp.ne.u32 f, %4, 0;\n\t@f mov.u32 %0, %1;\n\t@f mov.u32 %1, %2;\n\t"
                     "@f ld.shared.u32 %2, [%3];\n\t@f add.u32 %3, %3, 4;\n\t}"
                     : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(wa) : "r"(r) : "memory");
        q = qn;
    }
    __device__ __forceinline__ void step(uint32_t bits) { refill(q + bits); }
    // hot loops: the predicated register moves and the address bump as multiply-adds (FMA pipe)
    __device__ __forceinline__ void step_fma(uint32_t bits) {
        const uint32_t qn = q + bits;
        const uint32_t r = (qn ^ q) & 32u;
        asm volatile("{\n\t.reg .pred f;\n\tsetp.ne.u32 f, %4, 0;\n\t@f mad.lo.u32 %0, %1, %5, 0;\n\t@f mad.lo.u32 %1, %2, %5, 0;\n\t"
                     "@f ld.shared.u32 %2, [%3];\n\t@f mad.lo.u32 %3, %6, %5, %3;\n\t}"
                     : "+r"(w0), "+r"(w1), "+r"(w2), "+r"(wa) : "r"(r), "r"(k1), "r"(k4) : "memory");
        q = qn;
    }
};

// byte offset of the emit-table entry for the next EB bits
template <int EB>
__device__ __forceinline__ uint32_t emit_off(uint32_t x) {
    uint32_t y;
    asm("and.b32 %0, %1, %2;" : "=r"(y) : "r"(x), "n"(~((1u << (32 - EB)) - 1u)));
    return y >> (30 - EB);
}

// Advance from q over whole code words: first code-word start >= q_stop, kEnd32 if a code word does not end at or before
// q_avail.  `letters` counts the code words passed.  (Phase A and the count pass.)
template <int EB>
__device__ __forceinline__ uint32_t fused_run(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_stop,
                                              uint32_t q_avail, uint32_t &letters, uint32_t k1 = 1u, uint32_t k4 = 4u) {
    letters = 0;
    if (q == kEnd32) return kEnd32;
    if (q >= q_stop) return q;
    FReader rd;
    rd.init(win, q, k1, k4);
    uint32_t acc = 0;                                      // sum of the entries' top bytes: bits << 4 | count
    const uint32_t q_begin = q;
    const uint32_t lim = min(q_stop, q_avail);
    if (lim >= 2u * EB) {
        const uint32_t last2 = lim - 2 * EB;        // both lookups of a trip start at or before lim - kEmitBits
#pragma unroll 1
        while (rd.q <= last2) {
            // two lookups per peek: the 32 peeked bits always hold a second kEmitBits-bit window behind the first
            // entry's <= kEmitBits bits, so one refill test serves both
            const uint32_t x = rd.peek();
            const uint32_t e1 = lds32(lut + emit_off<EB>(x));
            const uint32_t b1 = e1 >> 28;
            const uint32_t e2 = lds32(lut + emit_off<EB>(x << b1));
            rd.step_fma(b1 + (e2 >> 28));
            acc += (e1 >> 24) + (e2 >> 24);
        }
    }
    uint32_t n1 = 0, ret = 0;
#pragma unroll 1
    for (;;) {                                             // the last few letters before q_stop, one at a time
        if (rd.q >= q_stop) { ret = rd.q; break; }
        const uint32_t e = lds32(lut + emit_off<EB>(rd.peek()));
        const uint32_t len = lds8(lens + (e & 0xFFu));
        if (rd.q + len > q_avail) { ret = kEnd32; break; }
        rd.step(len);
        acc += len << 4;
        n1++;
    }
    letters = acc - ((rd.q - q_begin) << 4) + n1;
    return ret;
}

// Careful emit (threads near the end of the owned stream range, and the thread with the ragged head of the whole output):
// letter by letter with every check, straight into registers: 32 letters, one 256-bit store per output row.
template <int EB>
__device__ __noinline__ void fused_careful_rows(uint32_t win, uint32_t lut, uint32_t lens, uint32_t entry, uint32_t q_own_end,
                                                uint32_t q_avail, uint32_t count, uint8_t *out, uintptr_t out_addr, uint64_t D,
                                                uint64_t out_cap) {
    if (entry == kEnd32 || count == 0) return;
    FReader rd;
    rd.init(win, entry);
    auto next = [&](uint32_t &letter) -> bool {                // false: no further letter of the owned range exists
        if (rd.q >= q_own_end) return false;
        const uint32_t e = lds32(lut + emit_off<EB>(rd.peek()));
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
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int b8 = 0; b8 < 4; b8++)
                    if (static_cast<uint32_t>(4 * k + b8) < got && pos + 4 * k + b8 < out_cap)
                        out[pos + 4 * k + b8] = static_cast<uint8_t>(v[k] >> (8 * b8));
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

extern __shared__ __align__(128) uint8_t fused_smem[];

// -DHB_FUSED_TIMING: per-phase SM clock stamps (hb_ctx_fused_phase_cycles); off in the product build
#ifdef HB_FUSED_TIMING
#define HB_TK(i) tk[i] = clock()
#else
#define HB_TK(i) ((void)0)
#endif

// s_misc words
enum { kMWarp = 0, kMFlag = 8, kMBaseLo = 10, kMBaseHi = 11, kMExit = 13, kMEntry = 14, kMNext = 15,
       kMBar = 16 /* 8 bytes */, kMPrefetched = 18 };

template <int EB>
__global__ void __launch_bounds__(kFTeam * kFMaxTeams, 1)
dec_fused_kernel(const FusedParams p) {
    const int team = threadIdx.x / kFTeam, tt = threadIdx.x % kFTeam;
    const int lane = tt & 31;
    uint32_t *s_lut = reinterpret_cast<uint32_t *>(fused_smem);
    uint8_t *s_lens = fused_smem + (static_cast<size_t>(1) << EB) * 4;
    uint8_t *tb = fused_smem + fused_shared_bytes(EB) + team * fused_team_bytes();
    uint32_t *s_win = reinterpret_cast<uint32_t *>(tb);
    uint32_t *s_ring = s_win + kFWinAlloc + 12;             // 64-byte aligned
    uint32_t *s_exit = s_ring + kFTeam * kFRingStride;
    uint32_t *s_misc = s_exit + kFTeam;

    for (int i = threadIdx.x; i < (1 << EB); i += blockDim.x) s_lut[i] = p.emit[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_lens)[i] = reinterpret_cast<const uint32_t *>(p.code_len)[i];

    uint32_t b = smem_addr(fused_smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(b));             // one opaque base register (see hb_decode.cuh)
    const uint32_t a_lut = b;
    const uint32_t a_lens = b + (1u << EB) * 4u;
    const uint32_t a_win = b + static_cast<uint32_t>(fused_shared_bytes(EB) + team * fused_team_bytes());
    // ring word k of this thread is at ring_t + k * (4 * kFTeam)
    const uint32_t ring_t = a_win + kFWinAlloc * 4u + 48u + static_cast<uint32_t>(tt) * 4u;
    const uint32_t a_bar = smem_addr(s_misc + kMBar);
    const uintptr_t out_addr = reinterpret_cast<uintptr_t>(p.out);
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
    uint32_t tk[7];
#endif

    for (;;) {
        const uint32_t j = s_misc[kMNext];                 // chunk, relative to first_chunk (ticket order)
        const bool prefetched = s_misc[kMPrefetched] != 0;
        if (j >= p.n_chunks) break;
        HB_TK(0);
        const uint32_t chunk = p.first_chunk + j;

        // ---- stage [chunk * kFChunkWords - kFHalo, + kFWinWords) MSB-first
        if (prefetched) {
            // the bulk copy was issued at the end of the previous chunk: wait for its bytes, then byte-swap in place
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
                entry = fused_run<EB>(a_win, a_lut, a_lens, q0, q_lo, q_avail, dummy, p.k1, p.k4);
            }
        }

        // ---- phase B (count) + in-team verification to a fixed point
        uint32_t count = 0, exitq = kEnd32;
        bool redo = active;
        for (int round = 0;; round++) {
            if (round > kFTeam + 1) asm volatile("trap;");
            if (redo) exitq = fused_run<EB>(a_win, a_lut, a_lens, entry, q_hi, q_avail, count, p.k1, p.k4);
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
        HB_TK(3);

        // ---- team scan of the letter counts; the ticket of the next chunk
        const uint32_t incl = warp_incl_scan(count);
        if (lane == 31) s_misc[kMWarp + (tt >> 5)] = incl;
        const bool is_last_active = active && (tt == kFTeam - 1 || q_sub + kFSubBits >= q_own_end);
        if (is_last_active) s_misc[kMExit] = exitq;
        if (active && !has_pred) s_misc[kMEntry] = entry;
        if (tt == 0) s_misc[kMNext] = atomicAdd(p.ticket, 1u);
        team_sync(team);
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
                // 128 descriptors per trip (4 per lane, all loads in flight together)
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
        team_sync(team);
        HB_TK(5);
        const uint64_t base = (static_cast<uint64_t>(s_misc[kMBaseHi]) << 32) | s_misc[kMBaseLo];
        const uint64_t D = base + off;                          // output position of my first letter

        // ---- phase C (emit): my rows = the 32-byte rows of the output that start inside my letters
        if (count) {
            const uint32_t p0 = static_cast<uint32_t>((out_addr + D) & 31);        // where my first letter sits in its row
            const uint32_t p_end = p0 + count;                                      // row space: my letters are [p0, p_end)
            const uint32_t target = (p_end + 31u) & ~31u;                           // my last row ends here
            // the fast loop may read (target - p_end) + one block of letters past my own: all of it must be real stream
            const uint32_t reach = q_hi + (kFOverrunLetters + 1) * EB + 64;
            const bool careful = reach > min(q_avail, q_own_end) || (D == 0 && p0 != 0) ||
                                 static_cast<uint64_t>(D - p0) + target > p.out_cap;
            if (careful) {
                fused_careful_rows<EB>(a_win, a_lut, a_lens, entry, q_own_end, q_avail, count, p.out, out_addr, D, p.out_cap);
                atomicAdd(&p.result->careful_threads, 1u);
            } else if (target > 32u || p0 == 0) {
                FReader rd;
                rd.init(a_win, entry, p.k1, p.k4);
                const uint32_t k1 = p.k1;
                uint32_t acc = 0;
                uint32_t pp = p0;                              // letter position, exact in its low 4 bits
                uint32_t wp = p0 & ~3u;                        // bytes of completed words (row space)
                uint32_t flushed = p0 ? 32u : 0u;              // rows below this are written (row 0 is my predecessor's)
                uint8_t *row_ptr = p.out + (D - p0) + flushed;
                auto append = [&](uint32_t e) {
                    const uint32_t L = e & 0xFFFFFFu;
                    const uint32_t s = pp << 3;                // funnel shifts take the amount modulo 32
                    acc |= __funnelshift_l(0u, L, s);          // L << (8 * (pp & 3))
                    const uint32_t hi = __funnelshift_l(L, 0u, s);
                    const uint32_t pn = pp + (e >> 24);        // adds count (bits 0..1) + junk above bit 3
                    const uint32_t fl = (pn ^ pp) & 4u;        // <= 3 letters per append: at most one word completes
                    const uint32_t wa_ring = ring_t + (wp & 60u) * static_cast<uint32_t>(kFTeam);   // word (wp / 4) % 16, word-major
                    asm volatile("{\n\t.reg .pred f;\n\tsetp.ne.u32 f, %3, 0;\n\t@f st.shared.u32 [%2], %0;\n\t"
                                 "@f mad.lo.u32 %0, %1, %4, 0;\n\t}"
                                 : "+r"(acc) : "r"(hi), "r"(wa_ring), "r"(fl), "r"(k1) : "memory");
                    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(wp) : "r"(fl), "r"(k1));   // wp += fl
                    pp = pn;
                };
#pragma unroll 1
                while (flushed < target) {
#pragma unroll
                    for (int t2 = 0; t2 < kFEmitTrips; t2++) {
                        const uint32_t x = rd.peek();
                        const uint32_t e1 = lds32(a_lut + emit_off<EB>(x));
                        const uint32_t b1 = e1 >> 28;
                        const uint32_t e2 = lds32(a_lut + emit_off<EB>(x << b1));
                        rd.step_fma(b1 + (e2 >> 28));
                        append(e1);
                        append(e2);
                    }
                    if (wp >= flushed + 32u) {                 // a row is complete in the ring: out it goes
                        const uint32_t half = flushed & 32u;
                        uint32_t v[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) v[k] = lds32(ring_t + (half + 4u * k) * static_cast<uint32_t>(kFTeam));
                        stg256(row_ptr, v);
                        row_ptr += 32;
                        flushed += 32;
                    }
                }
            }
        }
        team_sync(team);                                        // the window and the rings are reused
        // ---- fetch the next chunk's window in bulk while the other teams of the CTA compute
        if (tt == 0) {
            const bool will_prefetch = next_j < p.n_chunks && bulk_ok(next_j);
            s_misc[kMPrefetched] = will_prefetch ? 1u : 0u;
            if (will_prefetch) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses before async writes
                const long long wb = static_cast<long long>(p.first_chunk + next_j) * kFChunkWords - kFHalo;
                mbar_expect_tx(a_bar, kFWinWords * 4u);
                bulk_g2s(a_win, p.words + wb, kFWinWords * 4u, a_bar);
            }
#ifdef HB_FUSED_TIMING
            HB_TK(6);
#pragma unroll
            for (int k = 0; k < 6; k++) atomicAdd(&p.result->phase_cycles[k], static_cast<unsigned long long>(tk[k + 1] - tk[k]));
            atomicAdd(&p.result->phase_cycles[6], 1ull);
#endif
        }
        team_sync(team);                                        // kMNext / kMPrefetched are read at the top
    }
}

}  // namespace hb
