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
constexpr int kFTeam = 256;                                    // threads per team
constexpr int kFMaxTeams = 4;
constexpr int kFChunkWords = kFTeam * kFSubWords;
static_assert(kFChunkWords % 4 == 0, "chunks must keep 16-byte alignment");
constexpr int kFHalo = 16;                                     // words staged before and after the chunk
constexpr int kFWinWords = kFHalo + kFChunkWords + kFHalo;
constexpr int kFWinAlloc = kFWinWords + 4;                     // + look-ahead slack
constexpr uint32_t kFWinBits = kFWinWords * 32u;
constexpr int kEmitBits = HB_EMIT_BITS;
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
    return static_cast<size_t>(kFWinAlloc) * 4 + static_cast<size_t>(kFTeam) * slot_words * 4 + kFTeam * 4 * 2 + 64;
}
__host__ __device__ constexpr size_t fused_shared_bytes() { return (static_cast<size_t>(1) << kEmitBits) * 4 + 256; }

__device__ __forceinline__ void team_sync(int team) {
    asm volatile("bar.sync %0, %1;" :: "r"(team + 1), "r"(kFTeam) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_ull(const unsigned long long *p) {
    unsigned long long r;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_release_ull(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

// Bit reader over the DENSE staged window: two words + position; the funnel shift takes the position modulo 32 and a
// refill is due exactly when bit 5 of the position flips (a step is < 32 bits).
struct FReader {
    uint32_t w0, w1, q, wa;            // wa: shared address of the next word to load
    __device__ __forceinline__ void init(uint32_t win, uint32_t q0) {
        q = q0;
        wa = win + ((q0 >> 5) << 2);
        w0 = lds32(wa);
        w1 = lds32(wa + 4);
        wa += 8;
    }
    __device__ __forceinline__ uint32_t peek() const { return __funnelshift_l(w1, w0, q); }
    __device__ __forceinline__ void step(uint32_t bits) {
        const uint32_t qn = q + bits;
        if ((qn ^ q) & 32u) {
            w0 = w1;
            w1 = lds32(wa);
            wa += 4;
        }
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
// at or before q_avail.
__device__ __forceinline__ uint32_t fused_run(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_stop,
                                              uint32_t q_avail) {
    if (q >= q_stop) return q;
    FReader rd;
    rd.init(win, q);
    const uint32_t lim = min(q_stop, q_avail);
    if (lim >= static_cast<uint32_t>(kEmitBits)) {
        const uint32_t last = lim - kEmitBits;
        while (rd.q <= last) rd.step(lds32(lut + emit_off(rd.peek())) >> 28);
    }
    while (rd.q < q_stop) {
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        const uint32_t len = lds8(lens + (e & 0xFFu));
        if (rd.q + len > q_avail) return kEnd32;
        rd.step(len);
    }
    return rd.q;
}

// Letter sink of one thread: accumulator word + letter count + slot address.
struct FSink {
    uint32_t acc, p, wp, wend;
    __device__ __forceinline__ void init(uint32_t slot, uint32_t slot_end) { acc = 0; p = 0; wp = slot; wend = slot_end; }
    // append cnt (<= 3) letters packed in the low bytes of L (zero above)
    __device__ __forceinline__ void append(uint32_t L, uint32_t cnt) {
        const uint32_t s = p << 3;                         // shift amounts are taken modulo 32
        acc |= L << (s & 31u);
        const uint32_t hi = __funnelshift_l(L, 0u, s);     // letters that fall past the accumulator word
        const uint32_t pn = p + cnt;
        if ((pn ^ p) & 4u) {
            if (wp < wend) sts32(wp, acc);
            wp += 4;
            acc = hi;
        }
        p = pn;
    }
    __device__ __forceinline__ void finish() { if ((p & 3u) && wp < wend) sts32(wp, acc); }
};

// Decode [entry, q_hi) into the thread's slot.  Returns the exit (first code-word start >= q_hi, kEnd32 at the end of
// the stream); count = letters whose code word starts in [entry, q_hi) and ends <= q_avail.
__device__ __forceinline__ uint32_t fused_emit(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_hi,
                                               uint32_t q_avail, uint32_t slot, uint32_t slot_end, uint32_t &count) {
    count = 0;
    if (q == kEnd32) return kEnd32;
    if (q >= q_hi) return q;
    FReader rd;
    rd.init(win, q);
    FSink sk;
    sk.init(slot, slot_end);
    const uint32_t lim = min(q_hi, q_avail);
    if (lim >= static_cast<uint32_t>(kEmitBits)) {
        const uint32_t last = lim - kEmitBits;
        // two lookups per trip: the second one may start up to kEmitBits past `last`, which is still < lim
        if (last >= static_cast<uint32_t>(kEmitBits)) {
            const uint32_t last2 = last - kEmitBits;
            while (rd.q <= last2) {
                const uint32_t e1 = lds32(lut + emit_off(rd.peek()));
                rd.step(e1 >> 28);
                sk.append(e1 & 0xFFFFFFu, (e1 >> 24) & 3u);
                const uint32_t e2 = lds32(lut + emit_off(rd.peek()));
                rd.step(e2 >> 28);
                sk.append(e2 & 0xFFFFFFu, (e2 >> 24) & 3u);
            }
        }
        while (rd.q <= last) {
            const uint32_t e = lds32(lut + emit_off(rd.peek()));
            rd.step(e >> 28);
            sk.append(e & 0xFFFFFFu, (e >> 24) & 3u);
        }
    }
    uint32_t exitq = 0;
    for (;;) {
        if (rd.q >= q_hi) { exitq = rd.q; break; }
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        const uint32_t letter = e & 0xFFu;
        const uint32_t len = lds8(lens + letter);
        if (rd.q + len > q_avail) { exitq = kEnd32; break; }
        rd.step(len);
        sk.append(letter, 1u);
    }
    sk.finish();
    count = sk.p;
    return exitq;
}

// The < 32 letters that complete a thread's last output row: from the successors' slots, then -- past the team's last
// letter -- decoded from the halo behind the chunk.  Appended bytewise to the thread's own slot at offset `cnt`.
// Returns how many letters were appended (< need only at the end of the owned stream range).
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

// Slow path of a chunk whose slots overflowed: decode again, letter by letter, straight to global memory.
__device__ __noinline__ void fused_slow_write(uint32_t win, uint32_t lut, uint32_t lens, uint32_t q, uint32_t q_limit,
                                              uint32_t q_avail, uint32_t n, uint8_t *out, uint64_t pos, uint64_t out_cap) {
    if (q == kEnd32 || n == 0) return;
    FReader rd;
    rd.init(win, q);
    for (uint32_t i = 0; i < n && rd.q < q_limit; i++) {
        const uint32_t e = lds32(lut + emit_off(rd.peek()));
        const uint32_t letter = e & 0xFFu;
        const uint32_t len = lds8(lens + letter);
        if (rd.q + len > q_avail) break;
        rd.step(len);
        if (pos + i < out_cap) out[pos + i] = static_cast<uint8_t>(letter);
    }
}

extern __shared__ __align__(16) uint8_t fused_smem[];

__global__ void __launch_bounds__(kFTeam * kFMaxTeams, 1)
dec_fused_kernel(const FusedParams p) {
    const int team = threadIdx.x / kFTeam, tt = threadIdx.x % kFTeam;
    const int lane = tt & 31;
    // shared by the CTA: emit table + code lengths; then one block per team
    uint32_t *s_lut = reinterpret_cast<uint32_t *>(fused_smem);
    uint8_t *s_lens = fused_smem + (static_cast<size_t>(1) << kEmitBits) * 4;
    const size_t team_bytes = fused_team_bytes(p.slot_words);
    uint8_t *tb = fused_smem + fused_shared_bytes() + team * team_bytes;
    uint32_t *s_win = reinterpret_cast<uint32_t *>(tb);
    uint32_t *s_slots = s_win + kFWinAlloc;
    uint32_t *s_exit = s_slots + static_cast<size_t>(kFTeam) * p.slot_words;
    uint32_t *s_cnt = s_exit + kFTeam;
    uint32_t *s_misc = s_cnt + kFTeam;                      // [0..7] warp sums, [8] flag, [9] chunk, [10..11] base, [12] flag2

    for (int i = threadIdx.x; i < (1 << kEmitBits); i += blockDim.x) s_lut[i] = p.emit[i];
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
        reinterpret_cast<uint32_t *>(s_lens)[i] = reinterpret_cast<const uint32_t *>(p.code_len)[i];
    __syncthreads();

    uint32_t b = smem_addr(fused_smem);
    asm volatile("mov.u32 %0, %0;" : "+r"(b));             // one opaque base register (see hb_decode.cuh)
    const uint32_t a_lut = b;
    const uint32_t a_lens = b + (1u << kEmitBits) * 4u;
    const uint32_t a_win = b + static_cast<uint32_t>(fused_shared_bytes() + team * team_bytes);
    const uint32_t a_slots = a_win + kFWinAlloc * 4u;
    const uint32_t slot_bytes = p.slot_words * 4u;
    const uint32_t a_slot = a_slots + tt * slot_bytes;
    const uintptr_t out_addr = reinterpret_cast<uintptr_t>(p.out);

    for (;;) {
        if (tt == 0) s_misc[9] = atomicAdd(p.ticket, 1u);
        team_sync(team);
        const uint32_t j = s_misc[9];                      // chunk, relative to first_chunk
        if (j >= p.n_chunks) break;
        const uint32_t chunk = p.first_chunk + j;

        // ---- stage [chunk * kFChunkWords - kFHalo, + kFWinWords) MSB-first; all loads of a thread in flight together
        {
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
            if (tt < 4) s_win[kFWinWords + tt] = 0;
        }
        team_sync(team);

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
                uint32_t window = tt == 0 ? static_cast<uint32_t>(kLeadLookbackBits) : static_cast<uint32_t>(kLookbackBits);
                if (p.fixed_len) window = 0;
                if (p.spoil_speculation && tt == 0) window = 0;
                uint32_t q0 = q_lo > window ? q_lo - window : 0;
                if (q0 < q_buf0) q0 = q_buf0;
                if (p.len_gcd > 1) {
                    const unsigned long long abs0 = static_cast<unsigned long long>(win_bit0 + q0) + p.stream_bit0;
                    const uint32_t rem = static_cast<uint32_t>(abs0 % p.len_gcd);
                    if (rem) q0 += p.len_gcd - rem;
                }
                entry = fused_run(a_win, a_lut, a_lens, q0, q_lo, q_avail);
            }
        }

        // ---- phase B + in-team verification to a fixed point
        uint32_t count = 0, exitq = kEnd32;
        bool redo = active;
        for (int round = 0;; round++) {
            if (round > kFTeam + 1) asm volatile("trap;");
            if (redo) exitq = fused_emit(a_win, a_lut, a_lens, entry, q_hi, q_avail, a_slot, a_slot + slot_bytes, count);
            s_exit[tt] = exitq;
            if (tt == 0) s_misc[8] = 0;
            team_sync(team);
            redo = false;
            if (has_pred) {
                const uint32_t want = s_exit[tt - 1];
                if (want != entry) { entry = want; redo = true; s_misc[8] = 1; }
            }
            team_sync(team);
            const uint32_t any = s_misc[8];
            team_sync(team);
            if (!any) break;
        }
        if (!active) count = 0;

        // ---- team scan of the letter counts; overflow flag
        const uint32_t cap_bytes = slot_bytes;
        const bool overflow = count + 32u > cap_bytes;
        const uint32_t incl = warp_incl_scan(count);
        if (lane == 31) s_misc[tt >> 5] = incl;
        s_cnt[tt] = count;
        if (tt == 0) s_misc[12] = 0;
        team_sync(team);
        if (overflow) s_misc[12] = 1;
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < kFTeam / 32; k++) { const uint32_t w = s_misc[k]; if (k < (tt >> 5)) before += w; total += w; }
        const uint32_t off = before + incl - count;

        // the team's last active thread and its exit
        const bool is_last_active = active && (tt == kFTeam - 1 || q_sub + kFSubBits >= q_own_end);
        if (is_last_active) s_misc[13] = exitq;
        if (active && !has_pred) s_misc[14] = entry;
        team_sync(team);
        const uint32_t q_chunk_exit = s_misc[13];
        const bool slow = s_misc[12] != 0;

        // ---- decoupled look-back (first warp of the team): exclusive letter offset of this chunk
        if (tt < 32) {
            const uint32_t q_chunk_end = (kFHalo + kFChunkWords) * 32u;
            const uint32_t exit_rel = q_chunk_exit == kEnd32 ? kDescExitEnd : (q_chunk_exit - q_chunk_end) & 0xFFFFFu;
            const uint32_t q_entry = s_misc[14];
            const unsigned long long mine = (static_cast<unsigned long long>(exit_rel) << kDescExitShift) | total;
            if (lane == 0 && j + 1 < p.n_chunks) st_release_ull(p.desc + j, kDescAgg | mine);
            unsigned long long excl = 0;
            if (j > 0) {
                long long look = static_cast<long long>(j) - 1;
                bool first_window = true;
                for (;;) {
                    const long long idx = look - lane;
                    unsigned long long d;
                    do {
                        d = idx >= 0 ? ld_acquire_ull(p.desc + idx) : kDescPrefix;
                    } while (__any_sync(0xFFFFFFFFu, (d >> 62) == 0));
                    if (first_window && lane == 0) {
                        // my entry must be where the chunk before me stopped
                        const uint32_t pred_exit = static_cast<uint32_t>(d >> kDescExitShift) & 0xFFFFFu;
                        const uint32_t entry_rel = q_entry == kEnd32 ? kDescExitEnd : (q_entry - kFHalo * 32u) & 0xFFFFFu;
                        if (pred_exit != entry_rel) atomicOr(&p.result->error, 1u);
                    }
                    first_window = false;
                    const uint32_t pref_mask = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                    const int stop = pref_mask ? __ffs(pref_mask) - 1 : 31;
                    unsigned long long v = lane <= stop ? (d & kDescValueMask) : 0ull;
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, sft);
                    excl += v;
                    if (pref_mask) break;
                    look -= 32;
                }
            }
            if (lane == 0) {
                if (j + 1 < p.n_chunks) {
                    st_release_ull(p.desc + j, kDescPrefix | (static_cast<unsigned long long>(exit_rel) << kDescExitShift) |
                                               ((excl + total) & kDescValueMask));
                } else {
                    p.result->total_letters = excl + total;
                    p.result->exit_last = q_chunk_exit == kEnd32 ? kEnd64 : static_cast<unsigned long long>(win_bit0 + q_chunk_exit);
                }
                if (j == 0) p.result->entry0 = q_entry == kEnd32 ? kEnd64 : static_cast<unsigned long long>(win_bit0 + q_entry);
                if (slow) atomicAdd(&p.result->slow_chunks, 1u);
                s_misc[10] = static_cast<uint32_t>(excl);
                s_misc[11] = static_cast<uint32_t>(excl >> 32);
            }
        }
        team_sync(team);
        const uint64_t base = (static_cast<uint64_t>(s_misc[11]) << 32) | s_misc[10];
        const uint64_t D = base + off;                          // output position of my first letter

        // ---- compaction
        if (slow) {
            // a slot overflowed somewhere in the team: every thread writes its own letters letter by letter; the team's
            // last active thread also completes the last row (the next chunk expects it to be written)
            if (active && count) fused_slow_write(a_win, a_lut, a_lens, entry, q_hi, q_avail, count, p.out, D, p.out_cap);
            if (is_last_active && exitq != kEnd32) {
                const uint64_t end = D + count;
                const uint32_t need = static_cast<uint32_t>((0 - (out_addr + end)) & 31);
                fused_slow_write(a_win, a_lut, a_lens, exitq, q_own_end, q_avail, need, p.out, end, p.out_cap);
            }
        } else if (count) {
            const uint64_t end = D + count;
            const uint32_t head = D == 0 ? 0u : static_cast<uint32_t>((0 - (out_addr + D)) & 31);   // letters my predecessor's row holds
            const uint32_t need = static_cast<uint32_t>((0 - (out_addr + end)) & 31);
            uint32_t cnt_eff = count;
            if (need && head < count + need)
                cnt_eff += fused_pull(a_slots, slot_bytes, s_cnt, tt, count, need, a_win, a_lut, a_lens, q_chunk_exit,
                                      q_own_end, q_avail);
            uint32_t o = head;
            if (D == 0) {                                       // nobody precedes the first letter: its ragged head is mine
                const uint32_t rag = static_cast<uint32_t>((0 - out_addr) & 31);
                for (; o < rag && o < cnt_eff; o++)
                    if (o < p.out_cap) p.out[o] = static_cast<uint8_t>(lds8(a_slot + o));
            }
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
        team_sync(team);                                        // slots and window are reused by the next chunk
    }
}

}  // namespace hb
