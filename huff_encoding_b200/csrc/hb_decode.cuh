// hb_decode.cuh -- K3: decompress (comp.rs:487-519) of the reference's unindexed bit stream.
//
// The reference walks the tree bit by bit from the first bit; the stream carries no index and no length, so a
// parallel decoder has to find code-word boundaries itself.  Self-synchronising speculative decode:
//
//   stream bits are cut into 1024-bit subsequences (one per thread), 256 subsequences (32 KB) per CTA.
//   count pass   (dec_count_kernel)  CTA stages its 32 KB (+halo) in shared memory with coalesced loads.
//       phase A  every thread finds a candidate for its ENTRY (first code-word start >= its subsequence) by decoding
//                a look-back window that starts W bits earlier, aligned to gcd(code lengths): prefix codes
//                resynchronise, so after W bits the decoder is on a true boundary with high probability
//                (fixed-length code sets need no window: every aligned position is a boundary).
//       phase B  decode from the entry to the end of the subsequence: letter COUNT and EXIT (= the next
//                subsequence's true entry *if* this entry was right).
//       verify   entry[t] must equal exit[t-1]; threads whose candidate was wrong adopt exit[t-1] and redo B.
//                Iterated to a fixed point inside the CTA (terminates: thread t is final after t rounds).
//   chain check  (dec_verify_kernel) the same test across CTA boundaries; mismatching CTAs are re-run serially
//                with the corrected entry by dec_fix_kernel (rare: a CTA's first thread looks back 384 bits, and the
//                measured miss rate falls ~27x per 64 bits from 1e-4 at 192; worst case = strictly sequential
//                propagation, still exact).
//   offsets      two small scan kernels turn per-CTA letter counts into 64-bit output offsets.
//   write pass   (dec_write_kernel)  no shared-memory staging of the output: thread t owns the output bytes between
//                the 32-byte boundaries at or after its first letter and at or after its successor's first letter.  It
//                decodes from its true entry (discarding the < 32 letters before its first boundary, which its
//                predecessor writes), packs 32 letters into eight registers with static byte inserts and stores one
//                full 32-byte sector per lane with a 256-bit store (gap-free rewrite, every sector written once).
//
// Decoding uses a 12-bit first-level table (letter, length), a second-level table for codes of 13..20 bits (one
// 256-entry table per tree node at depth 12, in global memory, read through L1 by those codes only) and a tree walk
// behind it for longer codes (any depth); the count pass adds a 13-bit multi-letter table (bits consumed, letters
// completed) that skips several short codes per lookup.  Both passes come in two instances, chosen per launch by the
// tree's longest code: the common one has no tests for codes the tables do not cover.  Lookups and stream words come from shared memory (stream rows padded by one word per
// 32 so that threads walking their own subsequence in lock step hit 32 different banks).  The inner loops keep a
// 64-bit bit window in registers and address shared memory through 32-bit shared-space pointers.
//
// Algorithmic HBM bytes: C + N.  This implementation moves 2C + N (the stream is read by both passes) plus 4 bytes
// of per-subsequence metadata per 128 stream bytes.
#pragma once

#include "hb_common.cuh"

namespace hb {

#ifndef HB_DEC_THREADS
#define HB_DEC_THREADS 256
#endif
#ifndef HB_LEAD_LOOKBACK_BITS
#define HB_LEAD_LOOKBACK_BITS 384        // look-back of a CTA's first thread (nobody verifies it inside the CTA)
#endif
#ifndef HB_LOOKBACK_BITS
#define HB_LOOKBACK_BITS 192
#endif
constexpr int kDecThreads = HB_DEC_THREADS;
constexpr int kSubWords = 32;                                  // 1024-bit subsequence per thread
constexpr int kSubBits = kSubWords * 32;
constexpr int kChunkWords = kDecThreads * kSubWords;           // 8192 words = 32 KB per CTA
constexpr int kHaloWords = 32;                                 // before and after the chunk
constexpr int kWinWords = kHaloWords + kChunkWords + kHaloWords;
constexpr int kWinPhys = kWinWords + (kWinWords >> 5) + 4;     // padded: phys(i) = i + i/32 (+ slack for look-ahead)
constexpr uint32_t kWinBits = kWinWords * 32u;
constexpr int kLutBits = 12;
#ifndef HB_CNT_BITS
#define HB_CNT_BITS 13                                            // 12 and 14 measured slower
#endif
constexpr int kCntBits = HB_CNT_BITS;                           // index width of the multi-letter count table
constexpr int kCntBitsMax = 14;
constexpr uint32_t kEnd32 = 0xFFFFFFFFu;
constexpr uint64_t kEnd64 = ~0ull;
constexpr int kLeadLookbackBits = HB_LEAD_LOOKBACK_BITS;
constexpr int kLookbackBits = HB_LOOKBACK_BITS;                             // in-CTA look-back window W (a CTA's first thread: kLeadLookbackBits);
                                                               // measured resynchronisation distance: mean 16, p99 < 90 bits
constexpr int kScanGroup = 1024;                               // CTAs per offset-scan group
constexpr int kGroup = 32;                                     // letters per 256-bit store in the write pass

struct DecTables {                     // device resident, built on the host from the hb_tree
    uint16_t lut[1 << kLutBits];       // letter | len << 11, or (code longer than 12 bits) slot | 0x100 with len 0
    uint8_t  cnt[1 << kCntBitsMax];    // indexed by the next cnt_bits bits: (bits consumed << 4) | letters completed,
                                       // 0 if not even one code fits
    uint32_t nodes[HB_MAX_NODES];      // left | right << 16 ; leaf: left = 0xFFFF, right = letter
    uint32_t root;
    // second level (read from global memory / L1, only for codes of 13..20 bits): one 256-entry table per tree node
    // at depth 12 ("slot"), indexed by stream bits 12..19: letter | total len << 11, or slot | 0x100 if still longer
    uint16_t slot_node[256];           // slot -> node index (to continue a bit-serial walk for codes > 20 bits)
    uint16_t lut2[256 * 256];
};

struct DecParams {
    const uint32_t *words;             // stream as little-endian u32 words (byte k of the stream = byte k of memory)
    uint64_t n_words_readable;         // words that may be loaded
    uint64_t avail_bits;               // code words must end at or before this bit
    uint64_t own_begin, own_end;       // letters whose code word starts in [entry, own_end) are ours
    int64_t  entry_bit;                // >= 0: known first code-word start (>= own_begin); < 0: speculate
    uint64_t stream_bit0;              // stream bit index of buffer bit 0 (phase of the gcd alignment)
    uint32_t len_gcd, fixed_len;       // gcd of code lengths; fixed_len != 0 when all codes have that length
    uint32_t max_len;                  // longest code (bounds how far a thread may read past its subsequence)
    uint32_t cnt_bits;                 // index width of DecTables::cnt in use (= kCntBits)
    uint32_t spoil_speculation;        // test hook: CTA-leading threads skip their look-back (forces the repair path)
    uint32_t first_block, n_blocks;    // CTAs cover chunks first_block .. first_block + n_blocks - 1
    uint32_t *sub_info;                // per subsequence (relative to first_block): entry_rel << 16 | count
    uint64_t *blk_entry, *blk_exit;    // per CTA, absolute buffer bits (kEnd64 = none)
    uint32_t *blk_count;               // per CTA letters
};

// ---------------------------------------------------------------- shared-space access (32-bit addresses in registers)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void stg256(void *p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

struct DecShared {                     // shared-space byte addresses + the second-level table in global memory
    uint32_t win, lut, cnt, nodes;
    const uint16_t *lut2;
    const uint16_t *slot_node;
};

__device__ __forceinline__ uint32_t win_word_addr(uint32_t win, uint32_t i) { return win + ((i + (i >> 5)) << 2); }
__device__ __forceinline__ uint32_t win_bit(uint32_t win, uint32_t q) {
    return (lds32(win_word_addr(win, q >> 5)) >> (31 - (q & 31))) & 1u;
}

// ---------------------------------------------------------------- register bit window
// Two consecutive stream words and a bit offset: peek = one funnel shift, consume = one add, and a predicated
// one-word refill whenever the offset crosses 32.
struct BitReader {
    uint32_t w0, w1;        // stream words wi-2 and wi-1 (MSB first)
    uint32_t s, wi, q;      // bit offset inside w0 (< 32), next word to load, stream position (= 32*(wi-2) + s)
    __device__ __forceinline__ void init(uint32_t win, uint32_t q0) {
        q = q0; wi = q0 >> 5; s = q0 & 31;
        w0 = lds32(win_word_addr(win, wi));
        w1 = lds32(win_word_addr(win, wi + 1));
        wi += 2;
    }
    // byte offset of the first-level entry for the next kLutBits bits (mask + one LEA.HI with the table base)
    // (the AND is opaque to the optimiser, which would otherwise turn it back into shift + mask + add)
    __device__ __forceinline__ uint32_t peek_lut_off() const {
        uint32_t y;
        asm("and.b32 %0, %1, %2;" : "=r"(y) : "r"(__funnelshift_l(w1, w0, s)), "n"(~((1u << (32 - kLutBits)) - 1u)));
        return y >> (31 - kLutBits);
    }
    __device__ __forceinline__ void consume(uint32_t win, uint32_t l) {  // l < 32; leaves q stale (see users)
        s += l;
        if (s >= 32) {
            s -= 32;
            w0 = w1;
            w1 = lds32(win_word_addr(win, wi));
            wi++;
        }
    }
};

// Table entry (both levels), 16 bits: letter (bits 0-7) | long flag (bit 8) | code length (bits 11-15).  The length in
// the top bits makes "position += length" one shift-add (LEA.HI) with no mask; a long entry has length 0 and carries
// the second-level slot in the letter field.
constexpr uint32_t kLutLongFlag = 0x100u;
constexpr int kLutLenShift = 11;
__device__ __forceinline__ uint32_t lut_is_long(uint32_t e) { return e & kLutLongFlag; }
__device__ __forceinline__ uint32_t lut_len(uint32_t e) { return e >> kLutLenShift; }
__device__ __forceinline__ uint32_t lut_letter(uint32_t e) { return e & 0xFFu; }
__device__ __forceinline__ uint32_t lut_slot(uint32_t e) { return e & 0xFFu; }
// Second-level lookup for a first-level entry with the long flag: stream bits 12..19 after the reader's position.
// Returns an entry of the same layout (len 13..20 | letter << 8), or one with the long flag still set (> 20 bits).
__device__ __forceinline__ uint32_t lut_resolve(const DecShared &s, uint32_t e, const BitReader &rd) {
#ifdef HB_NO_LUT2                      // A/B switch: every code longer than 12 bits takes the bit-serial walk
    return e;
#endif
    const uint32_t x = __funnelshift_l(rd.w1, rd.w0, rd.s);
    return __ldg(s.lut2 + (lut_slot(e) << 8) + ((x >> (32 - kLutBits - 8)) & 0xFFu));
}

// Decode one code word starting at window bit q (table + tree walk, any length).  Returns its length, 0 if it would
// end after q_avail.
__device__ __noinline__ uint32_t dec_one_slow(DecShared s, uint32_t q, uint32_t q_avail, uint32_t &letter) {
    const uint32_t i = q >> 5;
    const uint32_t x = __funnelshift_l(lds32(win_word_addr(s.win, i + 1)), lds32(win_word_addr(s.win, i)), q & 31);
    const uint32_t e = lds16(s.lut + ((x >> (32 - kLutBits)) << 1));
    uint32_t len;
    if (!lut_is_long(e)) {
        len = lut_len(e);
        letter = lut_letter(e);
    } else {
        uint32_t node = __ldg(s.slot_node + lut_slot(e));
        len = kLutBits;
        for (;;) {
            const uint32_t nd = lds32(s.nodes + (node << 2));
            if ((nd & 0xFFFFu) == 0xFFFFu) { letter = nd >> 16; break; }
            if (q + len >= q_avail) return 0;
            node = win_bit(s.win, q + len) ? (nd >> 16) : (nd & 0xFFFFu);
            len++;
        }
    }
    return (q + len <= q_avail) ? len : 0;
}

// Advance from q over whole code words while q < q_stop; count them.  Returns the first code-word start >= q_stop,
// or kEnd32 when a code word does not fit below q_avail.  kLong = the tree has codes longer than the count table's
// index (else every table entry completes at least one letter and the zero tests disappear).
//
// The common-case loop keeps only (w0, w1, position, next word): the funnel shift takes the position modulo 32 by
// itself and a refill is due exactly when bit 5 of the position flips (a step is < 32 bits).  Every step adds
// bits << 4 | letters to ONE accumulator; the letter count falls out at the end as acc - 16 * (bits consumed).
template <bool kLong>
__device__ __forceinline__ uint32_t dec_run(DecShared s, uint32_t q, uint32_t q_stop, uint32_t q_avail, uint32_t &count) {
    constexpr int CB = kCntBits;
    if (q == kEnd32) { count = 0; return kEnd32; }
    if (q < q_stop && q_stop + 2 * CB <= q_avail) {
        // common case (everything but the very end of the stream): no code word can run past q_avail here
        const uint32_t q_begin = q;
        uint32_t wi = q >> 5;
        uint32_t w0 = lds32(win_word_addr(s.win, wi));
        uint32_t w1 = lds32(win_word_addr(s.win, wi + 1));
        wi += 2;
        uint32_t acc = 0;
        auto step = [&](uint32_t bits) {                                 // bits < 32
            const uint32_t qn = q + bits;
            if ((qn ^ q) & 32u) {
                w0 = w1;
                w1 = lds32(win_word_addr(s.win, wi));
                wi++;
            }
            q = qn;
        };
        auto peek = [&]() { return __funnelshift_l(w1, w0, q); };
        // one letter through the letter tables: the last few before q_stop, or a code longer than CB bits
        // (second-level table up to 20 bits, else the bit-serial walk).  false: it would end after q_avail.
        auto one_letter = [&]() -> bool {
            uint32_t y;
            asm("and.b32 %0, %1, %2;" : "=r"(y) : "r"(peek()), "n"(~((1u << (32 - kLutBits)) - 1u)));
            uint32_t e = lds16(s.lut + (y >> (31 - kLutBits)));
            if (lut_is_long(e))
                e = __ldg(s.lut2 + (lut_slot(e) << 8) + ((peek() >> (32 - kLutBits - 8)) & 0xFFu));
            uint32_t len = lut_len(e);
            if (lut_is_long(e)) {
                uint32_t letter;
                len = dec_one_slow(s, q, q_avail, letter);
                if (!len) return false;
                acc += (len << 4) + 1;
                q += len;                                                // any length: reload the two words
                wi = q >> 5;
                w0 = lds32(win_word_addr(s.win, wi));
                w1 = lds32(win_word_addr(s.win, wi + 1));
                wi += 2;
            } else {
                step(len);
                acc += (len << 4) + 1;
            }
            return true;
        };
        // two multi-letter steps per trip: one position check per trip, straight-line code between the lookups
        // (measured: 2 steps per trip = 21 % faster than 1).  A zero entry (kLong only) = the next code is longer
        // than CB bits: take that one letter and stay in this loop.
        if (q_stop >= 2 * CB) {
            const uint32_t q_lim2 = q_stop - 2 * CB;
            while (q <= q_lim2) {
                const uint32_t c1 = lds8(s.cnt + (peek() >> (32 - CB)));
                if (kLong && !c1) {
                    if (!one_letter()) { count = acc - ((q - q_begin) << 4); return kEnd32; }
                    continue;
                }
                step(c1 >> 4);
                const uint32_t c2 = lds8(s.cnt + (peek() >> (32 - CB)));
                if (kLong && !c2) {
                    acc += c1;
                    if (!one_letter()) { count = acc - ((q - q_begin) << 4); return kEnd32; }
                    continue;
                }
                step(c2 >> 4);
                acc += c1 + c2;
            }
        }
        for (;;) {
            const bool multi = q + CB <= q_stop;                         // a multi-letter step stays inside [q, q_stop)
            if (!multi && q >= q_stop) break;
            if (multi) {
                const uint32_t c = lds8(s.cnt + (peek() >> (32 - CB)));
                if (!kLong || c) { step(c >> 4); acc += c; continue; }
            }
            if (!one_letter()) { count = acc - ((q - q_begin) << 4); return kEnd32; }
        }
        count = acc - ((q - q_begin) << 4);
        return q;
    }
    uint32_t n = 0;
    while (q < q_stop) {                                                 // end of the stream: every step checked
        uint32_t letter;
        const uint32_t len = dec_one_slow(s, q, q_avail, letter);
        if (!len) { count = n; return kEnd32; }
        q += len;
        n++;
    }
    count = n;
    return q;
}

__device__ __forceinline__ void dec_load_tables(const DecTables *__restrict__ t, uint16_t *s_lut, uint8_t *s_cnt,
                                                uint32_t *s_nodes, int cnt_bits = kLutBits) {
    const uint32_t *src_lut = reinterpret_cast<const uint32_t *>(t->lut);
    uint32_t *dst_lut = reinterpret_cast<uint32_t *>(s_lut);
    for (int i = threadIdx.x; i < (1 << kLutBits) / 2; i += blockDim.x) dst_lut[i] = src_lut[i];
    if (s_cnt) {
        const uint32_t *src_cnt = reinterpret_cast<const uint32_t *>(t->cnt);
        uint32_t *dst_cnt = reinterpret_cast<uint32_t *>(s_cnt);
        for (int i = threadIdx.x; i < (1 << cnt_bits) / 4; i += blockDim.x) dst_cnt[i] = src_cnt[i];
    }
    for (int i = threadIdx.x; i < HB_MAX_NODES; i += blockDim.x) s_nodes[i] = t->nodes[i];
}

// Stage window words [chunk*8192 - 32, chunk*8192 + 8192 + 32) of the stream into padded shared memory, MSB-first.
// All global loads of a thread are issued before the first dependent use (9 x 128-bit in flight per thread).
__device__ __forceinline__ void dec_load_window(const DecParams &p, uint32_t chunk, uint32_t *win) {
    const long long w_begin = static_cast<long long>(chunk) * kChunkWords - kHaloWords;
    constexpr int kVecs = kWinWords / 4;                                   // 2064 uint4
    constexpr int kPerThread = (kVecs + kDecThreads - 1) / kDecThreads;    // 9
    if ((reinterpret_cast<uintptr_t>(p.words) & 15) == 0) {
        uint4 v[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; k++) {
            const int i4 = threadIdx.x + k * kDecThreads;
            const long long gw = w_begin + 4ll * i4;
            v[k] = make_uint4(0, 0, 0, 0);
            if (i4 < kVecs && gw >= 0 && static_cast<uint64_t>(gw) < p.n_words_readable) {
                if (static_cast<uint64_t>(gw) + 4 <= p.n_words_readable) {
                    v[k] = ld_stream_u4(reinterpret_cast<const uint4 *>(p.words + gw));
                } else {                       // last vector of the stream: never read past the last stream word
                    const uint64_t left = p.n_words_readable - static_cast<uint64_t>(gw);
                    v[k].x = ld_stream_u32(p.words + gw);
                    if (left > 1) v[k].y = ld_stream_u32(p.words + gw + 1);
                    if (left > 2) v[k].z = ld_stream_u32(p.words + gw + 2);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kPerThread; k++) {
            const int i4 = threadIdx.x + k * kDecThreads;
            if (i4 < kVecs) {
                const int i = 4 * i4;
                uint32_t *dst = win + i + (i >> 5);                        // the 4 words share a row: contiguous
                dst[0] = bswap32(v[k].x); dst[1] = bswap32(v[k].y); dst[2] = bswap32(v[k].z); dst[3] = bswap32(v[k].w);
            }
        }
    } else {
        for (int base = 0; base < kWinWords; base += 8 * kDecThreads) {
            uint32_t v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int i = base + threadIdx.x + k * kDecThreads;
                const long long gw = w_begin + i;
                v[k] = 0;
                if (i < kWinWords && gw >= 0 && static_cast<uint64_t>(gw) < p.n_words_readable) v[k] = ld_stream_u32(p.words + gw);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int i = base + threadIdx.x + k * kDecThreads;
                if (i < kWinWords) win[i + (i >> 5)] = bswap32(v[k]);
            }
        }
    }
    if (threadIdx.x < 3) win[kWinWords + (kWinWords >> 5) + threadIdx.x] = 0;   // look-ahead slack past the window
}

// The count pass for one chunk.
template <bool kLong>
__device__ void dec_count_block(const DecParams &p, uint32_t blk, uint64_t entry_override, bool use_override,
                                uint32_t *win, DecShared s, uint32_t *s_exit, uint32_t *s_red) {
    const int t = threadIdx.x;
    const uint32_t chunk = p.first_block + blk;
    dec_load_window(p, chunk, win);
    __syncthreads();

    // window bit q  <->  buffer bit  win_bit0 + q   (win_bit0 may be negative for chunk 0)
    const long long win_bit0 = (static_cast<long long>(chunk) * kChunkWords - kHaloWords) * 32;
    auto to_win = [&](uint64_t abs_bit) -> uint32_t {       // clamp an absolute bit into [0, kWinBits]
        const long long q = static_cast<long long>(abs_bit) - win_bit0;
        return q < 0 ? 0u : (q > static_cast<long long>(kWinBits) ? kWinBits : static_cast<uint32_t>(q));
    };
    const uint32_t q_avail = to_win(p.avail_bits);
    const uint32_t q_own_begin = to_win(p.own_begin);
    const uint32_t q_own_end = to_win(p.own_end);
    const uint32_t q_buf0 = to_win(0);                        // first real bit of the buffer

    const uint32_t q_sub = (kHaloWords + t * kSubWords) * 32u;             // my subsequence [q_sub, q_sub + 1024)
    const uint32_t q_lo = max(q_sub, q_own_begin);
    const uint32_t q_hi = min(q_sub + kSubBits, q_own_end);
    const bool active = q_lo < q_sub + kSubBits && (q_sub < q_own_end);   // subsequence intersects the owned range
    const bool is_first = active && (q_own_begin >= q_sub);               // contains own_begin: no predecessor
    const bool has_pred = active && !is_first && t > 0;

    // ---- phase A: entry candidate
    uint32_t entry = kEnd32;
    if (active) {
        if (is_first && p.entry_bit >= 0) {
            entry = to_win(static_cast<uint64_t>(p.entry_bit));
        } else if (!has_pred && use_override && !is_first) {
            entry = entry_override == kEnd64 ? kEnd32 : to_win(entry_override);
        } else {
            // A CTA's first thread is checked only across CTAs (dec_verify_kernel; a miss costs a serial repair), so
            // it looks back further -- but not the whole halo: its warp, and with it the CTA's barrier, waits for it.
            // The miss rate falls ~27x per 64 bits (measured 2.7e-3 / 1e-4 / < 5e-5 at 128 / 192 / 256 bits).
            uint32_t window = is_first ? static_cast<uint32_t>(kHaloWords * 32)
                                       : (t == 0 ? static_cast<uint32_t>(kLeadLookbackBits) : static_cast<uint32_t>(kLookbackBits));
            if (p.fixed_len) window = 0;
            if (p.spoil_speculation && t == 0 && !is_first) window = 0;   // deliberately bad guess (tests only)
            uint32_t q0 = q_lo > window ? q_lo - window : 0;
            if (q0 < q_buf0) q0 = q_buf0;
            if (p.len_gcd > 1) {           // align to the phase of the stream: (stream_bit0 + abs) % gcd == 0
                const unsigned long long abs0 = static_cast<unsigned long long>(win_bit0 + q0) + p.stream_bit0;
                const uint32_t rem = static_cast<uint32_t>(abs0 % p.len_gcd);
                if (rem) q0 += p.len_gcd - rem;
            }
            uint32_t dummy;
            entry = q0 >= q_lo ? q0 : dec_run<kLong>(s, q0, q_lo, q_avail, dummy);
        }
    }

    // ---- phase B + in-CTA verification to a fixed point
    uint32_t count = 0, exitq = kEnd32;
    bool redo = active;
    for (int round = 0;; round++) {
        if (round > kDecThreads + 1) asm volatile("trap;");   // cannot happen: thread t is final after t rounds
        if (redo) exitq = dec_run<kLong>(s, entry, q_hi, q_avail, count);
        s_exit[t] = exitq;
        __syncthreads();
        redo = false;
        if (has_pred) {
            const uint32_t want = s_exit[t - 1];
            if (want != entry) { entry = want; redo = true; }
        }
        if (!__syncthreads_or(redo)) break;
    }

    // ---- results
    const uint32_t sub = blk * kDecThreads + t;
    uint32_t entry_rel = 0xFFFFu;
    if (active && entry != kEnd32) entry_rel = entry - q_sub;              // < max code length + alignment slack
    p.sub_info[sub] = (entry_rel << 16) | (active ? count : 0u);

    // CTA totals: letters, entry of the first active thread, exit of the last active thread
    uint32_t c = active ? count : 0u;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, sft);
    if ((t & 31) == 0) s_red[t >> 5] = c;
    __syncthreads();
    if (t == 0) {
        uint32_t total = 0;
        for (int k = 0; k < kDecThreads / 32; k++) total += s_red[k];
        p.blk_count[blk] = total;
    }
    if (active && !has_pred)                                               // exactly one such thread per CTA
        p.blk_entry[blk] = entry == kEnd32 ? kEnd64 : static_cast<uint64_t>(win_bit0 + entry);
    if (active && (t == kDecThreads - 1 || q_sub + kSubBits >= q_own_end))   // the last active thread of the CTA
        p.blk_exit[blk] = exitq == kEnd32 ? kEnd64 : static_cast<uint64_t>(win_bit0 + exitq);
    __syncthreads();
}

extern __shared__ __align__(16) uint8_t dec_smem[];

constexpr size_t dec_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
constexpr size_t kDecOffLut = dec_align16(kWinPhys * 4);
constexpr size_t kDecOffNodes = kDecOffLut + (1 << kLutBits) * 2;
constexpr size_t kDecOffRed = kDecOffNodes + dec_align16(HB_MAX_NODES * 4);
constexpr size_t kDecOffExit = kDecOffRed + 64;
constexpr size_t kDecOffCnt = kDecOffExit + kDecThreads * 4;
constexpr size_t dec_smem_count(int cnt_bits) { return kDecOffCnt + (static_cast<size_t>(1) << cnt_bits); }
constexpr size_t kDecSmemWrite = kDecOffExit;                           // window + lut + nodes + red

struct DecCarve {
    uint32_t *win; uint16_t *lut; uint8_t *cnt; uint32_t *nodes; uint32_t *exit; uint32_t *red;
    DecShared sh;
};
__device__ __forceinline__ DecCarve dec_carve(uint8_t *base, const DecTables *tables) {
    DecCarve c;
    c.sh.lut2 = tables->lut2;
    c.sh.slot_node = tables->slot_node;
    c.win = reinterpret_cast<uint32_t *>(base);
    c.lut = reinterpret_cast<uint16_t *>(base + kDecOffLut);
    c.nodes = reinterpret_cast<uint32_t *>(base + kDecOffNodes);
    c.red = reinterpret_cast<uint32_t *>(base + kDecOffRed);
    c.cnt = base + kDecOffCnt;
    c.exit = reinterpret_cast<uint32_t *>(base + kDecOffExit);
    // one opaque base register: otherwise the compiler rematerialises the shared-window base (S2UR + ULEA + IMAD)
    // inside the decode loops instead of keeping it live
    uint32_t b = smem_addr(base);
    asm volatile("mov.u32 %0, %0;" : "+r"(b));
    c.sh.win = b;
    c.sh.lut = b + static_cast<uint32_t>(kDecOffLut);
    c.sh.cnt = b + static_cast<uint32_t>(kDecOffCnt);
    c.sh.nodes = b + static_cast<uint32_t>(kDecOffNodes);
    return c;
}

template <bool kLong>
__global__ void __launch_bounds__(kDecThreads)
dec_count_kernel(DecParams p, const DecTables *__restrict__ tables) {
    DecCarve c = dec_carve(dec_smem, tables);
    dec_load_tables(tables, c.lut, c.cnt, c.nodes, kCntBits);
    for (uint32_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x)
        dec_count_block<kLong>(p, blk, 0, false, c.win, c.sh, c.exit, c.red);
}

// dirty[j] = 1 when CTA j's entry is not its predecessor's exit.  n_dirty accumulates.
__global__ void dec_verify_kernel(DecParams p, uint32_t *dirty, uint32_t *n_dirty) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.n_blocks) return;
    uint32_t d = 0;
    if (j > 0 && p.blk_entry[j] != p.blk_exit[j - 1]) d = 1;
    dirty[j] = d;
    if (d) atomicAdd(n_dirty, 1u);
}

// Serial repair of mismatching CTAs (single CTA).  Each repaired chunk may change its exit and dirty its successor.
__global__ void __launch_bounds__(kDecThreads)
dec_fix_kernel(DecParams p, const DecTables *__restrict__ tables, uint32_t *dirty) {
    DecCarve c = dec_carve(dec_smem, tables);
    __shared__ uint32_t s_next;
    dec_load_tables(tables, c.lut, c.cnt, c.nodes, kCntBits);
    uint32_t cur = 1;
    for (;;) {
        // find the next dirty chunk at or after cur
        __syncthreads();
        if (threadIdx.x == 0) s_next = 0xFFFFFFFFu;
        __syncthreads();
        const volatile uint32_t *vdirty = dirty;
        for (uint32_t base = cur; base < p.n_blocks; base += kDecThreads) {
            const uint32_t j = base + threadIdx.x;
            if (j < p.n_blocks && vdirty[j]) atomicMin(&s_next, j);
            __syncthreads();
            const uint32_t found = s_next;
            __syncthreads();
            if (found != 0xFFFFFFFFu) break;
        }
        __syncthreads();
        const uint32_t j = s_next;
        if (j == 0xFFFFFFFFu) break;
        const volatile uint64_t *vexit = p.blk_exit;
        const uint64_t old_exit = vexit[j];
        const uint64_t entry = vexit[j - 1];
        __syncthreads();
        dec_count_block<true>(p, j, entry, true, c.win, c.sh, c.exit, c.red);
        if (threadIdx.x == 0) {
            dirty[j] = 0;
            if (j + 1 < p.n_blocks && vexit[j] != old_exit) dirty[j + 1] = 1;
            __threadfence();
        }
        cur = j + 1;
    }
}

constexpr int kScanThreads = kScanGroup / 4;                  // each scan thread owns 4 CTA counts
// ---- output offsets: group-local exclusive scan of CTA counts + scan of group totals
__global__ void __launch_bounds__(kScanThreads)
dec_scan_groups_kernel(const uint32_t *__restrict__ blk_count, uint32_t n_blocks, uint32_t *__restrict__ blk_local,
                       uint64_t *__restrict__ group_total) {
    __shared__ uint32_t s_w[kScanThreads / 32];
    const uint32_t base = blockIdx.x * kScanGroup + threadIdx.x * 4;
    uint32_t v[4];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k] = (base + k < n_blocks) ? blk_count[base + k] : 0u; sum += v[k]; }
    // per-group letters < 1024 * 256 * 1024 = 2^28: u32 is enough
    const uint32_t incl = warp_incl_scan(sum);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int k = 0; k < kScanThreads / 32; k++) { if (k < (threadIdx.x >> 5)) before += s_w[k]; total += s_w[k]; }
    uint32_t run = before + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n_blocks) blk_local[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) group_total[blockIdx.x] = total;
}

// single CTA: exclusive scan of group totals in place -> group offsets; writes the grand total
__global__ void __launch_bounds__(kScanThreads)
dec_scan_totals_kernel(uint64_t *group_total, uint32_t n_groups, uint64_t *grand_total) {
    __shared__ unsigned long long s_w[kScanThreads / 32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_groups; base += kScanThreads) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_groups ? group_total[i] : 0ull;
        const unsigned long long incl = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long before = 0, total = 0;
        for (int k = 0; k < kScanThreads / 32; k++) { if (k < (threadIdx.x >> 5)) before += s_w[k]; total += s_w[k]; }
        const unsigned long long carry = s_carry;
        if (i < n_groups) group_total[i] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = s_carry;
}

// ---------------------------------------------------------------- write pass
// Letter source that works anywhere: the staged window while the position is safely inside it, the stream in global
// memory (bit-serial tree walk) once a thread has to read past its window (only with very long codes).
// bit-serial walk from the root over global memory (comp.rs:496-509 literally); returns letter | bits consumed << 8
__device__ __noinline__ uint32_t dec_one_global(const uint32_t *__restrict__ words, const uint32_t *__restrict__ nodes_g,
                                                uint32_t root, unsigned long long pos) {
    uint32_t nd = nodes_g[root];
    if ((nd & 0xFFFFu) == 0xFFFFu) return (nd >> 16) | (1u << 8);            // lone root: one letter per bit
    uint32_t used = 0;
    while ((nd & 0xFFFFu) != 0xFFFFu) {
        const unsigned long long b = pos + used;
        const uint32_t bit = (bswap32(words[b >> 5]) >> (31 - (b & 31))) & 1u;
        nd = nodes_g[bit ? (nd >> 16) : (nd & 0xFFFFu)];
        used++;
    }
    return (nd >> 16) | (used << 8);
}

struct LetterSource {
    DecShared s;
    const uint32_t *words;
    const uint32_t *nodes_g;
    uint32_t root;
    long long win_bit0;
    BitReader rd;
    uint32_t q_safe;        // window positions below this never read outside the staged words
    bool global_mode;
    unsigned long long gpos;
};

// one letter, any code length, any position (slow paths only: ragged ends, long codes, window overrun)
__device__ __noinline__ uint32_t slow_next(LetterSource &src) {
    if (src.global_mode || src.rd.q >= src.q_safe) {
        if (!src.global_mode) { src.global_mode = true; src.gpos = static_cast<unsigned long long>(src.win_bit0 + src.rd.q); }
        const uint32_t r = dec_one_global(src.words, src.nodes_g, src.root, src.gpos);
        src.gpos += r >> 8;
        return r & 0xFFu;
    }
    uint32_t letter = 0;
    const uint32_t len = dec_one_slow(src.s, src.rd.q, kWinBits, letter);
    src.rd.init(src.s.win, src.rd.q + len);
    return letter;
}

// 32 letters from the register bit window, one table lookup each, static byte inserts, one 256-bit store.
// kResolve = false: branch-free body for trees without codes longer than the first-level table (the launch-uniform
// common case).  kResolve = true: a first-level miss is resolved per letter in the second-level table (13..20 bits).
// (Measured alternatives for kResolve: testing once per four letters and redoing the quad -- a long entry has length 0
// and repeats itself, so one test finds it -- is 3x SLOWER at 1.5 % long letters: some lane of the warp hits one in
// 86 % of the quads; an optimistic whole-group attempt is worse still.)
// Returns false with nothing stored and `reader` untouched when a code is longer than the tables cover; the caller
// then redoes the group letter by letter.
template <bool kResolve>
__device__ __forceinline__ bool dec_fast_group(const DecShared &sh, BitReader &reader, uint8_t *dst) {
    BitReader rd = reader;
    uint32_t v[8];
    uint32_t escape = 0;
#pragma unroll
    for (int q = 0; q < kGroup / 4; q++) {
        uint32_t e[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            e[k] = lds16(sh.lut + rd.peek_lut_off());
            if (kResolve) {
                if (lut_is_long(e[k])) e[k] = lut_resolve(sh, e[k], rd);  // rare, divergent
                escape |= e[k];                                           // (no long entries exist when !kResolve)
            }
            // consume (the stream position rd.q is recomputed after the group)
            rd.s += lut_len(e[k]);
            if (rd.s >= 32) {
                rd.s -= 32;
                rd.w0 = rd.w1;
                rd.w1 = lds32(win_word_addr(sh.win, rd.wi));
                rd.wi++;
            }
        }
        // the letter sits in byte 0 of an entry: three PRMTs gather four letters into one output word
        uint32_t w = __byte_perm(e[0], e[1], 0x4440);       // bytes 0, 1 (bytes 2, 3 are replaced next)
        w = __byte_perm(w, e[2], 0x3410);
        v[q] = __byte_perm(w, e[3], 0x4210);
    }
    if (lut_is_long(escape)) return false;
    rd.q = ((rd.wi - 2) << 5) + rd.s;
    stg256(dst, v);
    reader = rd;
    return true;
}

__global__ void __launch_bounds__(kDecThreads)
dec_write_kernel(DecParams p, const DecTables *__restrict__ tables, const uint32_t *__restrict__ blk_local,
                 const uint64_t *__restrict__ group_off, uint8_t *__restrict__ out, uint64_t total_letters) {
    DecCarve c = dec_carve(dec_smem, tables);
    __shared__ uint32_t s_w[kDecThreads / 32];
    dec_load_tables(tables, c.lut, nullptr, c.nodes);
    const int t = threadIdx.x;
    const uintptr_t out_addr = reinterpret_cast<uintptr_t>(out);
    // between two safety checks a thread reads at most one group of letters (+ refill look-ahead)
    const uint32_t reach = (kGroup + 1) * min(p.max_len, 255u) + 96;
    const uint32_t q_safe_group = kWinBits > reach ? kWinBits - reach : 0;
    const bool has_long = p.max_len > kLutBits;                  // CTA-uniform: codes beyond the first-level table exist

    for (uint32_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
        __syncthreads();
        const uint32_t chunk = p.first_block + blk;
        dec_load_window(p, chunk, c.win);

        const uint32_t info = p.sub_info[blk * kDecThreads + t];
        const uint32_t my_count = info & 0xFFFFu;
        const uint32_t entry_rel = info >> 16;

        // CTA-wide exclusive scan of letter counts
        const uint32_t incl = warp_incl_scan(my_count);
        if ((t & 31) == 31) s_w[t >> 5] = incl;
        __syncthreads();                                     // also: window staged
        uint32_t before = 0;
        for (int k = 0; k < kDecThreads / 32; k++) if (k < (t >> 5)) before += s_w[k];
        const uint64_t first = group_off[blk / kScanGroup] + blk_local[blk] + before + incl - my_count;   // my first letter
        if (my_count == 0) continue;

        // output range I own: [lo, hi) = 32-byte boundaries at/after my first letter and at/after my successor's
        const uint64_t next_first = first + my_count;
        uint64_t lo = first + ((0 - (out_addr + first)) & 31);
        uint64_t hi = next_first + ((0 - (out_addr + next_first)) & 31);
        if (first == 0) lo = 0;                              // nobody precedes the first letter: its ragged head is mine
        if (hi > total_letters) hi = total_letters;          // ragged tail of the whole output
        if (lo >= hi) continue;

        LetterSource src;
        src.s = c.sh;
        src.words = p.words;
        src.nodes_g = tables->nodes;
        src.root = tables->root;
        src.win_bit0 = (static_cast<long long>(chunk) * kChunkWords - kHaloWords) * 32;
        src.q_safe = kWinBits - (min(p.max_len, 255u) + 96);
        src.global_mode = false;
        src.gpos = 0;
        src.rd.init(c.sh.win, (kHaloWords + t * kSubWords) * 32u + entry_rel);

        uint64_t pos = first;
        if (pos < lo && src.rd.q < q_safe_group) {                            // letters my predecessor writes (< 32)
            // the warp runs this loop for its slowest lane (~30 trips): 32-bit trip counter, no position tracking,
            // and no long-code test at all when the tree has no code beyond the first-level table
            BitReader rd = src.rd;
            uint32_t k = static_cast<uint32_t>(lo - pos);
            if (!has_long) {
#pragma unroll 1
                for (; k; k--) rd.consume(c.sh.win, lut_len(lds16(c.sh.lut + rd.peek_lut_off())));
            } else {
#pragma unroll 1
                for (; k; k--) {
                    uint32_t e = lds16(c.sh.lut + rd.peek_lut_off());
                    if (lut_is_long(e)) {
                        e = lut_resolve(c.sh, e, rd);
                        if (lut_is_long(e)) break;                            // > 20 bits: letter by letter below
                    }
                    rd.consume(c.sh.win, lut_len(e));
                }
            }
            rd.q = ((rd.wi - 2) << 5) + rd.s;
            pos = lo - k;
            src.rd = rd;
        }
        for (; pos < lo; pos++) (void)slow_next(src);
        for (; pos < hi && ((out_addr + pos) & 31); pos++)                    // ragged head (first letter only)
            out[pos] = static_cast<uint8_t>(slow_next(src));
        while (pos + kGroup <= hi) {
            bool done = false;
            if (!src.global_mode && src.rd.q < q_safe_group)
                done = has_long ? dec_fast_group<true>(c.sh, src.rd, out + pos) : dec_fast_group<false>(c.sh, src.rd, out + pos);
            if (!done)
                for (int j = 0; j < kGroup; j++) out[pos + j] = static_cast<uint8_t>(slow_next(src));
            pos += kGroup;
        }
        for (; pos < hi; pos++) out[pos] = static_cast<uint8_t>(slow_next(src));   // ragged tail (end of output)
    }
}

}  // namespace hb
