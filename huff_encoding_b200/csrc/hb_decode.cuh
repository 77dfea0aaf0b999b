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
//                with the corrected entry by dec_fix_kernel (rare: needs a code that fails to resynchronise
//                within a whole 1024-bit halo; worst case = strictly sequential propagation, still exact).
//   offsets      two small scan kernels turn per-CTA letter counts into 64-bit output offsets.
//   write pass   (dec_write_kernel)  every thread decodes exactly its count letters from its true entry into a
//                shared-memory window; windows are copied out with coalesced 128-bit stores (gap-free rewrite).
//
// Decoding uses a 12-bit first-level table (letter, length) with a tree walk behind it for longer codes (any depth),
// and for the count pass a 12-bit multi-letter table (bits consumed, letters completed) that skips several
// short codes per lookup.  Lookups and stream words come from shared memory (stream rows padded by one word per
// 32 so that threads walking their own subsequence in lock step hit 32 different banks).
//
// Algorithmic HBM bytes: C + N.  This implementation moves 2C + N (the stream is read by both passes) plus 4 bytes
// of per-subsequence metadata per 128 stream bytes.
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kDecThreads = 256;
constexpr int kSubWords = 32;                                  // 1024-bit subsequence per thread
constexpr int kSubBits = kSubWords * 32;
constexpr int kChunkWords = kDecThreads * kSubWords;           // 8192 words = 32 KB per CTA
constexpr int kHaloWords = 32;                                 // before and after the chunk
constexpr int kWinWords = kHaloWords + kChunkWords + kHaloWords;
constexpr int kWinPhys = kWinWords + (kWinWords >> 5) + 1;     // padded: phys(i) = i + i/32
constexpr uint32_t kWinBits = kWinWords * 32u;
constexpr int kLutBits = 12;
constexpr uint32_t kEnd32 = 0xFFFFFFFFu;
constexpr uint64_t kEnd64 = ~0ull;
constexpr int kLookbackBits = 512;                             // in-CTA look-back window W (thread 0 uses the full halo)
constexpr int kOutWindow = 49152;                              // letters staged per copy-out round in the write pass
constexpr int kScanGroup = 1024;                               // CTAs per offset-scan group

struct DecTables {                     // device resident, built on the host from the hb_tree
    uint16_t lut[1 << kLutBits];       // bit15 = 0: letter | len << 8 ; bit15 = 1: node index to continue from
    uint8_t  cnt[1 << kLutBits];       // (bits consumed << 4) | letters completed, 0 if the first code is longer than 12
    uint32_t nodes[HB_MAX_NODES];      // left | right << 16 ; leaf: left = 0xFFFF, right = letter
    uint32_t root_is_leaf;
    uint32_t lut2[1 << kLutBits];      // write pass: letter0 | letter1 << 8 | len0 << 16 | len(0+1) << 24 (0 = no 2nd);
                                       // len0 == 0: first code longer than 12 bits (slow path)
};

struct DecParams {
    const uint32_t *words;             // stream as little-endian u32 words (byte k of the stream = byte k of memory)
    uint64_t n_words_readable;         // words that may be loaded
    uint64_t avail_bits;               // code words must end at or before this bit
    uint64_t own_begin, own_end;       // letters whose code word starts in [entry, own_end) are ours
    int64_t  entry_bit;                // >= 0: known first code-word start (>= own_begin); < 0: speculate
    uint64_t stream_bit0;              // stream bit index of buffer bit 0 (phase of the gcd alignment)
    uint32_t len_gcd, fixed_len;       // gcd of code lengths; fixed_len != 0 when all codes have that length
    uint32_t first_block, n_blocks;    // CTAs cover chunks first_block .. first_block + n_blocks - 1
    uint32_t *sub_info;                // per subsequence (relative to first_block): entry_rel << 16 | count
    uint64_t *blk_entry, *blk_exit;    // per CTA, absolute buffer bits (kEnd64 = none)
    uint32_t *blk_count;               // per CTA letters
};

__device__ __forceinline__ uint32_t win_phys(uint32_t i) { return i + (i >> 5); }

__device__ __forceinline__ uint32_t win_peek32(const uint32_t *win, uint32_t q) {
    const uint32_t i = q >> 5;
    return __funnelshift_l(win[win_phys(i + 1)], win[win_phys(i)], q & 31);
}
__device__ __forceinline__ uint32_t win_bit(const uint32_t *win, uint32_t q) {
    return (win[win_phys(q >> 5)] >> (31 - (q & 31))) & 1u;
}

// Decode one code word starting at window bit q.  Returns its length (0 if it would end after q_avail).
__device__ __forceinline__ uint32_t dec_one(const uint32_t *win, const uint16_t *lut, const uint32_t *nodes,
                                            uint32_t q, uint32_t q_avail, uint32_t &letter) {
    const uint32_t x = win_peek32(win, q);
    const uint32_t e = lut[x >> (32 - kLutBits)];
    uint32_t len;
    if (!(e & 0x8000u)) {
        len = (e >> 8) & 0xFu;
        letter = e & 0xFFu;
    } else {
        uint32_t node = e & 0x3FFu;
        len = kLutBits;
        for (;;) {
            const uint32_t nd = nodes[node];
            if ((nd & 0xFFFFu) == 0xFFFFu) { letter = nd >> 16; break; }
            if (q + len >= q_avail) return 0;
            node = win_bit(win, q + len) ? (nd >> 16) : (nd & 0xFFFFu);
            len++;
        }
    }
    return (q + len <= q_avail) ? len : 0;
}

// Register bit window over the staged stream: buf holds the next `have` bits, MSB first.
struct BitReader {
    const uint32_t *win;
    unsigned long long buf;
    uint32_t have, wi, q;
    __device__ __forceinline__ void init(const uint32_t *w, uint32_t q0) {
        win = w; q = q0; wi = q0 >> 5;
        const uint32_t s = q0 & 31;
        const unsigned long long two = (static_cast<unsigned long long>(win[win_phys(wi)]) << 32) | win[win_phys(wi + 1)];
        buf = two << s;
        have = 64 - s;
        wi += 2;
    }
    __device__ __forceinline__ void refill() {          // afterwards have >= 33
        if (have <= 32) {
            buf |= static_cast<unsigned long long>(win[win_phys(wi)]) << (32 - have);
            have += 32;
            wi++;
        }
    }
    __device__ __forceinline__ uint32_t peek() const { return static_cast<uint32_t>(buf >> (64 - kLutBits)); }
    __device__ __forceinline__ void skip(uint32_t l) { buf <<= l; have -= l; q += l; }
};

// Advance from q over whole code words while q < q_stop; count them.  Returns the first code-word start >= q_stop,
// or kEnd32 when a code word does not fit below q_avail.
__device__ __forceinline__ uint32_t dec_run(const uint32_t *win, const uint16_t *lut, const uint8_t *cnt_lut,
                                            const uint32_t *nodes, uint32_t q, uint32_t q_stop, uint32_t q_avail,
                                            uint32_t &count) {
    uint32_t n = 0;
    if (q == kEnd32) { count = 0; return kEnd32; }
    const uint32_t fast_stop = min(q_stop, q_avail);
    if (q + kLutBits <= fast_stop) {
        BitReader rd;
        rd.init(win, q);
        while (rd.q + kLutBits <= fast_stop) {
            rd.refill();
            const uint32_t c = cnt_lut[rd.peek()];
            if (c) {
                rd.skip(c >> 4);
                n += c & 15u;
            } else {                                        // first code longer than 12 bits: tree walk, then re-prime
                uint32_t letter;
                const uint32_t len = dec_one(win, lut, nodes, rd.q, q_avail, letter);
                if (!len) { count = n; return kEnd32; }
                n++;
                rd.init(win, rd.q + len);
            }
        }
        q = rd.q;
    }
    while (q < q_stop) {
        uint32_t letter;
        const uint32_t len = dec_one(win, lut, nodes, q, q_avail, letter);
        if (!len) { count = n; return kEnd32; }
        q += len;
        n++;
    }
    count = n;
    return q;
}

__device__ __forceinline__ void dec_load_tables(const DecTables *__restrict__ t, uint16_t *s_lut, uint8_t *s_cnt,
                                                uint32_t *s_nodes, uint32_t *s_lut2 = nullptr) {
    if (s_lut2)
        for (int i = threadIdx.x; i < (1 << kLutBits); i += blockDim.x) s_lut2[i] = t->lut2[i];
    const uint32_t *src_lut = reinterpret_cast<const uint32_t *>(t->lut);
    uint32_t *dst_lut = reinterpret_cast<uint32_t *>(s_lut);
    for (int i = threadIdx.x; i < (1 << kLutBits) / 2; i += blockDim.x) dst_lut[i] = src_lut[i];
    if (s_cnt) {
        const uint32_t *src_cnt = reinterpret_cast<const uint32_t *>(t->cnt);
        uint32_t *dst_cnt = reinterpret_cast<uint32_t *>(s_cnt);
        for (int i = threadIdx.x; i < (1 << kLutBits) / 4; i += blockDim.x) dst_cnt[i] = src_cnt[i];
    }
    for (int i = threadIdx.x; i < HB_MAX_NODES; i += blockDim.x) s_nodes[i] = t->nodes[i];
}

// Stage window words [chunk*8192 - 32, chunk*8192 + 8192 + 32) of the stream into padded shared memory, MSB-first.
__device__ __forceinline__ void dec_load_window(const DecParams &p, uint32_t chunk, uint32_t *win) {
    const long long w_begin = static_cast<long long>(chunk) * kChunkWords - kHaloWords;
    for (int i = threadIdx.x; i < kWinWords; i += blockDim.x) {
        const long long gw = w_begin + i;
        uint32_t v = 0;
        if (gw >= 0 && static_cast<uint64_t>(gw) < p.n_words_readable) v = bswap32(ld_stream_u32(p.words + gw));
        win[win_phys(i)] = v;
    }
    if (threadIdx.x == 0) win[win_phys(kWinWords)] = 0;      // win_peek32 may touch one word past the window
}

// The count pass for one chunk.  entry_override: kEnd64-1 => none (speculate / use p.entry_bit).
__device__ void dec_count_block(const DecParams &p, uint32_t blk, uint64_t entry_override, bool use_override,
                                uint32_t *win, const uint16_t *s_lut, const uint8_t *s_cnt, const uint32_t *s_nodes,
                                uint32_t *s_exit, uint32_t *s_red) {
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
    const bool is_first = active && (q_own_begin >= q_sub) ;              // contains own_begin: no predecessor
    const bool has_pred = active && !is_first && t > 0;

    // ---- phase A: entry candidate
    uint32_t entry = kEnd32;
    if (active) {
        if (is_first && p.entry_bit >= 0) {
            entry = to_win(static_cast<uint64_t>(p.entry_bit));
        } else if (!has_pred && use_override && !is_first) {
            entry = entry_override == kEnd64 ? kEnd32 : to_win(entry_override);
        } else {
            uint32_t window = (t == 0 || is_first) ? static_cast<uint32_t>(kHaloWords * 32) : static_cast<uint32_t>(kLookbackBits);
            if (p.fixed_len) window = 0;
            uint32_t q0 = q_lo > window ? q_lo - window : 0;
            if (q0 < q_buf0) q0 = q_buf0;
            if (p.len_gcd > 1) {           // align to the phase of the stream: (stream_bit0 + abs) % gcd == 0
                const unsigned long long abs0 = static_cast<unsigned long long>(win_bit0 + q0) + p.stream_bit0;
                const uint32_t rem = static_cast<uint32_t>(abs0 % p.len_gcd);
                if (rem) q0 += p.len_gcd - rem;
            }
            uint32_t dummy;
            entry = q0 >= q_lo ? q0 : dec_run(win, s_lut, s_cnt, s_nodes, q0, q_lo, q_avail, dummy);
        }
    }

    // ---- phase B + in-CTA verification to a fixed point
    uint32_t count = 0, exitq = kEnd32;
    bool redo = active;
    for (int round = 0;; round++) {
        if (round > kDecThreads + 1) asm volatile("trap;");   // cannot happen: thread t is final after t rounds
        if (redo) {
            exitq = dec_run(win, s_lut, s_cnt, s_nodes, entry, q_hi, q_avail, count);
            if (entry != kEnd32 && entry >= q_hi) { exitq = entry; count = 0; }
        }
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

    // CTA totals: letters, entry of the first active thread, exit of the last thread
    uint32_t c = active ? count : 0u;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, s);
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

struct DecSmem {
    uint32_t *win; uint16_t *lut; uint8_t *cnt; uint32_t *nodes; uint32_t *exit; uint32_t *red; uint32_t *lut2; uint32_t *stage;
};
constexpr size_t dec_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }
constexpr size_t kDecOffLut = dec_align16(kWinPhys * 4);
constexpr size_t kDecOffNodes = kDecOffLut + (1 << kLutBits) * 2;
constexpr size_t kDecOffRed = kDecOffNodes + dec_align16(HB_MAX_NODES * 4);
constexpr size_t kDecOffVar = kDecOffRed + 64;                       // count: cnt + exit ; write: lut2 + stage
constexpr size_t kDecSmemCount = kDecOffVar + (1 << kLutBits) + kDecThreads * 4;
constexpr int kStageWords = kOutWindow / 4 + 32;                     // + slack for the 4-byte alignment shift, XOR-swizzled rows
constexpr size_t kDecSmemWrite = kDecOffVar + (1 << kLutBits) * 4 + kStageWords * 4;

__device__ __forceinline__ DecSmem dec_carve(uint8_t *base) {
    DecSmem s;
    s.win = reinterpret_cast<uint32_t *>(base);
    s.lut = reinterpret_cast<uint16_t *>(base + kDecOffLut);
    s.nodes = reinterpret_cast<uint32_t *>(base + kDecOffNodes);
    s.red = reinterpret_cast<uint32_t *>(base + kDecOffRed);
    s.cnt = base + kDecOffVar;
    s.exit = reinterpret_cast<uint32_t *>(base + kDecOffVar + (1 << kLutBits));
    s.lut2 = reinterpret_cast<uint32_t *>(base + kDecOffVar);
    s.stage = reinterpret_cast<uint32_t *>(base + kDecOffVar + (1 << kLutBits) * 4);
    return s;
}

__global__ void __launch_bounds__(kDecThreads)
dec_count_kernel(DecParams p, const DecTables *__restrict__ tables) {
    DecSmem s = dec_carve(dec_smem);
    dec_load_tables(tables, s.lut, s.cnt, s.nodes);
    for (uint32_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x)
        dec_count_block(p, blk, 0, false, s.win, s.lut, s.cnt, s.nodes, s.exit, s.red);
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
    DecSmem s = dec_carve(dec_smem);
    __shared__ uint32_t s_next;
    dec_load_tables(tables, s.lut, s.cnt, s.nodes);
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
        const uint32_t j = s_next;
        if (j == 0xFFFFFFFFu) break;
        const volatile uint64_t *vexit = p.blk_exit;
        const uint64_t old_exit = vexit[j];
        const uint64_t entry = vexit[j - 1];
        __syncthreads();
        dec_count_block(p, j, entry, true, s.win, s.lut, s.cnt, s.nodes, s.exit, s.red);
        if (threadIdx.x == 0) {
            dirty[j] = 0;
            if (j + 1 < p.n_blocks && vexit[j] != old_exit) dirty[j + 1] = 1;
            __threadfence();
        }
        cur = j + 1;
    }
}

// ---- output offsets: group-local exclusive scan of CTA counts + scan of group totals
__global__ void __launch_bounds__(kDecThreads)
dec_scan_groups_kernel(const uint32_t *__restrict__ blk_count, uint32_t n_blocks, uint32_t *__restrict__ blk_local,
                       uint64_t *__restrict__ group_total) {
    __shared__ uint32_t s_w[kDecThreads / 32];
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
    for (int k = 0; k < kDecThreads / 32; k++) { if (k < (threadIdx.x >> 5)) before += s_w[k]; total += s_w[k]; }
    uint32_t run = before + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; k++) { if (base + k < n_blocks) blk_local[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) group_total[blockIdx.x] = total;
}

// single CTA: exclusive scan of group totals in place -> group offsets; writes the grand total
__global__ void __launch_bounds__(kDecThreads)
dec_scan_totals_kernel(uint64_t *group_total, uint32_t n_groups, uint64_t *grand_total) {
    __shared__ unsigned long long s_w[kDecThreads / 32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_groups; base += kDecThreads) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n_groups ? group_total[i] : 0ull;
        const unsigned long long incl = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long before = 0, total = 0;
        for (int k = 0; k < kDecThreads / 32; k++) { if (k < (threadIdx.x >> 5)) before += s_w[k]; total += s_w[k]; }
        const unsigned long long carry = s_carry;
        if (i < n_groups) group_total[i] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = s_carry;
}

// ---- write pass
// staging bytes live in 32-bit words whose index is XOR-swizzled by its row, so that lanes writing at a stride of a
// multiple of 128 bytes (uniform data: 128 letters per thread) still hit 32 different banks
__device__ __forceinline__ uint32_t stage_word(uint32_t w) { return w ^ ((w >> 5) & 31u); }
__device__ __forceinline__ void stage_put(uint32_t *stage, uint32_t a, uint32_t letter) {
    reinterpret_cast<uint8_t *>(stage)[(stage_word(a >> 2) << 2) | (a & 3u)] = static_cast<uint8_t>(letter);
}

__global__ void __launch_bounds__(kDecThreads)
dec_write_kernel(DecParams p, const DecTables *__restrict__ tables, const uint32_t *__restrict__ blk_local,
                 const uint64_t *__restrict__ group_off, uint8_t *__restrict__ out) {
    DecSmem s = dec_carve(dec_smem);
    __shared__ uint32_t s_w[kDecThreads / 32];
    dec_load_tables(tables, s.lut, nullptr, s.nodes, s.lut2);
    const int t = threadIdx.x;
    for (uint32_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
        __syncthreads();
        const uint32_t chunk = p.first_block + blk;
        dec_load_window(p, chunk, s.win);

        const uint32_t info = p.sub_info[blk * kDecThreads + t];
        const uint32_t my_count = info & 0xFFFFu;
        const uint32_t entry_rel = info >> 16;
        const long long win_bit0 = (static_cast<long long>(chunk) * kChunkWords - kHaloWords) * 32;
        const long long q_av = static_cast<long long>(p.avail_bits) - win_bit0;
        const uint32_t q_avail = q_av < 0 ? 0u : (q_av > static_cast<long long>(kWinBits) ? kWinBits : static_cast<uint32_t>(q_av));

        // CTA-wide exclusive scan of letter counts
        const uint32_t incl = warp_incl_scan(my_count);
        if ((t & 31) == 31) s_w[t >> 5] = incl;
        __syncthreads();                                     // also: window staged
        uint32_t before = 0, total = 0;
        for (int k = 0; k < kDecThreads / 32; k++) { if (k < (t >> 5)) before += s_w[k]; total += s_w[k]; }
        const uint32_t my_off = before + incl - my_count;
        const uint64_t out_base = group_off[blk / kScanGroup] + blk_local[blk];

        BitReader rd;
        if (my_count) rd.init(s.win, (kHaloWords + t * kSubWords) * 32u + entry_rel);
        uint32_t produced = 0;
        for (uint32_t win_start = 0; win_start < total; win_start += kOutWindow) {
            const uint32_t win_len = min(static_cast<uint32_t>(kOutWindow), total - win_start);
            uint8_t *dst = out + out_base + win_start;
            const uint32_t shift = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst) & 3);   // word-align stage to dst
            // my letters that fall into [win_start, win_start + win_len)
            uint32_t limit = min(my_count, win_start + win_len > my_off ? win_start + win_len - my_off : 0u);
            uint32_t a = shift + my_off + produced - win_start;                                  // staging byte address
            while (produced < limit) {
                rd.refill();
                const uint32_t e = s.lut2[rd.peek()];
                const uint32_t len0 = (e >> 16) & 0xFFu;
                if (len0) {
                    const uint32_t len01 = e >> 24;
                    stage_put(s.stage, a, e & 0xFFu);
                    if (len01 && produced + 1 < limit) {
                        stage_put(s.stage, a + 1, (e >> 8) & 0xFFu);
                        rd.skip(len01);
                        produced += 2; a += 2;
                    } else {
                        rd.skip(len0);
                        produced += 1; a += 1;
                    }
                } else {                                                                          // code longer than 12 bits
                    uint32_t letter = 0;
                    const uint32_t len = dec_one(s.win, s.lut, s.nodes, rd.q, q_avail, letter);
                    stage_put(s.stage, a, letter);
                    produced += 1; a += 1;
                    rd.init(s.win, rd.q + len);
                }
            }
            __syncthreads();
            // copy out whole 32-bit words; the ragged first / last word byte by byte
            const uint32_t n_words = (shift + win_len + 3) / 4;
            uint8_t *dst_al = dst - shift;
            for (uint32_t w = t; w < n_words; w += kDecThreads) {
                const uint32_t v = s.stage[stage_word(w)];
                const uint32_t lo = w * 4, hi = lo + 4;
                if (lo >= shift && hi <= shift + win_len) {
                    st_stream_u32(reinterpret_cast<uint32_t *>(dst_al + lo), v);
                } else {
                    for (uint32_t k = max(lo, shift); k < min(hi, shift + win_len); k++)
                        dst_al[k] = static_cast<uint8_t>(v >> (8 * (k - lo)));
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace hb
