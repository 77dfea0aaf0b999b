// hb_encode.cuh -- K2: the packing loop of compress_with_tree (comp.rs:422-447) as one single-pass kernel.
//
// The reference appends each letter's code bit by bit, MSB first, to one gap-free stream.  Here:
//   * the input is cut into one contiguous REGION per CTA (one persistent CTA per SM).  The histogram kernel
//     (hb_hist.cuh) has already counted every region separately, so the exact bit offset of region r is
//     sum_{i<r} sum_b region_hist[i][b] * len[b]: every CTA computes its own base in a short prologue.  There is no
//     inter-CTA communication at all (no look-back descriptors, no cooperative launch, nothing to wait on);
//   * inside a region the 32 warps of the CTA take 32 consecutive TILES per round; one __syncthreads per round
//     exchanges the 32 tile bit totals through shared memory and every warp derives its 64-bit global bit offset;
//   * a tile is 32 lanes x 32 (S = 4; else 16) CONSECUTIVE letters: one 256-bit load per lane (a warp reads 1 KiB), and
//     a lane's codes are contiguous in the stream, so ONE warp scan per tile (of the lanes' bit totals) places them;
//   * codes come from a lane-replicated shared-memory table: entry (b, lane) lives at [b*32 + lane] as (code
//     left-aligned in 32 bits, len), so the 32 lookups of a warp never conflict whatever the data;
//   * every lane streams its codes through a 64-bit register packer in PIECES of <= 32 bits (S = 4: all codes
//     <= 16 bits, a piece is two letters; S = 2: codes <= 32 bits, one letter; S = 1: codes <= 64 bits, a letter is
//     two pieces) and ORs each completed 32-bit word into the per-warp shared-memory staging stream: one shared
//     atomic per 32 stream bits (the words two lanes share need the OR; the others take it as a plain store);
//   * the staging stream is laid out at (global offset % 32), so its words ARE the global stream words and go out as
//     coalesced big-endian 32-bit stores.  The word two tiles share is written once, by the later tile, which
//     re-derives the last 32 bits of its predecessor from the predecessor's last 32 letters (every code has >= 1
//     bit): no pre-zeroed output, no global atomics, no second pass over the input.
//
// Algorithmic HBM bytes per launch: N (letters read once) + C (stream written once) (+ 1 KiB per region of counts).
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kEncWarps = 32;
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kEncRoundLetters = 32768;                         // region sizes are multiples of this (whole rounds)

// consecutive letters per lane: 32 (one 256-bit load) when a piece is two letters, else 16; 16 pieces either way
__host__ __device__ constexpr int enc_lane_letters(int S) { return S == 4 ? 32 : 16; }
__host__ __device__ constexpr int enc_tile_letters(int S) { return 32 * enc_lane_letters(S); }   // one warp
__host__ __device__ constexpr int enc_max_bits(int S) { return S == 4 ? 16 : (S == 2 ? 32 : 64); }
__host__ __device__ constexpr int enc_stage_words(int S) { return enc_tile_letters(S) * enc_max_bits(S) / 32 + 4; }   // + shared first word + slack
static_assert(kEncRoundLetters % (kEncWarps * enc_tile_letters(4)) == 0 && kEncRoundLetters % (kEncWarps * enc_tile_letters(1)) == 0,
              "regions must hold whole rounds");

// device-resident code table: codes LEFT-aligned
struct EncTable {
    uint2 lo[256];        // (first min(len, 32) code bits, left-aligned in 32 bits; len)
    uint32_t hi[256];     // the remaining len - 32 bits, left-aligned (len > 32 only)
    uint32_t packed[256]; // len <= 16 only: code left-aligned in the upper half | len  (one 32-bit lookup; else 0)
};

// S == 4 (the hot variant) looks letters up in a lane-replicated table of PACKED 32-bit entries (code << 16 | len: one
// shared-memory wavefront per warp lookup instead of two) pinned at the ABSOLUTE shared address 0x10000 with 256 bytes
// per letter: entry (b, lane) is at 0x10000 | b << 8 | lane << 2, which ONE byte-permute builds from the raw input
// word (no shift, mask, add).  The shared memory below 0x10000 (minus the static variables) is left unused; the
// staging and a plain 2 KiB (code, len) table for the once-per-tile lookups follow the table.
constexpr uint32_t kEncTabAbs = 0x10000u;
constexpr size_t kEncPackedBytes = 256 * 256;
constexpr size_t enc_smem_bytes(int S) {
    return (S == 4 ? static_cast<size_t>(kEncTabAbs) + kEncPackedBytes + 256 * sizeof(uint2)
                   : 256 * 32 * sizeof(uint2) + (S == 1 ? 256 * sizeof(uint32_t) : 0)) +
           static_cast<size_t>(kEncWarps) * enc_stage_words(S) * sizeof(uint32_t);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
    return r;
}

// exact bit count of a region from its histogram: sum_b hist[b] * len[b]  (block-wide, all threads get the result)
__device__ __forceinline__ unsigned long long enc_region_base(const uint32_t *__restrict__ region_hist, uint32_t n_before,
                                                              const uint2 *s_tab_lane0, int tab_shift, unsigned long long *s_red) {
    unsigned long long acc = 0;
    for (uint32_t k = threadIdx.x; k < n_before * 256u; k += blockDim.x)
        acc += static_cast<unsigned long long>(region_hist[k]) * s_tab_lane0[(k & 255u) << tab_shift].y;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, sft);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    unsigned long long total = 0;
    for (uint32_t k = 0; k < blockDim.x / 32; k++) total += s_red[k];
    __syncthreads();
    return total;
}

extern __shared__ __align__(16) uint8_t enc_smem[];

// 64-bit register bit packer of one lane.  `hi` holds the `fill` (< 32) pending bits left-aligned; a piece of L <= 32
// bits (left-aligned in `pl`, zeros below) is appended and a completed word is OR-ed into the staging stream.
struct EncPacker {
    uint32_t *ptr;
    uint32_t hi, fill;
    __device__ __forceinline__ void append(uint32_t pl, uint32_t L) {
        hi |= pl >> fill;
        const uint32_t lo = __funnelshift_r(0u, pl, fill);        // the bits of pl that fall past the pending word
        fill += L;
        if (fill >= 32) {
            atomicOr(ptr, hi);
            ptr++;
            hi = lo;
            fill -= 32;
        }
    }
    // a piece of L <= 64 bits, left-aligned in (ph:pl): up to two completed words; the second one lies entirely
    // inside this piece, so it is a plain store
    __device__ __forceinline__ void append64(uint32_t ph, uint32_t pl, uint32_t L) {
        const uint32_t w0 = hi | (ph >> fill);
        const uint32_t w1 = __funnelshift_r(pl, ph, fill);
        const uint32_t w2 = __funnelshift_r(0u, pl, fill);
        const uint32_t nf = fill + L;
        // straight-line: two predicated stores, two selects (lanes differ in how many words they complete)
        if (nf >= 32) atomicOr(ptr, w0);
        if (nf >= 64) ptr[1] = w1;
        hi = nf >= 64 ? w2 : (nf >= 32 ? w1 : w0);
        ptr += nf >> 5;
        fill = nf & 31;
    }
    __device__ __forceinline__ void finish() { if (fill) atomicOr(ptr, hi); }
};

// The one cut-short tile at the end of the input (n % tile != 0): letter by letter, no register arrays.
template <int S>
__device__ __noinline__ uint32_t enc_partial_bits(const uint8_t *__restrict__ data, size_t lane_base, uint32_t n_mine,
                                                  const uint2 *my_tab) {
    uint32_t bits = 0;
    constexpr int kTabShift = S == 4 ? 0 : 5;
    for (uint32_t j = 0; j < n_mine; j++) bits += my_tab[static_cast<uint32_t>(data[lane_base + j]) << kTabShift].y;
    return bits;
}
template <int S>
__device__ __noinline__ void enc_partial_append(const uint8_t *__restrict__ data, size_t lane_base, uint32_t n_mine,
                                                const uint2 *my_tab, const uint32_t *s_hi, uint32_t *ptr, uint32_t fill) {
    EncPacker pk;
    pk.ptr = ptr;
    pk.fill = fill;
    pk.hi = 0;
    constexpr int kTabShift = S == 4 ? 0 : 5;
    for (uint32_t j = 0; j < n_mine; j++) {
        const uint32_t b = data[lane_base + j];
        const uint2 e = my_tab[b << kTabShift];
        pk.append(e.x, min(e.y, 32u));
        if (S == 1 && e.y > 32) pk.append(s_hi[b], e.y - 32);
    }
    pk.finish();
}

template <int S>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_regions_kernel(const uint8_t *__restrict__ data, size_t n, const EncTable *__restrict__ table,
                      uint32_t start_bit, uint32_t *__restrict__ out32, const uint32_t *__restrict__ region_hist,
                      size_t region_letters, unsigned long long *__restrict__ total_bits_out) {
    constexpr int kLane = enc_lane_letters(S);
    constexpr int kTile = enc_tile_letters(S);
    constexpr int kPieces = 16;                                  // register-held pieces per lane
    constexpr int kStageWords = enc_stage_words(S);

    // S == 4: packed lane-replicated table at absolute shared address kEncTabAbs, then a plain [256] (code, len) table;
    // else: lane-replicated (code, len) table [256][32] at the start (+ the code tails for S == 1).  See enc_smem_bytes.
    constexpr int kTabShift = S == 4 ? 0 : 5;                   // index shift of the (code, len) table
    uint8_t *tab_bytes = enc_smem + (S == 4 ? kEncTabAbs - smem_addr(enc_smem) : 0u);
    uint2 *s_tab = reinterpret_cast<uint2 *>(tab_bytes + (S == 4 ? kEncPackedBytes : 0));
    uint32_t *s_hi = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_tab) + (S == 4 ? 256 : 256 * 32) * sizeof(uint2));   // [256] (S == 1)
    uint32_t *s_stage_all = s_hi + (S == 1 ? 256 : 0);
    __shared__ unsigned long long s_red[kEncWarps];
    __shared__ uint32_t s_tile_bits[2][kEncWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (S == 4) {
        uint32_t *packed = reinterpret_cast<uint32_t *>(tab_bytes);
        for (int i = threadIdx.x; i < 256 * 32; i += kEncThreads) packed[(i >> 5) * 64 + (i & 31)] = table->packed[i >> 5];
        for (int i = threadIdx.x; i < 256; i += kEncThreads) s_tab[i] = table->lo[i];
    } else {
        for (int i = threadIdx.x; i < 256 * 32; i += kEncThreads) s_tab[i] = table->lo[i >> 5];
    }
    if (S == 1) for (int i = threadIdx.x; i < 256; i += kEncThreads) s_hi[i] = table->hi[i];
    __syncthreads();

    const size_t region_begin = static_cast<size_t>(blockIdx.x) * region_letters;
    if (region_begin >= n) return;
    const size_t region_end = min(n, region_begin + region_letters);
    // global bit offset of this region: everything the regions before it emit (+ the caller's start bit)
    unsigned long long running = start_bit + enc_region_base(region_hist, blockIdx.x, s_tab, kTabShift, s_red);

    uint32_t *stage = s_stage_all + warp * kStageWords;         // word m = global stream word (tile offset / 32) + m
    const bool aligned32 = (reinterpret_cast<uintptr_t>(data) & 31) == 0;
    const uint2 *my_tab = s_tab + (S == 4 ? 0 : lane);
    const uint32_t n_rounds = static_cast<uint32_t>((region_end - region_begin + kEncWarps * kTile - 1) / (kEncWarps * kTile));

    // this lane's letters of a full tile: one 256-bit load (or 128-bit loads)
    uint32_t raw[kLane / 4];
    auto load_raw = [&](size_t at) {
        if (kLane == 32 && aligned32) {
            const u32x8 v = ld_stream_256(data + at);
#pragma unroll
            for (int j = 0; j < 8; j++) raw[j] = v.v[j];
        } else {
#pragma unroll
            for (int h = 0; h < kLane / 16; h++) {
                const uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(data + at) + h);
                raw[4 * h + 0] = v.x; raw[4 * h + 1] = v.y; raw[4 * h + 2] = v.z; raw[4 * h + 3] = v.w;
            }
        }
    };
    // S == 4 software-pipelines the loads: the registers of `raw` are refilled for the NEXT round as soon as this
    // round's lookups are done, so the load latency hides behind the scan, the barrier and the packing
    constexpr bool kPipelined = S == 4;
    if (kPipelined) {
        const size_t first = region_begin + static_cast<size_t>(warp) * kTile;
        if (first + kTile <= n && first < region_end) load_raw(first + static_cast<size_t>(lane) * kLane);
    }

    for (uint32_t round = 0; round < n_rounds; round++) {
        const size_t tile_base = region_begin + (static_cast<size_t>(round) * kEncWarps + warp) * kTile;
        const bool live = tile_base < region_end;               // region_end == n whenever a tile is cut short
        const bool full = tile_base + kTile <= n;
        const size_t lane_base = tile_base + static_cast<size_t>(lane) * kLane;

        if (!kPipelined && lane_base + static_cast<size_t>(kEncWarps) * kTile < n)   // next round's letters -> L2
            asm volatile("prefetch.global.L2 [%0];" :: "l"(data + lane_base + static_cast<size_t>(kEncWarps) * kTile));

        // one of the predecessor tile's last 32 letters (its tail bits are re-derived below): issued early
        const uint32_t pred_letter = (live && tile_base > 0) ? data[tile_base - 32 + lane] : 0u;

        // ---- pass 1: this lane's consecutive letters, table lookups; the pieces and their lengths (one byte each)
        //      stay in registers.  The single cut-short tile of the input takes the letter-by-letter path.
        uint32_t pv[kPieces];
        uint32_t plen[S == 4 ? 8 : kPieces / 4];                // S == 4: one length per quad; else a byte per piece
        uint32_t lane_bits = 0;
        const uint32_t n_mine = full ? kLane
                                     : (lane_base >= n ? 0u : static_cast<uint32_t>(min(static_cast<size_t>(kLane), n - lane_base)));
        if (full) {
            if (!kPipelined) load_raw(lane_base);
            if (S == 4) {
                // quads: four letters -> one <= 64-bit piece (qh:ql, left-aligned) and its length
                const uint32_t lane_addr = kEncTabAbs + (static_cast<uint32_t>(lane) << 2);
                constexpr uint32_t kCode = 0xFFFF0000u;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t w = raw[q];
                    const uint32_t e0 = lds_u32(__byte_perm(w, lane_addr, 0x7604));   // 0x10000 | byte << 8 | lane << 2
                    const uint32_t e1 = lds_u32(__byte_perm(w, lane_addr, 0x7614));
                    const uint32_t e2 = lds_u32(__byte_perm(w, lane_addr, 0x7624));
                    const uint32_t e3 = lds_u32(__byte_perm(w, lane_addr, 0x7634));
                    // entry = code << 16 | len (len <= 16): the wrapping funnel shift takes len from the low 5 bits
                    const uint32_t v01 = (e0 & kCode) | __funnelshift_r(e1 & kCode, 0u, e0), l01 = (e0 + e1) & 0x3Fu;
                    const uint32_t v23 = (e2 & kCode) | __funnelshift_r(e3 & kCode, 0u, e2), l23 = (e2 + e3) & 0x3Fu;
                    pv[2 * q] = v01 | __funnelshift_rc(v23, 0u, l01);
                    pv[2 * q + 1] = __funnelshift_rc(0u, v23, l01);
                    plen[q] = l01 + l23;
                    lane_bits += l01 + l23;
                }
                // `raw` is free: next round's tile (same warp), if it is a full one of this region
                const size_t next_base = tile_base + static_cast<size_t>(kEncWarps) * kTile;
                if (next_base + kTile <= n && next_base < region_end) load_raw(next_base + static_cast<size_t>(lane) * kLane);
            } else {
#pragma unroll
                for (int k = 0; k < kPieces / 4; k++) plen[k] = 0;
#pragma unroll
                for (int k = 0; k < kPieces; k++) {
                    const uint32_t b = (raw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                    const uint2 e = my_tab[b << kTabShift];
                    pv[k] = e.x;
                    plen[k >> 2] |= e.y << (8 * (k & 3));      // S == 1: up to 64, the second piece comes from s_hi
                    lane_bits += e.y;
                }
            }
        } else if (live) {
            lane_bits = enc_partial_bits<S>(data, lane_base, n_mine, my_tab);
        }

        // ---- bit offsets: ONE warp scan per tile (a lane totals at most 1024 bits)
        uint32_t tile_bits, off;
        {
            const uint32_t incl = warp_incl_scan(lane_bits);
            tile_bits = __shfl_sync(0xFFFFFFFFu, incl, 31);
            off = incl - lane_bits;
        }

        // ---- one barrier per round: exchange the 32 tile totals, derive this tile's global bit offset.  (A
        //      barrier-free hand-over from tile to tile through shared-memory flags was measured 2x SLOWER: 32
        //      dependent hops per round are longer than the round's work.)
        if (!live) tile_bits = 0;
        if (lane == 0) s_tile_bits[round & 1][warp] = tile_bits;
        __syncthreads();
        unsigned long long excl;
        {
            const uint32_t mine = s_tile_bits[round & 1][lane];
            const uint32_t incl = warp_incl_scan(mine);
            const uint32_t before = __shfl_sync(0xFFFFFFFFu, incl - mine, warp);
            const uint32_t round_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            excl = running + before;
            running += round_total;
        }
        if (!live) continue;

        // ---- last 32 bits of the predecessor tile, recomputed from its last 32 letters
        uint32_t pred_tail = 0;
        if (tile_base > 0) {
            const uint32_t b = pred_letter;
            const uint2 e = my_tab[b << kTabShift];
            // suffix sum of the lengths of the letters after mine
            uint32_t after = e.y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, after, d);
                if (lane + d < 32) after += o;
            }
            after -= e.y;
            // the low 32 bits of my code, right-aligned
            uint32_t low;
            if (S == 1 && e.y > 32)
                low = static_cast<uint32_t>(((static_cast<unsigned long long>(e.x) << 32) | s_hi[b]) >> (64 - e.y));
            else
                low = e.y ? e.x >> (32 - e.y) : 0u;
            const uint32_t piece = after < 32 ? (low << after) : 0u;     // the part that lands in the last 32 bits
            pred_tail = __reduce_or_sync(0xFFFFFFFFu, piece);
        }

        // ---- OR the lanes' bit strings into the zeroed staging stream, already shifted by (global offset % 32):
        //      staging word m is global word W0 + m, whose first rr bits are the predecessor's last rr bits
        const uint32_t rr = static_cast<uint32_t>(excl & 31);
        const uint32_t n_local_vecs = ((rr + tile_bits + 31) / 32 + 3 + 3) / 4;           // <= kStageWords / 4
#pragma unroll
        for (int k = 0; k < (kStageWords / 4 + 31) / 32; k++)
            if (lane + 32 * k < n_local_vecs) reinterpret_cast<uint4 *>(stage)[lane + 32 * k] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        if (lane == 0 && rr) atomicOr(&stage[0], pred_tail << (32 - rr));   // OR: other lanes' bits share word 0
        if (full) {
            const uint32_t at = off + rr;
            EncPacker pk;
            pk.ptr = stage + (at >> 5);
            pk.fill = at & 31;
            pk.hi = 0;
            if (S == 4) {
#pragma unroll
                for (int q = 0; q < 8; q++) pk.append64(pv[2 * q], pv[2 * q + 1], plen[q]);
            } else {
#pragma unroll
                for (int k = 0; k < kPieces; k++) {
                    const uint32_t l = (plen[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                    if (S == 1) {
                        pk.append(pv[k], min(l, 32u));
                        if (l > 32) {                            // rare: the code's remaining bits
                            const uint32_t b = (raw[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                            pk.append(s_hi[b], l - 32);
                        }
                    } else {
                        pk.append(pv[k], l);
                    }
                }
            }
            pk.finish();
        } else {
            const uint32_t at = off + rr;
            enc_partial_append<S>(data, lane_base, n_mine, my_tab, s_hi, stage + (at >> 5), at & 31);
        }
        const bool is_last_tile = tile_base + kTile >= n;
        if (lane == 0 && is_last_tile && total_bits_out) *total_bits_out = excl + tile_bits - start_bit;
        __syncwarp();

        // ---- copy out: coalesced big-endian 32-bit stores of the whole words; the last partial word stays with the
        //      next tile, except at the very end of the stream
        const unsigned long long w0 = excl >> 5;
        const unsigned long long end_bit = excl + tile_bits;
        const uint32_t n_full = static_cast<uint32_t>((end_bit >> 5) - w0);
        uint32_t *dst = out32 + w0;
        {
            const uint32_t *sp = stage + lane;
            uint32_t *gp = dst + lane;
#pragma unroll 1
            for (uint32_t m = lane; m < n_full; m += 32, sp += 32, gp += 32) st_stream_u32(gp, bswap32(*sp));
        }
        if (is_last_tile && (end_bit & 31) && lane == 0) {
            // the stream's final partial word: pad bits are zero (comp.rs:446-447), write only the bytes that exist
            const uint32_t word = stage[n_full];
            const uint32_t n_bytes = (static_cast<uint32_t>(end_bit & 31) + 7) / 8;
            uint8_t *dst8 = reinterpret_cast<uint8_t *>(dst + n_full);
            for (uint32_t k = 0; k < n_bytes; k++) dst8[k] = static_cast<uint8_t>(word >> (24 - 8 * k));
        }
        __syncwarp();                                           // staging is reused by the next tile
    }
}

}  // namespace hb
