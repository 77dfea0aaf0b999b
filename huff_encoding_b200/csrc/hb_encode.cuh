// hb_encode.cuh -- K2: the packing loop of compress_with_tree (comp.rs:422-447) as one single-pass kernel.
//
// The reference appends each letter's code bit by bit, MSB first, to one gap-free stream.  Here:
//   * the input is cut into one contiguous REGION per CTA (one persistent CTA per SM).  The histogram kernel
//     (hb_hist.cuh) has already counted every region separately, so the exact bit offset of region r is
//     sum_{i<r} sum_b region_hist[i][b] * len[b]: every CTA computes its own base in a short prologue.  There is no
//     inter-CTA communication at all (no look-back descriptors, no cooperative launch, nothing to wait on);
//   * inside a region the 32 warps of the CTA take 32 consecutive TILES per round; one __syncthreads per round
//     exchanges the 32 tile bit totals through shared memory and every warp derives its 64-bit global bit offset;
//   * a tile is 32 lanes x 8 rounds of CHUNKS; a chunk is S consecutive letters (S = 4 when every code has <= 16
//     bits, 2 for <= 32 bits, 1 for <= 64 bits) merged into one <= 64-bit value.  Lane i takes chunk r*32 + i in
//     round r, so the 32 lanes read 32*S consecutive bytes (coalesced) and write adjacent stream words;
//   * codes come from a lane-replicated shared-memory table: entry (b, lane) lives at [b*32 + lane] as (code, len),
//     so the 32 lookups of a warp never conflict whatever the data;
//   * a warp scan of the chunk lengths (two rounds packed per 32-bit scan) gives every chunk its bit offset inside the
//     tile; chunks are OR-ed into a per-warp shared-memory staging stream (<= 3 shared atomics per chunk);
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
constexpr int kEncRounds = 8;                                   // chunks per lane per tile
constexpr int kEncStageWords = 32 * kEncRounds * 2 + 8;         // 256 chunks x 64 bits, + tail slot + slack

constexpr int kEncRoundLetters = kEncWarps * 32 * kEncRounds * 4;   // region sizes are multiples of this (S = 4 round)

// device-resident code table
struct EncTable {
    uint2 lo[256];        // (low <= 32 code bits, len) -- for len <= 32 this is the whole code
    uint32_t hi[256];     // code bits above 32 (len > 32 only)
};

constexpr size_t enc_smem_bytes(int S) {
    return 256 * 32 * sizeof(uint2) + (S == 1 ? 256 * sizeof(uint32_t) : 0) + kEncWarps * kEncStageWords * sizeof(uint32_t);
}

// exact bit count of a region from its histogram: sum_b hist[b] * len[b]  (block-wide, all threads get the result)
__device__ __forceinline__ unsigned long long enc_region_base(const uint32_t *__restrict__ region_hist, uint32_t n_before,
                                                              const uint2 *s_tab_lane0, unsigned long long *s_red) {
    unsigned long long acc = 0;
    for (uint32_t k = threadIdx.x; k < n_before * 256u; k += blockDim.x)
        acc += static_cast<unsigned long long>(region_hist[k]) * s_tab_lane0[(k & 255u) << 5].y;
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, sft);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    unsigned long long total = 0;
    for (uint32_t k = 0; k < blockDim.x / 32; k++) total += s_red[k];
    __syncthreads();
    return total;
}

template <int S> struct EncLoad;
template <> struct EncLoad<4> { using type = uint32_t; };
template <> struct EncLoad<2> { using type = uint16_t; };
template <> struct EncLoad<1> { using type = uint8_t; };

extern __shared__ __align__(16) uint8_t enc_smem[];

template <int S>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_regions_kernel(const uint8_t *__restrict__ data, size_t n, const EncTable *__restrict__ table,
                      uint32_t start_bit, uint32_t *__restrict__ out32, const uint32_t *__restrict__ region_hist,
                      size_t region_letters, unsigned long long *__restrict__ total_bits_out) {
    constexpr int kTile = 32 * kEncRounds * S;                  // letters per tile: 1024 / 512 / 256
    using load_t = typename EncLoad<S>::type;

    uint2 *s_tab = reinterpret_cast<uint2 *>(enc_smem);                                     // [256][32]
    uint32_t *s_hi = reinterpret_cast<uint32_t *>(enc_smem + 256 * 32 * sizeof(uint2));    // [256] (S == 1)
    uint32_t *s_stage_all = s_hi + (S == 1 ? 256 : 0);
    __shared__ unsigned long long s_red[kEncWarps];
    __shared__ uint32_t s_tile_bits[2][kEncWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 32; i += kEncThreads) s_tab[i] = table->lo[i >> 5];
    if (S == 1) for (int i = threadIdx.x; i < 256; i += kEncThreads) s_hi[i] = table->hi[i];
    __syncthreads();

    const size_t region_begin = static_cast<size_t>(blockIdx.x) * region_letters;
    if (region_begin >= n) return;
    const size_t region_end = min(n, region_begin + region_letters);
    // global bit offset of this region: everything the regions before it emit (+ the caller's start bit)
    unsigned long long running = start_bit + enc_region_base(region_hist, blockIdx.x, s_tab, s_red);

    uint32_t *stage = s_stage_all + warp * kEncStageWords;      // word m = global stream word (tile offset / 32) + m
    const uint2 *my_tab = s_tab + lane;
    const uint32_t n_rounds = static_cast<uint32_t>((region_end - region_begin + kEncWarps * kTile - 1) / (kEncWarps * kTile));

    for (uint32_t round = 0; round < n_rounds; round++) {
        const size_t tile_base = region_begin + (static_cast<size_t>(round) * kEncWarps + warp) * kTile;
        const bool live = tile_base < region_end;               // region_end == n whenever a tile is cut short
        const bool full = tile_base + kTile <= n;

        // ---- load this lane's 8 chunks (round r: chunk r*32 + lane) and merge each chunk's codes
        unsigned long long val[kEncRounds];
        uint32_t len[kEncRounds];
        {
            load_t raw[kEncRounds];
            const load_t *src = reinterpret_cast<const load_t *>(data + tile_base);
            if (full) {
#pragma unroll
                for (int r = 0; r < kEncRounds; r++) raw[r] = src[r * 32 + lane];
            } else {
#pragma unroll
                for (int r = 0; r < kEncRounds; r++) {
                    uint32_t v = 0;
                    const size_t at = tile_base + static_cast<size_t>(r * 32 + lane) * S;
#pragma unroll
                    for (int k = 0; k < S; k++)
                        if (at + k < n) v |= static_cast<uint32_t>(data[at + k]) << (8 * k);
                    raw[r] = static_cast<load_t>(v);
                }
            }
#pragma unroll
            for (int r = 0; r < kEncRounds; r++) {
                const size_t at = tile_base + static_cast<size_t>(r * 32 + lane) * S;
                if (S == 4) {
                    const uint32_t w = raw[r];
                    uint2 e0 = my_tab[(w & 0xFFu) << 5], e1 = my_tab[((w >> 8) & 0xFFu) << 5];
                    uint2 e2 = my_tab[((w >> 16) & 0xFFu) << 5], e3 = my_tab[(w >> 24) << 5];
                    if (!full) {
                        if (at + 0 >= n) e0 = make_uint2(0, 0);
                        if (at + 1 >= n) e1 = make_uint2(0, 0);
                        if (at + 2 >= n) e2 = make_uint2(0, 0);
                        if (at + 3 >= n) e3 = make_uint2(0, 0);
                    }
                    const uint32_t v01 = (e0.x << e1.y) | e1.x, l01 = e0.y + e1.y;   // <= 32 bits
                    const uint32_t v23 = (e2.x << e3.y) | e3.x, l23 = e2.y + e3.y;
                    val[r] = (static_cast<unsigned long long>(v01) << l23) | v23;
                    len[r] = l01 + l23;
                } else if (S == 2) {
                    const uint32_t w = raw[r];
                    uint2 e0 = my_tab[(w & 0xFFu) << 5], e1 = my_tab[((w >> 8) & 0xFFu) << 5];
                    if (!full) {
                        if (at + 0 >= n) e0 = make_uint2(0, 0);
                        if (at + 1 >= n) e1 = make_uint2(0, 0);
                    }
                    val[r] = (static_cast<unsigned long long>(e0.x) << e1.y) | e1.x;
                    len[r] = e0.y + e1.y;
                } else {
                    const uint32_t b = raw[r];
                    uint2 e0 = my_tab[b << 5];
                    uint32_t h = s_hi[b];
                    if (!full && at >= n) { e0 = make_uint2(0, 0); h = 0; }
                    val[r] = (static_cast<unsigned long long>(h) << 32) | e0.x;
                    len[r] = e0.y;
                }
            }
        }

        // ---- bit offsets: warp scans, two rounds per 32-bit word (a round totals at most 32 * 64 bits)
        uint32_t off[kEncRounds];
        uint32_t tile_bits = 0;
#pragma unroll
        for (int r = 0; r < kEncRounds; r += 2) {
            const uint32_t packed = len[r] | (len[r + 1] << 16);
            const uint32_t incl = warp_incl_scan(packed);
            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const uint32_t excl = incl - packed;
            off[r] = tile_bits + (excl & 0xFFFFu);
            tile_bits += tot & 0xFFFFu;
            off[r + 1] = tile_bits + (excl >> 16);
            tile_bits += tot >> 16;
        }

        // ---- one barrier per round: exchange the 32 tile totals, derive this tile's global bit offset
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
            const uint32_t b = data[tile_base - 32 + lane];
            const uint2 e = my_tab[b << 5];
            // suffix sum of the lengths of the letters after mine
            uint32_t after = e.y;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, after, d);
                if (lane + d < 32) after += o;
            }
            after -= e.y;
            const uint32_t piece = after < 32 ? (e.x << after) : 0u;   // low code bits that land in the last 32 bits
            pred_tail = __reduce_or_sync(0xFFFFFFFFu, piece);
        }

        // ---- OR the chunks into the zeroed staging stream, already shifted by (global offset % 32): staging word m
        //      is global word W0 + m, whose first rr bits are the predecessor's last rr bits
        const uint32_t rr = static_cast<uint32_t>(excl & 31);
        const uint32_t n_local_words = (rr + tile_bits + 31) / 32 + 3;
        for (uint32_t i = lane; i < n_local_words; i += 32) stage[i] = 0;
        __syncwarp();
        if (lane == 0 && rr) atomicOr(&stage[0], pred_tail << (32 - rr));   // OR: other lanes' chunks share word 0
#pragma unroll
        for (int r = 0; r < kEncRounds; r++) {
            const uint32_t L = len[r];
            if (L == 0) continue;                               // only tail tiles / letters without a code
            const unsigned long long top = val[r] << (64 - L);  // left-aligned chunk
            const uint32_t hi = static_cast<uint32_t>(top >> 32), lo = static_cast<uint32_t>(top);
            const uint32_t at = off[r] + rr;
            const uint32_t w = at >> 5, s = at & 31;
            atomicOr(&stage[w], hi >> s);
            if (s + L > 32) atomicOr(&stage[w + 1], __funnelshift_r(lo, hi, s));
            if (s + L > 64) atomicOr(&stage[w + 2], __funnelshift_r(0u, lo, s));
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
        for (uint32_t m = lane; m < n_full; m += 32) st_stream_u32(dst + m, bswap32(stage[m]));
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
