// hb_encode.cuh -- K2: the packing loop of compress_with_tree (comp.rs:422-447) as one single-pass kernel.
//
// The reference appends each letter's code bit by bit, MSB first, to one gap-free stream.  Here every WARP is
// autonomous (no CTA barrier in the steady state):
//   * one persistent CTA per SM; warp g of the grid owns tiles g, g + G, g + 2G, ... (G = warps in the grid, all
//     co-resident: the kernel is launched cooperatively so the look-back below can never wait on a warp that is
//     not running);
//   * a tile is 32 lanes x 8 rounds of CHUNKS; a chunk is S consecutive letters (S = 4 when every code has <= 16
//     bits, 2 for <= 32 bits, 1 for <= 64 bits) merged into one <= 64-bit value.  In round r lane i takes chunk
//     r*32 + i, so the 32 lanes of a round read 32*S consecutive bytes (coalesced) and write adjacent stream words;
//   * codes come from a lane-replicated shared-memory table: entry (b, lane) lives at [b*32 + lane] as (code, len),
//     so the 32 lookups of a warp never conflict whatever the data;
//   * a warp scan of the chunk lengths (two rounds packed per 32-bit scan) gives every chunk its bit offset inside the
//     tile; the tile total is published at once (status AGGREGATE) so successors never wait on the packing;
//   * chunks are OR-ed into a per-warp shared-memory staging stream (<= 3 shared atomics per chunk);
//   * a decoupled look-back over one 64-bit descriptor per tile yields the tile's 64-bit GLOBAL bit offset;
//   * the staging stream is tile-local; on the way out it is funnel-shifted by (global offset % 32) and stored as
//     coalesced big-endian 32-bit words.  The word two tiles share is written once, by the later tile, which
//     re-derives the last 32 bits of its predecessor from the predecessor's last 32 letters (every code has >= 1
//     bit) instead of waiting for them: no pre-zeroed output, no global atomics, no second pass over the input.
//
// Algorithmic HBM bytes per launch: N (letters read once) + C (stream written once).
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kEncWarps = 32;
constexpr int kEncThreads = kEncWarps * 32;
constexpr int kEncRounds = 8;                                   // chunks per lane per tile
constexpr int kEncStageWords = 32 * kEncRounds * 2 + 8;         // 256 chunks x 64 bits, + tail slot + slack

constexpr uint64_t kDescAggregate = 1ull << 62;
constexpr uint64_t kDescPrefix = 2ull << 62;
constexpr uint64_t kDescValueMask = (1ull << 62) - 1;

// device-resident code table
struct EncTable {
    uint2 lo[256];        // (low <= 32 code bits, len) -- for len <= 32 this is the whole code
    uint32_t hi[256];     // code bits above 32 (len > 32 only)
};

constexpr size_t enc_smem_bytes(int S) {
    return 256 * 32 * sizeof(uint2) + (S == 1 ? 256 * sizeof(uint32_t) : 0) + kEncWarps * kEncStageWords * sizeof(uint32_t);
}

template <int S> struct EncLoad;
template <> struct EncLoad<4> { using type = uint32_t; };
template <> struct EncLoad<2> { using type = uint16_t; };
template <> struct EncLoad<1> { using type = uint8_t; };

extern __shared__ __align__(16) uint8_t enc_smem[];

template <int S>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_warp_tiles_kernel(const uint8_t *__restrict__ data, size_t n, const EncTable *__restrict__ table,
                         uint32_t start_bit, uint32_t *__restrict__ out32, uint64_t *__restrict__ desc,
                         uint32_t n_tiles, unsigned long long *__restrict__ total_bits_out) {
    constexpr int kTile = 32 * kEncRounds * S;                  // letters per tile: 1024 / 512 / 256
    using load_t = typename EncLoad<S>::type;

    uint2 *s_tab = reinterpret_cast<uint2 *>(enc_smem);                                     // [256][32]
    uint32_t *s_hi = reinterpret_cast<uint32_t *>(enc_smem + 256 * 32 * sizeof(uint2));    // [256] (S == 1)
    uint32_t *s_stage_all = s_hi + (S == 1 ? 256 : 0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 32; i += kEncThreads) s_tab[i] = table->lo[i >> 5];
    if (S == 1) for (int i = threadIdx.x; i < 256; i += kEncThreads) s_hi[i] = table->hi[i];
    __syncthreads();                                            // the only CTA barrier

    uint32_t *stage = s_stage_all + warp * kEncStageWords;      // [0] = predecessor tail, [1 + m] = local word m
    const uint2 *my_tab = s_tab + lane;
    const uint32_t grid_warps = gridDim.x * kEncWarps;

    for (uint32_t tile = blockIdx.x * kEncWarps + warp; tile < n_tiles; tile += grid_warps) {
        const size_t tile_base = static_cast<size_t>(tile) * kTile;
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

        // ---- publish the aggregate right away; successors only need this to get past us
        if (lane == 0) {
            if (tile == 0) st_release_u64(desc, kDescPrefix | (static_cast<unsigned long long>(start_bit) + tile_bits));
            else st_release_u64(desc + tile, kDescAggregate | tile_bits);
        }

        // ---- OR the chunks into the zeroed staging stream
        const uint32_t n_local_words = (tile_bits + 31) / 32 + 3;
        for (uint32_t i = lane; i < n_local_words; i += 32) stage[1 + i] = 0;
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kEncRounds; r++) {
            const uint32_t L = len[r];
            if (L == 0) continue;                               // only tail tiles / letters without a code
            const unsigned long long top = val[r] << (64 - L);  // left-aligned chunk
            const uint32_t hi = static_cast<uint32_t>(top >> 32), lo = static_cast<uint32_t>(top);
            const uint32_t w = 1 + (off[r] >> 5), s = off[r] & 31;
            atomicOr(&stage[w], hi >> s);
            if (s + L > 32) atomicOr(&stage[w + 1], __funnelshift_r(lo, hi, s));
            if (s + L > 64) atomicOr(&stage[w + 2], __funnelshift_r(0u, lo, s));
        }
        __syncwarp();

        // ---- last 32 bits of the predecessor tile, recomputed from its last 32 letters
        uint32_t pred_tail = 0;
        if (tile > 0) {
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

        // ---- decoupled look-back: exclusive global bit offset of this tile
        unsigned long long excl = start_bit;
        if (tile > 0) {
            excl = 0;
            long long look = static_cast<long long>(tile) - 1;
            for (;;) {
                const long long idx = look - lane;
                unsigned long long d = kDescPrefix;            // virtual "prefix 0" before tile 0 (never the nearest)
                if (idx >= 0) {
                    uint32_t polls = 0;
                    do {
                        d = ld_acquire_u64(desc + idx);
                        if (++polls == (1u << 26)) asm volatile("trap;");   // a lost descriptor must not hang the GPU
                    } while ((d >> 62) == 0);
                }
                const unsigned pm = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                const int first_prefix = pm ? __ffs(pm) - 1 : 32;
                unsigned long long contrib = (lane <= first_prefix) ? (d & kDescValueMask) : 0ull;
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, sft);
                excl += contrib;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st_release_u64(desc + tile, kDescPrefix | (excl + tile_bits));
        }
        if (lane == 0) {
            stage[0] = pred_tail;
            if (tile == n_tiles - 1 && total_bits_out) *total_bits_out = excl + tile_bits - start_bit;
        }
        __syncwarp();

        // ---- copy out: global word W0+m = funnel(local[m-1], local[m]) >> r, stored big-endian
        const uint32_t rr = static_cast<uint32_t>(excl & 31);
        const unsigned long long w0 = excl >> 5;
        const unsigned long long end_bit = excl + tile_bits;
        const uint32_t n_full = static_cast<uint32_t>((end_bit >> 5) - w0);
        uint32_t *dst = out32 + w0;
        for (uint32_t m = lane; m < n_full; m += 32)
            st_stream_u32(dst + m, bswap32(__funnelshift_r(stage[1 + m], stage[m], rr)));
        if (tile == n_tiles - 1 && (end_bit & 31) && lane == 0) {
            // the stream's final partial word: pad bits are zero (comp.rs:446-447), write only the bytes that exist
            const uint32_t word = __funnelshift_r(stage[1 + n_full], stage[n_full], rr);
            const uint32_t n_bytes = (static_cast<uint32_t>(end_bit & 31) + 7) / 8;
            uint8_t *dst8 = reinterpret_cast<uint8_t *>(dst + n_full);
            for (uint32_t k = 0; k < n_bytes; k++) dst8[k] = static_cast<uint8_t>(word >> (24 - 8 * k));
        }
        __syncwarp();                                           // staging is reused by the next tile
    }
}

}  // namespace hb
