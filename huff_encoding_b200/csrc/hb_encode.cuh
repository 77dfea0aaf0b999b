// hb_encode.cuh -- K2: the packing loop of compress_with_tree (comp.rs:422-447) as one single-pass kernel.
//
// The reference appends each letter's code bit by bit, MSB first, to one gap-free stream.  Here:
//   * persistent CTAs take 4096-letter tiles in ticket order; a thread owns 16 consecutive letters (one
//     128-bit load), looks their (code, len) up in a shared-memory copy of the code table and sums the lengths;
//   * a CTA-wide scan gives every thread its bit offset inside the tile, a decoupled look-back over one
//     64-bit descriptor per tile gives the tile its 64-bit GLOBAL bit offset (inputs > 2^32 bits are in scope);
//   * threads shift-merge their codes into a shared-memory staging stream (plain 32-bit stores for words they
//     fully own, shared-memory atomicOr for the first/last partial word);
//   * the tile is copied out with coalesced 32-bit big-endian stores.  The staging stream is kept tile-local
//     (bit 0 = the tile's first bit) and funnel-shifted by (global offset % 32) on the way out, and every tile
//     publishes the last 32 bits of its local stream next to its descriptor, so the output word that straddles
//     two tiles is written exactly once, by the later tile: no pre-zeroed output, no global atomics, no second pass.
//
// Algorithmic HBM bytes per launch: N (letters read once) + C (stream written once).
// Code lengths up to 32 bits take the narrow path; 33..64 bits (e.g. the Fibonacci edge case, 40 bits) the wide
// path, which feeds every letter as a (high part, low 32 bits) pair through the same merge.
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kEncThreads = 256;
constexpr int kEncLettersPerThread = 16;
constexpr int kEncTile = kEncThreads * kEncLettersPerThread;        // 4096 letters

constexpr uint64_t kDescAggregate = 1ull << 62;
constexpr uint64_t kDescPrefix = 2ull << 62;
constexpr uint64_t kDescValueMask = (1ull << 62) - 1;

// device-resident code table: lo[b] = (low <=32 code bits, their count), hi[b] = (bits above 32, their count)
struct EncTable {
    uint2 lo[256];
    uint2 hi[256];
};

struct EncScratch {
    uint64_t *desc;        // one per tile, zeroed before the launch
    uint32_t *tails;       // last 32 bits of each tile's local stream
    uint32_t *ticket;      // zeroed before the launch
};

template <bool WIDE>
__global__ void __launch_bounds__(kEncThreads)
encode_tiles_kernel(const uint8_t *__restrict__ data, size_t n, const EncTable *__restrict__ table,
                    uint32_t start_bit, uint32_t *__restrict__ out32, EncScratch scratch, uint32_t n_tiles,
                    unsigned long long *__restrict__ total_bits_out) {
    constexpr int kMaxTileWords = kEncTile * (WIDE ? 2 : 1);        // worst case: every letter 32 (64) bits
    __shared__ uint2 s_lo[256];
    __shared__ uint2 s_hi[WIDE ? 256 : 1];
    __shared__ uint32_t s_stage[kMaxTileWords + 4];                  // [0] = predecessor tail, [1 + m] = local word m
    __shared__ uint32_t s_warp_bits[kEncThreads / 32];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_prefix;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    s_lo[tid] = table->lo[tid];
    if (WIDE) s_hi[tid] = table->hi[tid];

    for (;;) {
        __syncthreads();                                             // previous tile fully copied out; tables visible
        if (tid == 0) s_tile = atomicAdd(scratch.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;

        // ---- load 16 letters, look up codes
        const size_t base = static_cast<size_t>(tile) * kEncTile + static_cast<size_t>(tid) * kEncLettersPerThread;
        uint32_t w[4] = {0, 0, 0, 0};
        int valid = 0;
        if (base + kEncLettersPerThread <= n) {
            uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(data + base));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            valid = kEncLettersPerThread;
        } else if (base < n) {
            valid = static_cast<int>(n - base);
#pragma unroll
            for (int j = 0; j < kEncLettersPerThread; j++)           // static indices keep w[] in registers
                if (j < valid) w[j >> 2] |= static_cast<uint32_t>(data[base + j]) << (8 * (j & 3));
        }
        uint32_t code[kEncLettersPerThread], len[kEncLettersPerThread];
        uint32_t code_hi[WIDE ? kEncLettersPerThread : 1], len_hi[WIDE ? kEncLettersPerThread : 1];
        uint32_t my_bits = 0;
#pragma unroll
        for (int j = 0; j < kEncLettersPerThread; j++) {
            const uint32_t b = byte_of(w[j >> 2], j & 3);
            const bool ok = j < valid;
            uint2 e = s_lo[b];
            code[j] = ok ? e.x : 0u;
            len[j] = ok ? e.y : 0u;
            my_bits += len[j];
            if (WIDE) {
                uint2 h = s_hi[b];
                code_hi[j] = ok ? h.x : 0u;
                len_hi[j] = ok ? h.y : 0u;
                my_bits += len_hi[j];
            }
        }

        // ---- CTA-wide exclusive scan of bit counts
        const uint32_t incl = warp_incl_scan(my_bits);
        if (lane == 31) s_warp_bits[warp] = incl;
        __syncthreads();
        uint32_t before = 0, tile_bits = 0;
#pragma unroll
        for (int k = 0; k < kEncThreads / 32; k++) {
            const uint32_t wb = s_warp_bits[k];
            if (k < warp) before += wb;
            tile_bits += wb;
        }
        const uint32_t my_off = before + incl - my_bits;

        // ---- clear the part of the staging stream this tile will touch
        const uint32_t n_local_words = (tile_bits + 31) / 32 + 2;
        for (uint32_t i = tid; i < n_local_words; i += kEncThreads) s_stage[1 + i] = 0;
        __syncthreads();

        // ---- shift-merge this thread's codes into the staging stream
        {
            uint32_t widx = 1 + (my_off >> 5);
            uint32_t nb = my_off & 31;                 // bits of the current word that belong to earlier threads
            uint64_t acc = 0;
            bool first = true;
            auto append = [&](uint32_t c, uint32_t l) {
                acc = (acc << l) | c;
                nb += l;
                if (nb >= 32) {
                    nb -= 32;
                    const uint32_t word = static_cast<uint32_t>(acc >> nb);
                    if (first) atomicOr(&s_stage[widx], word); else s_stage[widx] = word;
                    first = false;
                    widx++;
                }
            };
#pragma unroll
            for (int j = 0; j < kEncLettersPerThread; j++) {
                if (WIDE) append(code_hi[j], len_hi[j]);
                append(code[j], len[j]);
            }
            if (nb) atomicOr(&s_stage[widx], static_cast<uint32_t>(acc << (32 - nb)));
        }
        __syncthreads();

        // ---- publish (aggregate, tail), look back for the exclusive global bit offset
        if (warp == 0) {
            unsigned long long excl = 0;
            uint32_t pred_tail = 0;
            if (lane == 0) {
                const uint32_t q = tile_bits & 31;
                const uint32_t last = tile_bits ? (tile_bits - 1) >> 5 : 0;
                // last 32 bits of the local stream (only consumed when tile_bits >= 32, i.e. for full tiles)
                const uint32_t tail = q ? __funnelshift_l(s_stage[1 + last], s_stage[last], q) : s_stage[1 + last];
                scratch.tails[tile] = tail;
            }
            if (tile == 0) {
                excl = start_bit;
                if (lane == 0) st_release_u64(scratch.desc, kDescPrefix | (excl + tile_bits));
            } else {
                if (lane == 0) st_release_u64(scratch.desc + tile, kDescAggregate | tile_bits);
                long long look = static_cast<long long>(tile) - 1;
                for (;;) {
                    const long long idx = look - lane;
                    unsigned long long d = kDescPrefix;            // virtual "prefix 0" before tile 0 (never the nearest)
                    if (idx >= 0) {
                        uint32_t polls = 0;
                        do {
                            d = ld_acquire_u64(scratch.desc + idx);
                            if (++polls == (1u << 26)) asm volatile("trap;");   // a lost descriptor must not hang the GPU
                        } while ((d >> 62) == 0);
                    }
                    const unsigned pm = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2);
                    const int first_prefix = pm ? __ffs(pm) - 1 : 32;
                    unsigned long long contrib = (lane <= first_prefix) ? (d & kDescValueMask) : 0ull;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, s);
                    excl += contrib;
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) {
                    st_release_u64(scratch.desc + tile, kDescPrefix | (excl + tile_bits));
                    pred_tail = ld_relaxed_u32(scratch.tails + tile - 1);   // ordered after lane 0's acquire of desc[tile-1]
                }
            }
            if (lane == 0) {
                s_prefix = excl;
                s_stage[0] = pred_tail;
                if (tile == n_tiles - 1 && total_bits_out) *total_bits_out = excl + tile_bits - start_bit;
            }
        }
        __syncthreads();

        // ---- copy out: global word W0+m = funnel(local[m-1], local[m]) >> r, stored big-endian
        const unsigned long long gbit = s_prefix;
        const uint32_t r = static_cast<uint32_t>(gbit & 31);
        const unsigned long long w0 = gbit >> 5;
        const unsigned long long end_bit = gbit + tile_bits;
        const uint32_t n_full = static_cast<uint32_t>((end_bit >> 5) - w0);
        uint32_t *dst = out32 + w0;
        for (uint32_t m = tid; m < n_full; m += kEncThreads) {
            const uint32_t word = __funnelshift_r(s_stage[1 + m], s_stage[m], r);
            dst[m] = bswap32(word);
        }
        if (tile == n_tiles - 1 && (end_bit & 31) && tid == 0) {
            // the stream's final partial word: pad bits are zero (comp.rs:446-447), write only the bytes that exist
            const uint32_t word = __funnelshift_r(s_stage[1 + n_full], s_stage[n_full], r);
            const uint32_t n_bytes = (static_cast<uint32_t>(end_bit & 31) + 7) / 8;
            uint8_t *dst8 = reinterpret_cast<uint8_t *>(dst + n_full);
            for (uint32_t k = 0; k < n_bytes; k++) dst8[k] = static_cast<uint8_t>(word >> (24 - 8 * k));
        }
    }
}

}  // namespace hb
