// hb_fixed.cuh -- fast path for FIXED-LENGTH code sets (every code has the same length L, L in {1, 2, 4, 8}).
//
// A Huffman tree whose leaves all sit at depth L is a perfect tree: 2^L letters, every L-bit pattern is a code.
// This is what uniform data produces (BASELINE.json configs[1]: 256 letters, all codes 8 bits) and what one- and
// two-letter inputs produce (L = 1).  The packing loop of compress_with_tree (comp.rs:422-447) then degenerates to
// "replace each letter by its L-bit code" and decompress (comp.rs:487-519) to the inverse table lookup: no bit
// offsets to scan, no code-word boundaries to find.  Both directions are one streaming kernel.
//
// Algorithmic HBM bytes: N + C (encode), C + N (decode), moved exactly once.
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kFixThreads = 512;

#ifndef HB_FIX_UNROLL
#define HB_FIX_UNROLL 4
#endif

// L == 8: out[i] = table[in[i]] for n bytes; table lane-replicated in shared memory ([256][32] u32, conflict-free).
// in / out 32-byte aligned -> 256-bit loads and stores (one full sector per lane); else 128-bit.
__global__ void __launch_bounds__(kFixThreads)
fixed8_translate_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n,
                        const uint8_t *__restrict__ table) {
    __shared__ uint32_t s_tab[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += kFixThreads) s_tab[i] = table[i >> 5];
    __syncthreads();
    const uint32_t *my = s_tab + lane_id();
    auto tr = [&](uint32_t w) -> uint32_t {
        return my[(w & 0xFFu) << 5] | (my[((w >> 8) & 0xFFu) << 5] << 8) | (my[((w >> 16) & 0xFFu) << 5] << 16) |
               (my[(w >> 24) << 5] << 24);
    };
    const size_t stride = static_cast<size_t>(gridDim.x) * kFixThreads;
    size_t i = static_cast<size_t>(blockIdx.x) * kFixThreads + threadIdx.x;
    size_t done = 0;                                           // bytes handled by the vector loops
#ifndef HB_FIX_NO_V8
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 31) == 0) {
        const size_t n_vec = n / 32;
        for (; i + (HB_FIX_UNROLL - 1) * stride < n_vec; i += HB_FIX_UNROLL * stride) {
            u32x8 v[HB_FIX_UNROLL];
#pragma unroll
            for (int u = 0; u < HB_FIX_UNROLL; u++) v[u] = ld_stream_256(in + 32 * (i + u * stride));
#pragma unroll
            for (int u = 0; u < HB_FIX_UNROLL; u++) {
#pragma unroll
                for (int k = 0; k < 8; k++) v[u].v[k] = tr(v[u].v[k]);
                st_stream_256(out + 32 * (i + u * stride), v[u]);
            }
        }
        for (; i < n_vec; i += stride) {
            u32x8 v = ld_stream_256(in + 32 * i);
#pragma unroll
            for (int k = 0; k < 8; k++) v.v[k] = tr(v.v[k]);
            st_stream_256(out + 32 * i, v);
        }
        done = n_vec * 32;
    } else
#endif
    {
        const size_t n_vec = n / 16;
        const uint4 *src = reinterpret_cast<const uint4 *>(in);
        uint4 *dst = reinterpret_cast<uint4 *>(out);
        for (; i + (HB_FIX_UNROLL - 1) * stride < n_vec; i += HB_FIX_UNROLL * stride) {
            uint4 v[HB_FIX_UNROLL];
#pragma unroll
            for (int u = 0; u < HB_FIX_UNROLL; u++) v[u] = ld_stream_u4(src + i + u * stride);
#pragma unroll
            for (int u = 0; u < HB_FIX_UNROLL; u++)
                st_stream_u4(dst + i + u * stride, make_uint4(tr(v[u].x), tr(v[u].y), tr(v[u].z), tr(v[u].w)));
        }
        for (; i < n_vec; i += stride) {
            const uint4 v = ld_stream_u4(src + i);
            st_stream_u4(dst + i, make_uint4(tr(v.x), tr(v.y), tr(v.z), tr(v.w)));
        }
        done = n_vec * 16;
    }
    if (blockIdx.x == 0) {
        const size_t at = done + threadIdx.x;                  // < 32 leftover bytes
        if (at < n) out[at] = static_cast<uint8_t>(my[static_cast<uint32_t>(in[at]) << 5]);
    }
}

// L in {1, 2, 4}: one thread per stream byte.  code[letter] = L-bit code; letters beyond n contribute zero bits.
__global__ void __launch_bounds__(kFixThreads)
fixed_pack_kernel(const uint8_t *__restrict__ in, size_t n, uint8_t *__restrict__ out, size_t out_bytes, uint32_t L,
                  const uint8_t *__restrict__ code) {
    __shared__ uint8_t s_code[256];
    if (threadIdx.x < 256) s_code[threadIdx.x] = code[threadIdx.x];
    __syncthreads();
    const uint32_t per = 8 / L;
    for (size_t j = static_cast<size_t>(blockIdx.x) * kFixThreads + threadIdx.x; j < out_bytes;
         j += static_cast<size_t>(gridDim.x) * kFixThreads) {
        uint32_t b = 0;
        for (uint32_t k = 0; k < per; k++) {
            const size_t at = j * per + k;
            const uint32_t c = at < n ? s_code[in[at]] : 0u;
            b |= c << (8 - L * (k + 1));
        }
        out[j] = static_cast<uint8_t>(b);
    }
}

// inverse: letter[pattern] for every L-bit pattern; n_letters letters starting at byte `in` bit 0
__global__ void __launch_bounds__(kFixThreads)
fixed_unpack_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t n_letters, uint32_t L,
                    const uint8_t *__restrict__ letter) {
    __shared__ uint8_t s_letter[256];
    if (threadIdx.x < 256) s_letter[threadIdx.x] = letter[threadIdx.x];
    __syncthreads();
    const uint32_t per = 8 / L, mask = (1u << L) - 1;
    const size_t n_bytes = (n_letters + per - 1) / per;
    for (size_t j = static_cast<size_t>(blockIdx.x) * kFixThreads + threadIdx.x; j < n_bytes;
         j += static_cast<size_t>(gridDim.x) * kFixThreads) {
        const uint32_t b = in[j];
        for (uint32_t k = 0; k < per; k++) {
            const size_t at = j * per + k;
            if (at < n_letters) out[at] = s_letter[(b >> (8 - L * (k + 1))) & mask];
        }
    }
}

}  // namespace hb
