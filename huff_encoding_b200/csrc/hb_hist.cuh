// hb_hist.cuh -- K1: 256-bin byte histogram (replaces build_weights_map, weights.rs:116-123, and
// ByteWeights::from_bytes, weights.rs:265-279).
//
// HBM-bound read of N bytes.  128-bit streaming loads; counts go to shared-memory bins that are privatised
// per LANE rather than per warp: bin b of lane l lives at word b*32 + l, so the 32 lanes of any warp always hit
// 32 different banks whatever the data looks like (a single-symbol input is as fast as uniform noise), and all
// warps of the CTA share the one 32 KB table through native shared-memory atomics.  Per-CTA u32 partials are
// folded into the global u64 bins with one atomic per non-empty bin.
//
// Algorithmic bytes: N read, 2 KiB written.
#pragma once

#include "hb_common.cuh"

namespace hb {

constexpr int kHistThreads = 512;
constexpr int kHistUnroll = 4;            // 128-bit loads in flight per thread

__device__ __forceinline__ void hist_red(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(1u) : "memory");
}

// one byte -> one shared-memory increment in 3 instructions: PRMT isolates the byte, IMAD forms lane_base + byte * 128,
// then the atomic.  (shift + mask + add + atomic = 4 before: 20 % fewer instructions, same 0.188 ms per GiB -- the
// kernel is bound by the shared-atomic rate, one per input byte, not by issue slots)
__device__ __forceinline__ void hist_word(uint32_t lane_base, uint32_t w) {
#ifdef HB_HIST_NO_PRMT
    hist_red(lane_base + ((w & 0xFFu) << 7));
    hist_red(lane_base + (((w >> 8) & 0xFFu) << 7));
    hist_red(lane_base + (((w >> 16) & 0xFFu) << 7));
    hist_red(lane_base + ((w >> 24) << 7));
#else
    hist_red(lane_base + (__byte_perm(w, 0u, 0x4440) << 7));
    hist_red(lane_base + (__byte_perm(w, 0u, 0x4441) << 7));
    hist_red(lane_base + (__byte_perm(w, 0u, 0x4442) << 7));
    hist_red(lane_base + (__byte_perm(w, 0u, 0x4443) << 7));
#endif
}

// data: any alignment.  hist: 256 x u64, zeroed by the caller (cudaMemsetAsync) before the launch.
__global__ void __launch_bounds__(kHistThreads)
hist_lane_columns_kernel(const uint8_t *__restrict__ data, size_t n, unsigned long long *__restrict__ hist) {
    __shared__ uint32_t cols[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += kHistThreads) cols[i] = 0;
    __syncthreads();

    const uint32_t lane_base = static_cast<uint32_t>(__cvta_generic_to_shared(cols)) + (lane_id() << 2);

    // split into unaligned head, 16-byte vectors, tail
    const uintptr_t addr = reinterpret_cast<uintptr_t>(data);
    size_t head = (16 - (addr & 15)) & 15;
    if (head > n) head = n;
    const size_t n_vec = (n - head) / 16;
    const size_t tail_begin = head + n_vec * 16;
    const uint4 *vec = reinterpret_cast<const uint4 *>(data + head);

    const size_t stride = static_cast<size_t>(gridDim.x) * kHistThreads;
    size_t i = static_cast<size_t>(blockIdx.x) * kHistThreads + threadIdx.x;
    // main loop: kHistUnroll independent loads, then the atomics
    for (; i + (kHistUnroll - 1) * stride < n_vec; i += kHistUnroll * stride) {
        uint4 v[kHistUnroll];
#pragma unroll
        for (int u = 0; u < kHistUnroll; u++) v[u] = ld_stream_u4(vec + i + u * stride);
#pragma unroll
        for (int u = 0; u < kHistUnroll; u++) {
            hist_word(lane_base, v[u].x);
            hist_word(lane_base, v[u].y);
            hist_word(lane_base, v[u].z);
            hist_word(lane_base, v[u].w);
        }
    }
    for (; i < n_vec; i += stride) {
        uint4 v = ld_stream_u4(vec + i);
        hist_word(lane_base, v.x);
        hist_word(lane_base, v.y);
        hist_word(lane_base, v.z);
        hist_word(lane_base, v.w);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) hist_red(lane_base + (static_cast<uint32_t>(data[threadIdx.x]) << 7));
        const size_t n_tail = n - tail_begin;
        if (threadIdx.x < n_tail) hist_red(lane_base + (static_cast<uint32_t>(data[tail_begin + threadIdx.x]) << 7));
    }
    __syncthreads();

    // fold the 32 lane columns of each bin (rotated start -> conflict-free) and publish
    if (threadIdx.x < 256) {
        const uint32_t b = threadIdx.x;
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) sum += cols[b * 32 + ((j + b) & 31)];
        if (sum) atomicAdd(hist + b, static_cast<unsigned long long>(sum));
    }
}

// Region variant used by compress(): the input is cut into `n_regions` contiguous regions (one per encoder CTA) and,
// besides the global 256-bin result, every region gets its own 256 x u32 histogram.  The encoder turns those into the
// exact bit offset of every region (sum_b region_hist[r][b] * len[b]) without reading the input a second time and
// without any inter-CTA look-back.  `data` must be 16-byte aligned; region_bytes is a multiple of 16.
// grid = n_regions * ctas_per_region.
__global__ void __launch_bounds__(kHistThreads)
hist_regions_kernel(const uint8_t *__restrict__ data, size_t n, size_t region_bytes, int ctas_per_region,
                    unsigned long long *__restrict__ hist, uint32_t *__restrict__ region_hist) {
    __shared__ uint32_t cols[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += kHistThreads) cols[i] = 0;
    __syncthreads();
    const uint32_t lane_base = static_cast<uint32_t>(__cvta_generic_to_shared(cols)) + (lane_id() << 2);

    const uint32_t region = blockIdx.x / ctas_per_region, part = blockIdx.x % ctas_per_region;
    const size_t begin = static_cast<size_t>(region) * region_bytes;
    if (begin < n) {
        const size_t end = min(n, begin + region_bytes);
        const size_t stride = static_cast<size_t>(ctas_per_region) * kHistThreads;
        const size_t n_vec = (end - begin) / 16;
        const uint4 *vec = reinterpret_cast<const uint4 *>(data + begin);
        size_t i = static_cast<size_t>(part) * kHistThreads + threadIdx.x;
        for (; i + (kHistUnroll - 1) * stride < n_vec; i += kHistUnroll * stride) {
            uint4 v[kHistUnroll];
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) v[u] = ld_stream_u4(vec + i + u * stride);
#pragma unroll
            for (int u = 0; u < kHistUnroll; u++) {
                hist_word(lane_base, v[u].x);
                hist_word(lane_base, v[u].y);
                hist_word(lane_base, v[u].z);
                hist_word(lane_base, v[u].w);
            }
        }
        for (; i < n_vec; i += stride) {
            uint4 v = ld_stream_u4(vec + i);
            hist_word(lane_base, v.x);
            hist_word(lane_base, v.y);
            hist_word(lane_base, v.z);
            hist_word(lane_base, v.w);
        }
        const size_t tail_begin = begin + n_vec * 16;          // < 16 bytes, only in the last region
        if (part == 0 && tail_begin + threadIdx.x < end)
            hist_red(lane_base + (static_cast<uint32_t>(data[tail_begin + threadIdx.x]) << 7));
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        const uint32_t b = threadIdx.x;
        uint32_t sum = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) sum += cols[b * 32 + ((j + b) & 31)];
        if (sum) {
            atomicAdd(hist + b, static_cast<unsigned long long>(sum));
            atomicAdd(region_hist + region * 256 + b, sum);
        }
    }
}


// region_hist[r][b] = sum of the 32 sub-region histograms of region r (only the region encoder needs it: codes > 15 bits)
__global__ void __launch_bounds__(256)
hist_fold_regions_kernel(const uint32_t *__restrict__ sub_hist, uint32_t n_sub, uint32_t *__restrict__ region_hist) {
    const uint32_t r = blockIdx.x, b = threadIdx.x;
    uint32_t sum = 0;
    for (uint32_t s = r * 32; s < min(n_sub, r * 32 + 32); s++) sum += sub_hist[static_cast<size_t>(s) * 256 + b];
    region_hist[r * 256 + b] = sum;
}

}  // namespace hb
