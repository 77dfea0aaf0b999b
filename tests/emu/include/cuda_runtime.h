// cuda_runtime.h -- TEST INFRASTRUCTURE (tests/emu): stands in for the CUDA toolkit's header when the product's kernel
// sources are compiled for the host CPU.  Device builtins map onto hb_emu.h; the runtime API is a synchronous model:
// "device" memory is host memory behind guard pages, streams execute immediately in issue order, events are no-ops.
#pragma once

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "hb_emu.h"

// ------------------------------------------------------------------ language keywords
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

#define threadIdx (hb_emu::g_thread->tid)
#define blockIdx (hb_emu::g_thread->bid)
#define blockDim (hb_emu::g_thread->bdim)
#define gridDim (hb_emu::g_thread->gdim)

// ------------------------------------------------------------------ vector types
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
using dim3 = hb_emu::Dim3;

// ------------------------------------------------------------------ integer intrinsics
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31u;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    s &= 31u;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) {
    if (s >= 32u) return hi;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {      // PTX prmt.b32, default mode
    const uint64_t src = (static_cast<uint64_t>(y) << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned c = (sel >> (4 * i)) & 0xFu;
        unsigned b = static_cast<unsigned>(src >> (8 * (c & 7u))) & 0xFFu;
        if (c & 8u) b = (b & 0x80u) ? 0xFFu : 0u;
        r |= b << (8 * i);
    }
    return r;
}
static inline int __ffs(unsigned x) { return __builtin_ffs(static_cast<int>(x)); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void *p) {
    return static_cast<size_t>(static_cast<const uint8_t *>(p) - hb_emu::smem_base());
}

// CUDA's min/max accept mixed integer types (usual arithmetic conversions)
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> min(A a, B b) {
    using T = std::common_type_t<A, B>;
    return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(a) : static_cast<T>(b);
}
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> max(A a, B b) {
    using T = std::common_type_t<A, B>;
    return static_cast<T>(a) < static_cast<T>(b) ? static_cast<T>(b) : static_cast<T>(a);
}

// ------------------------------------------------------------------ synchronisation and warp collectives
static inline void __syncthreads() { hb_emu::barrier(0, 0); }
static inline int __syncthreads_or(int pred) { return hb_emu::barrier_or(pred); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { hb_emu::warp_sync(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

namespace hb_emu {
template <typename T> static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8 && std::is_trivially_copyable<T>::value, "shuffle of a type wider than 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> static inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
static inline int lane() { return static_cast<int>(g_thread->tid.x & 31u); }
}  // namespace hb_emu

template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    return hb_emu::from_bits<T>(hb_emu::warp_exchange(hb_emu::to_bits(v), src & 31));
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) {
    return hb_emu::from_bits<T>(hb_emu::warp_exchange(hb_emu::to_bits(v), hb_emu::lane() ^ (m & 31)));
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int l = hb_emu::lane();
    return hb_emu::from_bits<T>(hb_emu::warp_exchange(hb_emu::to_bits(v), l >= static_cast<int>(d) ? l - static_cast<int>(d) : l));
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    const int l = hb_emu::lane();
    return hb_emu::from_bits<T>(hb_emu::warp_exchange(hb_emu::to_bits(v), l + static_cast<int>(d) < 32 ? l + static_cast<int>(d) : l));
}
static inline unsigned __ballot_sync(unsigned, int pred) { return hb_emu::warp_ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return hb_emu::warp_ballot(pred) != 0; }
static inline int __all_sync(unsigned, int pred) { return hb_emu::warp_ballot(!pred) == 0; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return static_cast<unsigned>(hb_emu::warp_reduce_or(v)); }

// ------------------------------------------------------------------ atomics (global or shared: the same host memory)
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicMin(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ------------------------------------------------------------------ runtime API (synchronous model)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct hb_emu_stream *cudaStream_t;
typedef struct hb_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};

cudaError_t cudaMalloc(void **p, size_t bytes);
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMallocHost(reinterpret_cast<void **>(p), bytes); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = nullptr);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void *dst, int value, size_t bytes);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaGetLastError();
const char *cudaGetErrorName(cudaError_t e);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int device);
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) {
    *n = 1;
    return cudaSuccess;
}
