// hb_emu.h -- TEST INFRASTRUCTURE.  A small CUDA execution model on the host CPU, so that the product's own kernel sources
// (huff_encoding_b200/csrc/*.cu, *.cuh -- translated by tests/emu/translate.py, never edited) can be compiled with g++ and
// run by the CPU test suite: one OS thread per resident CTA, one cooperative fiber per CUDA thread, real barriers
// (__syncthreads, bar.sync), warp collectives (shuffles, ballots), shared memory as one bounds-checked arena per CTA,
// global memory atomics, mbarrier + bulk copies, guarded device allocations.
//
// It exists to check LOGIC (indexing, halos, barriers, look-back protocol, host-side launch code) where no GPU is at
// hand, and to run the same parity tests against the oracle that `pytest -m gpu` runs on the B200.  It says nothing about
// performance, and the product never loads it: libhuffb200.so is built by nvcc from the same sources, untranslated.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <type_traits>

namespace hb_emu {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
    Dim3() = default;
    Dim3(unsigned x_) : x(x_) {}
    Dim3(int x_) : x(static_cast<unsigned>(x_)) {}
    Dim3(unsigned long x_) : x(static_cast<unsigned>(x_)) {}
    Dim3(unsigned long long x_) : x(static_cast<unsigned>(x_)) {}
    Dim3(long x_) : x(static_cast<unsigned>(x_)) {}
};

struct ThreadCtx {            // what a CUDA thread sees
    Dim3 tid, bid, bdim, gdim;
};
extern thread_local ThreadCtx *g_thread;         // the fiber running on this OS thread

// ---- launch: runs the grid to completion (kernels of a stream run in order; this model is synchronous)
void launch(Dim3 grid, Dim3 block, size_t dyn_smem_bytes, const char *name, std::function<void()> body);

// ---- shared memory: one arena per CTA: [static __shared__ variables | dynamic shared memory]
uint8_t *dyn_smem();                                               // start of the dynamic part
void *static_smem(size_t bytes, size_t align, int site);           // the static variable declared at `site`
uint8_t *smem_base();                                              // arena start = shared-space address 0
size_t smem_size();
void smem_check(uint32_t addr, uint32_t bytes, const char *what);  // aborts on an out-of-arena shared-space access

// ---- synchronisation
void yield();                                                      // spin-wait loops must call this
void barrier(int id, unsigned count);                              // bar.sync id, count  (count 0 = all live threads)
int barrier_or(int pred);                                          // __syncthreads_or
void warp_sync();
uint64_t warp_exchange(uint64_t v, int src_lane);                  // every lane contributes v, reads lane src_lane's
uint32_t warp_ballot(int pred);
uint64_t warp_reduce_or(uint64_t v);
[[noreturn]] void trap(const char *why);

// ---- mbarrier + 1-D bulk copy global -> shared.  The 8 bytes at `bar` hold the phase counter.  HB_EMU_BULK=eager (default):
//      the copy lands when it is issued; lazy: it lands only when somebody waits on its barrier -- the two ends of what the
//      hardware may do, so a read of the destination before the wait (or a write to it after the issue) shows as a difference
void mbar_init(uint32_t bar);
void mbar_wait(uint32_t bar, uint32_t parity);
void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar);

}  // namespace hb_emu
