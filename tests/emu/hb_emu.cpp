// hb_emu.cpp -- TEST INFRASTRUCTURE: the CPU execution model behind tests/emu/include/hb_emu.h (see there).
#include "hb_emu.h"

#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include "cuda_runtime.h"

#if !defined(__x86_64__)
#error "tests/emu: the fiber switch below is written for x86-64"
#endif

// void hb_emu_switch(void **save_sp, void *load_sp): callee-saved registers + stack pointer
extern "C" void hb_emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl hb_emu_switch
.type hb_emu_switch,@function
hb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size hb_emu_switch,.-hb_emu_switch
)");

namespace hb_emu {

thread_local ThreadCtx *g_thread = nullptr;

namespace {

constexpr size_t kStaticSmem = 48 * 1024;          // room for the static __shared__ variables, below the dynamic part
constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;
constexpr int kBarriers = 16;

struct Bar {
    unsigned arrived = 0, gen = 0;
    int acc = 0, result = 0;
};
struct Warp {
    unsigned arrived = 0, gen = 0, live = 32;
    uint64_t slot[32];
    uint64_t result = 0;
};
struct Fiber {
    ThreadCtx t;
    void *sp = nullptr;
    bool done = false;
};
struct PendingCopy { uint32_t dst; const void *src; uint32_t bytes, bar; };
struct Cta {                               // per worker thread, reused CTA after CTA
    std::vector<PendingCopy> pending;      // bulk copies issued but not yet landed (HB_EMU_BULK=lazy)
    std::vector<Fiber> fibers;
    std::vector<uint8_t *> stacks;         // kMaxThreads guarded stacks, allocated once
    unsigned n = 0, live = 0, cur = 0;
    Bar bars[kBarriers];
    Warp warps[kMaxThreads / 32];
    uint8_t *arena = nullptr;
    size_t arena_size = 0, static_top = 0;
    std::map<int, size_t> static_sites;
    void *main_sp = nullptr;
    const std::function<void()> *body = nullptr;
    const char *name = "";
    std::chrono::steady_clock::time_point last_progress;
    uint64_t spins = 0;
};
thread_local Cta *g_cta = nullptr;

double stall_limit_s() {
    static const double v = [] { const char *e = getenv("HB_EMU_STALL_S"); return e ? atof(e) : 120.0; }();
    return v;
}

void progress(Cta *c) {
    c->spins = 0;
}

[[noreturn]] void die(const char *fmt, const char *a = "", const char *b = "") {
    fprintf(stderr, "hb_emu: ");
    fprintf(stderr, fmt, a, b);
    if (g_cta && g_thread)
        fprintf(stderr, " [kernel %s, block %u, thread %u]", g_cta->name, g_thread->bid.x, g_thread->tid.x);
    fprintf(stderr, "\n");
    fflush(stderr);
    abort();
}

void switch_to(Cta *c, unsigned next) {
    const unsigned me = c->cur;
    c->cur = next;
    g_thread = &c->fibers[next].t;
    hb_emu_switch(&c->fibers[me].sp, c->fibers[next].sp);
}

// HB_EMU_ORDER: the order in which the threads of a CTA get their turns.  Code that is correct under the CUDA memory
// model gives the same result under every order; code that leans on an order (a missing barrier or __syncwarp) does not.
//   up (default): round robin, ascending thread index   down: descending   random[:seed]: a random runnable thread
int order_mode() {
    static const int m = [] {
        const char *e = getenv("HB_EMU_ORDER");
        if (!e || !strncmp(e, "up", 2)) return 0;
        if (!strncmp(e, "down", 4)) return 1;
        if (!strncmp(e, "random", 6)) return 2;
        return 0;
    }();
    return m;
}
uint64_t &rng_state() {
    static thread_local uint64_t s = [] {
        const char *e = getenv("HB_EMU_ORDER");
        const char *colon = e ? strchr(e, ':') : nullptr;
        return colon ? strtoull(colon + 1, nullptr, 10) * 2654435761ull + 1 : 88172645463325252ull;
    }();
    return s;
}
unsigned rnd(unsigned n) {
    uint64_t &x = rng_state();
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    return static_cast<unsigned>((x >> 11) % n);
}

// next runnable fiber after `from`; n if there is none
unsigned next_runnable(Cta *c, unsigned from) {
    const int mode = order_mode();
    if (mode == 2) from = rnd(c->n);
    for (unsigned k = 1; k <= c->n; k++) {
        const unsigned i = mode == 1 ? (from + c->n - k) % c->n : (from + k) % c->n;
        if (!c->fibers[i].done) return i;
    }
    return c->n;
}

void fiber_exit(Cta *c);

void fiber_main() {
    Cta *c = g_cta;
    (*c->body)();
    fiber_exit(c);
}

void complete_barriers_after_exit(Cta *c) {
    // hardware counts exited threads as arrived at __syncthreads (barrier 0)
    Bar &b = c->bars[0];
    if (c->live && b.arrived && b.arrived >= c->live) {
        b.arrived = 0;
        b.result = b.acc;
        b.acc = 0;
        b.gen++;
    }
}

void fiber_exit(Cta *c) {
    Fiber &f = c->fibers[c->cur];
    f.done = true;
    c->live--;
    Warp &w = c->warps[f.t.tid.x >> 5];
    w.live--;
    if (w.live && w.arrived && w.arrived >= w.live) die("a lane exited while its warp waits in a collective");
    complete_barriers_after_exit(c);
    progress(c);
    const unsigned nx = next_runnable(c, c->cur);
    if (nx == c->n) {
        void *dummy;
        g_thread = nullptr;
        hb_emu_switch(&dummy, c->main_sp);
    } else {
        switch_to(c, nx);
    }
    die("a finished fiber was resumed");
}

void run_cta(Cta *c, Dim3 bid, Dim3 block, Dim3 grid, size_t dyn_bytes, const char *name, const std::function<void()> &body) {
    const unsigned n = block.x * block.y * block.z;
    if (n == 0 || n > kMaxThreads) die("block size out of range", "", "");
    if (c->stacks.empty()) {
        c->stacks.resize(kMaxThreads);
        const size_t page = 4096;
        for (int i = 0; i < kMaxThreads; i++) {
            uint8_t *m = static_cast<uint8_t *>(mmap(nullptr, kStackBytes + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
            if (m == MAP_FAILED) die("mmap of a fiber stack failed");
            mprotect(m, page, PROT_NONE);                       // stack overflow -> fault, not silent corruption
            c->stacks[i] = m + page;
        }
        c->fibers.resize(kMaxThreads);
    }
    const size_t need = kStaticSmem + dyn_bytes;
    if (need > c->arena_size) {
        free(c->arena);
        if (posix_memalign(reinterpret_cast<void **>(&c->arena), 1024, need)) die("shared-memory arena allocation failed");
        c->arena_size = need;
    }
    memset(c->arena, 0xA5, need);                               // shared memory starts out undefined
    c->static_top = 0;
    c->static_sites.clear();
    c->pending.clear();
    c->n = c->live = n;
    c->body = &body;
    c->name = name;
    for (auto &b : c->bars) b = Bar();
    for (unsigned w = 0; w < (n + 31) / 32; w++) {
        c->warps[w] = Warp();
        c->warps[w].live = (w + 1) * 32 <= n ? 32 : n - w * 32;
    }
    for (unsigned i = 0; i < n; i++) {
        Fiber &f = c->fibers[i];
        f.done = false;
        f.t.tid = Dim3(i);                                      // 1-D blocks only (all the product's kernels)
        f.t.bid = bid;
        f.t.bdim = block;
        f.t.gdim = grid;
        // initial frame: six callee-saved registers, then fiber_main as the return address; rsp % 16 == 8 on entry
        uintptr_t top = (reinterpret_cast<uintptr_t>(c->stacks[i]) + kStackBytes) & ~static_cast<uintptr_t>(15);
        void **sp = reinterpret_cast<void **>(top);
        *--sp = nullptr;                                        // fake return address of fiber_main
        *--sp = reinterpret_cast<void *>(&fiber_main);
        for (int k = 0; k < 6; k++) *--sp = nullptr;
        f.sp = sp;
    }
    c->spins = 0;
    c->last_progress = std::chrono::steady_clock::now();
    c->cur = order_mode() == 1 ? n - 1 : (order_mode() == 2 ? rnd(n) : 0);
    g_thread = &c->fibers[c->cur].t;
    hb_emu_switch(&c->main_sp, c->fibers[c->cur].sp);
    g_thread = nullptr;
    if (c->live) die("kernel returned to the scheduler with live threads");
}

}  // namespace

void yield() {
    Cta *c = g_cta;
    if (++c->spins > (1u << 22)) {                              // rarely: look at the clock
        const auto now = std::chrono::steady_clock::now();
        if (c->spins == (1u << 22) + 1) c->last_progress = now;
        else if (std::chrono::duration<double>(now - c->last_progress).count() > stall_limit_s())
            die("no progress for too long: deadlock (barrier count mismatch, or a spin-wait that can never end)");
        if ((c->spins & 0xFFFFF) == 0) std::this_thread::yield();
    }
    const unsigned nx = next_runnable(c, c->cur);
    if (nx == c->n || nx == c->cur) return;                     // nobody else: the caller re-tests its condition
    switch_to(c, nx);
}

void barrier(int id, unsigned count) {
    Cta *c = g_cta;
    if (id < 0 || id >= kBarriers) die("barrier id out of range");
    Bar &b = c->bars[id];
    const unsigned gen = b.gen;
    const unsigned want = count ? count : c->live;
    if (++b.arrived >= want) {
        if (b.arrived > want) die("more threads arrived at a barrier than it expects");
        b.arrived = 0;
        b.result = b.acc;
        b.acc = 0;
        b.gen++;
        progress(c);
        return;
    }
    while (b.gen == gen) yield();
}

int barrier_or(int pred) {
    Cta *c = g_cta;
    Bar &b = c->bars[0];
    b.acc |= pred ? 1 : 0;
    barrier(0, 0);
    return b.result;     // stable until the next barrier 0 completes, which needs this thread too
}

static Warp &my_warp() { return g_cta->warps[g_thread->tid.x >> 5]; }

// all live lanes of the warp arrive; the last one runs `fn` (with every slot filled) before releasing the others
template <typename F>
static void warp_rendezvous(Warp &w, F fn) {
    const unsigned gen = w.gen;
    if (++w.arrived >= w.live) {
        fn();
        w.arrived = 0;
        w.gen++;
        progress(g_cta);
        return;
    }
    while (w.gen == gen) yield();
}

void warp_sync() {
    warp_rendezvous(my_warp(), [] {});
}

uint64_t warp_exchange(uint64_t v, int src_lane) {
    Warp &w = my_warp();
    const int lane = g_thread->tid.x & 31;
    w.slot[lane] = v;
    warp_rendezvous(w, [] {});
    const uint64_t r = w.slot[src_lane & 31];
    warp_rendezvous(w, [] {});                                   // nobody overwrites a slot before everyone has read
    return r;
}

uint32_t warp_ballot(int pred) {
    Warp &w = my_warp();
    const int lane = g_thread->tid.x & 31;
    w.slot[lane] = pred ? 1 : 0;
    const unsigned base_tid = g_thread->tid.x & ~31u;
    Cta *c = g_cta;
    warp_rendezvous(w, [&] {
        uint64_t m = 0;
        for (int l = 0; l < 32; l++)
            if (base_tid + l < c->n && !c->fibers[base_tid + l].done && w.slot[l]) m |= 1ull << l;
        w.result = m;
    });
    const uint32_t r = static_cast<uint32_t>(w.result);
    warp_rendezvous(w, [] {});
    return r;
}

uint64_t warp_reduce_or(uint64_t v) {
    Warp &w = my_warp();
    const int lane = g_thread->tid.x & 31;
    w.slot[lane] = v;
    const unsigned base_tid = g_thread->tid.x & ~31u;
    Cta *c = g_cta;
    warp_rendezvous(w, [&] {
        uint64_t m = 0;
        for (int l = 0; l < 32; l++)
            if (base_tid + l < c->n && !c->fibers[base_tid + l].done) m |= w.slot[l];
        w.result = m;
    });
    const uint64_t r = w.result;
    warp_rendezvous(w, [] {});
    return r;
}

void trap(const char *why) { die("trap: %s", why); }

static bool bulk_lazy() {
    static const bool v = [] { const char *e = getenv("HB_EMU_BULK"); return e && !strncmp(e, "lazy", 4); }();
    return v;
}
static void land(Cta *c, const PendingCopy &p) {
    memcpy(c->arena + p.dst, p.src, p.bytes);
    uint64_t ph;
    memcpy(&ph, c->arena + p.bar, 8);
    ph++;                                                       // complete_tx of all expected bytes: the phase ends
    memcpy(c->arena + p.bar, &ph, 8);
}
void mbar_init(uint32_t bar) {
    smem_check(bar, 8, "mbarrier.init");
    const uint64_t z = 0;
    memcpy(g_cta->arena + bar, &z, 8);
}
void bulk_copy(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    Cta *c = g_cta;
    if ((dst & 15u) || (reinterpret_cast<uintptr_t>(src) & 15) || (bytes & 15u)) trap("cp.async.bulk: addresses and size must be multiples of 16");
    smem_check(bar, 8, "cp.async.bulk mbarrier");
    if (static_cast<size_t>(dst) + bytes > c->arena_size) trap("cp.async.bulk: destination beyond the CTA's shared memory");
    const PendingCopy p{dst, src, bytes, bar};
    if (bulk_lazy()) c->pending.push_back(p);
    else land(c, p);
}
void mbar_wait(uint32_t bar, uint32_t parity) {
    Cta *c = g_cta;
    smem_check(bar, 8, "mbarrier.try_wait");
    for (size_t i = 0; i < c->pending.size();) {                // lazy copies land now, at the last possible moment
        if (c->pending[i].bar == bar) { land(c, c->pending[i]); c->pending.erase(c->pending.begin() + i); }
        else i++;
    }
    for (;;) {
        uint64_t ph;
        memcpy(&ph, c->arena + bar, 8);
        if ((ph & 1u) != parity) break;
        yield();
    }
}

uint8_t *smem_base() { return g_cta->arena; }
size_t smem_size() { return g_cta->arena_size; }
uint8_t *dyn_smem() { return g_cta->arena + kStaticSmem; }

void *static_smem(size_t bytes, size_t align, int site) {
    Cta *c = g_cta;
    auto it = c->static_sites.find(site);
    if (it != c->static_sites.end()) return c->arena + it->second;
    size_t off = (c->static_top + align - 1) / align * align;
    if (off + bytes > kStaticSmem) die("static __shared__ variables exceed the model's static area");
    c->static_top = off + bytes;
    c->static_sites[site] = off;
    return c->arena + off;
}

void smem_check(uint32_t addr, uint32_t bytes, const char *what) {
    Cta *c = g_cta;
    if (static_cast<size_t>(addr) + bytes > c->arena_size || (addr & (bytes - 1)))
        die("shared-space access out of the CTA's shared memory or misaligned (%s)", what);
}

void launch(Dim3 grid, Dim3 block, size_t dyn_smem_bytes, const char *name, std::function<void()> body) {
    const unsigned n_ctas = grid.x * grid.y * grid.z;
    if (n_ctas == 0) return;
    static const unsigned max_workers = [] {
        const char *e = getenv("HB_EMU_WORKERS");
        const int v = e ? atoi(e) : 8;
        return static_cast<unsigned>(v < 1 ? 1 : v);
    }();
    const unsigned workers = n_ctas < max_workers ? n_ctas : max_workers;
    // CTAs are handed out in index order to a few resident workers, like a GPU with that many SMs.  One grid at a time
    // (the model is synchronous); the workers' stacks and arenas are allocated once and reused.
    static std::mutex launch_mutex;
    std::lock_guard<std::mutex> lock(launch_mutex);
    static std::vector<Cta *> pool;
    while (pool.size() < workers) pool.push_back(new Cta());
    std::atomic<unsigned> next{0};
    auto work = [&](unsigned w) {
        Cta *cta = pool[w];
        g_cta = cta;
        for (;;) {
            const unsigned i = next.fetch_add(1);
            if (i >= n_ctas) break;
            run_cta(cta, Dim3(i), block, grid, dyn_smem_bytes, name, body);
        }
        g_cta = nullptr;
    };
    std::vector<std::thread> threads;
    for (unsigned w = 1; w < workers; w++) threads.emplace_back(work, w);
    work(0);
    for (auto &t : threads) t.join();
}

}  // namespace hb_emu

// marks this build: huff_encoding_b200/_lib.py refuses to load it outside a test run (HB_EMU=1)
extern "C" int hb_emu_is_model() { return 1; }

// ------------------------------------------------------------------ runtime API
namespace {
struct Alloc { void *map; size_t map_bytes; };
std::mutex g_alloc_mutex;
std::map<void *, Alloc> g_allocs;
}  // namespace

cudaError_t cudaMalloc(void **p, size_t bytes) {
    // the allocation ends at a guard page (rounded up to 256 bytes, the allocation granularity contracts may rely on)
    const size_t page = 4096;
    const size_t rounded = (bytes + 255) / 256 * 256;
    const size_t body = (rounded + page - 1) / page * page;
    uint8_t *m = static_cast<uint8_t *>(mmap(nullptr, body + 2 * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
    mprotect(m, page, PROT_NONE);
    mprotect(m + page + body, page, PROT_NONE);
    uint8_t *user = m + page + body - rounded;
    memset(user, 0xCD, rounded);                                 // device memory starts out undefined
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    g_allocs[user] = Alloc{m, body + 2 * page};
    *p = user;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_alloc_mutex);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return cudaErrorInvalidValue;
    munmap(it->second.map, it->second.map_bytes);
    g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) {
    *p = malloc(bytes ? bytes : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    if (bytes) memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t) {
    if (bytes) memset(dst, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemset(void *dst, int value, size_t bytes) { return cudaMemsetAsync(dst, value, bytes, nullptr); }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = reinterpret_cast<cudaEvent_t>(malloc(8)); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorName(cudaError_t e) { return e == cudaSuccess ? "cudaSuccess" : "cudaError(model)"; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "error in the CPU model of the CUDA runtime"; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int) {
    memset(prop, 0, sizeof *prop);
    snprintf(prop->name, sizeof prop->name, "CPU model of an sm_100a device (tests/emu)");
    prop->major = 10;
    prop->minor = 0;
    const char *e = getenv("HB_EMU_SMS");
    prop->multiProcessorCount = e ? atoi(e) : 148;      // the B200's count: same region, sub-region and grid sizes as on the device
    if (prop->multiProcessorCount < 1) prop->multiProcessorCount = 1;
    return cudaSuccess;
}
