// nccl_emu.cpp -- TEST INFRASTRUCTURE (tests/emu): the five NCCL entry points hb_api.cu resolves with dlopen, for ranks
// that are threads of one process OR separate processes on this host (each with its own hb_ctx).  The model library is
// built to dlopen this file instead of libnccl.so.2, so hb_comm_init / hb_compress_shard_dev run their real multi-rank
// code path on the CPU.  A communicator is a POSIX shared-memory segment named after its unique id: arrival counters for a
// sense-reversing barrier and one slot per rank.  ncclAllGather is synchronous here: publish the send block, wait for all
// ranks, copy every rank's block, wait again.
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

#include "nccl.h"

namespace {
constexpr int kMaxRanks = 64;
constexpr size_t kSlotBytes = 4096;                  // >= the largest block the library gathers (256 x u64 = 2 KiB)

struct Shared {
    std::atomic<uint32_t> ready;                     // set by the creator once the header is initialised
    std::atomic<uint32_t> n;
    std::atomic<uint32_t> arrived;
    std::atomic<uint32_t> generation;
    std::atomic<uint32_t> detached;
    uint8_t slots[kMaxRanks][kSlotBytes];
};

void barrier(Shared *s) {
    const uint32_t gen = s->generation.load(std::memory_order_acquire);
    if (s->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == s->n.load()) {
        s->arrived.store(0, std::memory_order_relaxed);
        s->generation.fetch_add(1, std::memory_order_acq_rel);
    } else {
        unsigned spins = 0;
        while (s->generation.load(std::memory_order_acquire) == gen) {
            if (++spins > 200) { sched_yield(); }
            if (spins > 20000) { struct timespec ts = {0, 200000}; nanosleep(&ts, nullptr); }
        }
    }
}

void shm_name(const ncclUniqueId &id, char out[64]) {
    uint64_t a, b;
    memcpy(&a, id.internal + 8, 8);
    memcpy(&b, id.internal + 16, 8);
    snprintf(out, 64, "/hb_emu_nccl_%016llx%016llx", static_cast<unsigned long long>(a), static_cast<unsigned long long>(b));
}
}  // namespace

struct ncclComm {
    Shared *shared;
    int rank, n;
    char name[64];
};

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
    memset(id, 0, sizeof *id);
    memcpy(id->internal, "hb_emu", 6);
    uint64_t r[2] = {0, 0};
    FILE *f = fopen("/dev/urandom", "rb");
    if (f) { if (fread(r, 8, 2, f) != 2) r[0] = 0; fclose(f); }
    if (!r[0]) { r[0] = static_cast<uint64_t>(getpid()) << 32 | static_cast<uint64_t>(time(nullptr)); r[1] = reinterpret_cast<uintptr_t>(id); }
    memcpy(id->internal + 8, r, 16);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int n, ncclUniqueId id, int rank) {
    if (n < 1 || n > kMaxRanks || rank < 0 || rank >= n) return static_cast<ncclResult_t>(4);
    ncclComm *c = new ncclComm();
    c->rank = rank;
    c->n = n;
    shm_name(id, c->name);
    bool creator = true;
    int fd = shm_open(c->name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) {
        creator = false;
        for (int tries = 0; fd < 0 && tries < 20000; tries++) {             // the creator may not have got there yet
            fd = shm_open(c->name, O_RDWR, 0600);
            if (fd < 0) { struct timespec ts = {0, 500000}; nanosleep(&ts, nullptr); }
        }
        if (fd < 0) { delete c; return static_cast<ncclResult_t>(2); }
    }
    if (creator && ftruncate(fd, sizeof(Shared)) != 0) { close(fd); delete c; return static_cast<ncclResult_t>(2); }
    if (!creator) {                                                          // wait until the segment has its size
        struct stat st;
        for (int tries = 0; tries < 20000; tries++) {
            if (fstat(fd, &st) == 0 && static_cast<size_t>(st.st_size) >= sizeof(Shared)) break;
            struct timespec ts = {0, 500000};
            nanosleep(&ts, nullptr);
        }
    }
    void *m = mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { delete c; return static_cast<ncclResult_t>(2); }
    c->shared = static_cast<Shared *>(m);
    if (creator) {
        c->shared->n.store(static_cast<uint32_t>(n));
        c->shared->arrived.store(0);
        c->shared->generation.store(0);
        c->shared->detached.store(0);
        c->shared->ready.store(1, std::memory_order_release);
    } else {
        while (c->shared->ready.load(std::memory_order_acquire) != 1) sched_yield();
        if (c->shared->n.load() != static_cast<uint32_t>(n)) { munmap(m, sizeof(Shared)); delete c; return static_cast<ncclResult_t>(4); }
    }
    barrier(c->shared);                              // like NCCL: returns when every rank has joined
    *comm = c;
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    if (!comm) return ncclSuccess;
    if (comm->shared->detached.fetch_add(1) + 1 == static_cast<uint32_t>(comm->n)) shm_unlink(comm->name);   // last one out
    munmap(comm->shared, sizeof(Shared));
    delete comm;
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t type, ncclComm_t comm, cudaStream_t) {
    const size_t bytes = count * 8;
    if (type != ncclUint64 || bytes > kSlotBytes) return static_cast<ncclResult_t>(4);
    Shared *s = comm->shared;
    memcpy(s->slots[comm->rank], send, bytes);
    barrier(s);
    for (int r = 0; r < comm->n; r++) memcpy(static_cast<uint8_t *>(recv) + static_cast<size_t>(r) * bytes, s->slots[r], bytes);
    barrier(s);                                      // nobody overwrites its slot before everyone has read it
    return ncclSuccess;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "error in the NCCL model"; }

}  // extern "C"
