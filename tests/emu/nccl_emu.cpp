// nccl_emu.cpp -- TEST INFRASTRUCTURE (tests/emu): the five NCCL entry points hb_api.cu resolves with dlopen, for ranks that
// are THREADS of one process (each with its own hb_ctx).  The model library is built to dlopen this file instead of
// libnccl.so.2, so hb_comm_init / hb_compress_shard_dev run their real multi-rank code path on the CPU.
// ncclAllGather is synchronous here: publish the send pointer, wait for all ranks, copy every rank's block, wait again.
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "nccl.h"

namespace {
struct Group {
    int n = 0, joined = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned gen = 0;
    std::vector<const void *> send;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
std::mutex g_mutex;
std::map<std::string, std::shared_ptr<Group>> g_groups;
uint64_t g_next_id = 1;
}  // namespace

struct ncclComm {
    std::shared_ptr<Group> group;
    int rank;
};

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
    std::lock_guard<std::mutex> lk(g_mutex);
    memset(id, 0, sizeof *id);
    const uint64_t v = g_next_id++;
    memcpy(id->internal, "hb_emu", 6);
    memcpy(id->internal + 8, &v, 8);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int n, ncclUniqueId id, int rank) {
    std::shared_ptr<Group> g;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        const std::string key(id.internal, sizeof id.internal);
        auto &slot = g_groups[key];
        if (!slot) { slot = std::make_shared<Group>(); slot->n = n; slot->send.resize(n); }
        g = slot;
    }
    if (g->n != n || rank < 0 || rank >= n) return static_cast<ncclResult_t>(4);
    *comm = new ncclComm{g, rank};
    g->barrier();                                   // like NCCL: returns when every rank has joined
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    delete comm;
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t type, ncclComm_t comm, cudaStream_t) {
    if (type != ncclUint64) return static_cast<ncclResult_t>(4);
    Group &g = *comm->group;
    const size_t bytes = count * 8;
    g.send[comm->rank] = send;
    g.barrier();
    for (int r = 0; r < g.n; r++) {
        uint8_t *dst = static_cast<uint8_t *>(recv) + static_cast<size_t>(r) * bytes;
        if (dst != g.send[r]) memcpy(dst, g.send[r], bytes);        // in place: a rank's own block is already there
    }
    g.barrier();                                    // nobody reuses its send buffer before everyone has read it
    return ncclSuccess;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "error in the NCCL model"; }

}  // extern "C"
