// hb_api.cu -- C ABI of libhuffb200.so (see include/huffb200.h): context, launch logic, host<->device plumbing.
// No CPU fallback: every compute entry point runs the sm_100a kernels or fails with HB_ERR_CUDA.
#include "../../include/huffb200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "hb_decode.cuh"
#include "hb_decode_fused.cuh"
#include "hb_encode.cuh"
#include "hb_encode_warps.cuh"
#include "hb_fixed.cuh"
#include "hb_hist.cuh"
#include "hb_tables.cuh"

namespace {

thread_local std::string g_last_error;

hb_status cuda_fail(cudaError_t e, const char *what, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "%s failed at hb_api.cu:%d: %s (%s)", what, line, cudaGetErrorName(e), cudaGetErrorString(e));
    g_last_error = buf;
    return HB_ERR_CUDA;
}

#define HB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__);        \
    } while (0)
#define HB_TRY(call)                                 \
    do {                                             \
        hb_status s_ = (call);                       \
        if (s_ != HB_OK) return s_;                  \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;      // elements
    hb_status reserve(size_t n) {
        if (n <= cap) return HB_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc", __LINE__); }
        cap = want;
        return HB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DecResult {           // device -> host after the count pass
    uint64_t total_letters;
    uint64_t entry0;
    uint64_t exit_last;
    uint32_t n_dirty;
    uint32_t pad;
};

}  // namespace

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy the process already uses, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) { g_last_error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return nullptr; }
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) {
        g_last_error = "libnccl.so.2 lacks the expected symbols";
        return nullptr;
    }
    api.lib = h;
    return &api;
}
hb_status nccl_fail(NcclApi *api, ncclResult_t r, const char *what) {
    g_last_error = std::string(what) + " failed: " + (api && api->GetErrorString ? api->GetErrorString(r) : "NCCL error");
    return HB_ERR_CUDA;
}
}  // namespace

struct hb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;

    // histogram
    unsigned long long *d_hist = nullptr;
    int hist_grid = 0;
    int hist_ctas_per_sm = 1;
    // per-region histograms of the last hb_histogram call (consumed by the encoder)
    uint32_t *d_region_hist = nullptr;   // [sm_count][256]
    const uint8_t *region_ptr = nullptr;
    size_t region_n = 0, region_letters = 0;
    bool region_valid = false;
    // per-sub-region (= per encoder warp) cumulative histograms of the same call (hb_encode_warps.cuh)
    uint32_t *d_sub_cum = nullptr;       // [sm_count * 32][256]
    unsigned long long *d_sub_bits = nullptr;   // [sm_count * 32]
    uint32_t *d_enc_err = nullptr;       // != 0: the last encode met a letter without a usable code
    uint32_t n_sub = 0, subs_per_cta = 1;
    size_t sub_letters = 0;
    uint32_t enc_max_len = 0;
    bool enc_warps = true;               // HB_NO_ENCODE_WARPS=1: always the region kernel (A/B, tests)

    // encoder
    hb::EncTable *d_enc_table = nullptr;
    hb_tree enc_tree_cached;
    bool enc_tree_valid = false;
    int enc_chunk = 4;                   // letters per chunk: 4 (codes <= 16 bits), 2 (<= 32), 1 (<= 64)
    unsigned long long *d_total_bits = nullptr;

    // fixed-length fast path (hb_fixed.cuh)
    bool fastpath = true;                // HB_NO_FASTPATH=1 turns it off (profiling the general path on uniform data)
    uint8_t *d_fix_enc = nullptr;        // letter -> L-bit code
    uint8_t *d_fix_dec = nullptr;        // L-bit pattern -> letter
    uint32_t enc_fixed_len = 0, dec_fixed_len = 0;
    int fix_grid = 0;
    bool last_dec_fixed = false;
    const uint8_t *last_fix_src = nullptr;
    uint32_t last_fix_len = 0;

    // decoder
    hb::DecTables *d_dec_tables = nullptr;
    uint32_t *d_emit = nullptr;          // multi-letter emit table of the fused decoder (1 << kEmitBits entries)
    uint8_t *d_code_len = nullptr;       // 256 code lengths (fused decoder: letter-by-letter tails)
    DevBuf<unsigned long long> fused_desc;
    uint32_t *d_fused_ctl = nullptr;     // [0] ticket, then a FusedResult (16-byte aligned)
    hb::FusedResult *h_fused_result = nullptr;   // pinned
    bool fused_enabled = true;           // HB_NO_FUSED=1: always take the two-pass decoder
    int fused_teams_forced = 0;                  // HB_FUSED_TEAMS: cap on the teams per CTA (A/B measurements)
    uint32_t last_fused = 0;             // 1: the last decompress ran the fused kernel, 2: it was refuted and redone two-pass
    uint64_t refuted_codes[4] = {0, 0, 0, 0};    // code sets whose speculation was refuted lately: straight to two-pass
    int refuted_next = 0;
    uint32_t last_fused_slow_chunks = 0;
    hb_tree dec_tree_cached;
    bool dec_tree_valid = false;         // dec_tree_cached / dec_fixed_len describe the last tree seen
    bool dec_tables_valid = false;       // d_dec_tables / d_emit were built for dec_tree_cached
    DevBuf<uint32_t> sub_info, blk_count, blk_local, dirty;
    DevBuf<uint64_t> blk_entry, blk_exit, group_total;
    DecResult *d_dec_result = nullptr;
    DecResult *h_dec_result = nullptr;   // pinned
    uint32_t *d_n_dirty = nullptr;
    int dec_count_grid = 0, dec_write_grid = 0;
    bool spoil_speculation = false;      // HB_DEBUG_SPOIL_SPECULATION=1 (tests): force the cross-CTA repair path
    uint32_t last_repairs = 0;           // chunks repaired by dec_fix_kernel in the last count pass
    hb::DecParams last_dec;
    uint64_t last_dec_total = 0;
    bool last_dec_valid = false;

    // multi-GPU: one rank of a communicator (NCCL, resolved at run time: hb_comm_init)
    ncclComm_t comm = nullptr;
    int comm_ranks = 1, comm_rank = 0;
    unsigned long long *d_gather = nullptr;      // [comm_ranks][256] gathered shard histograms
    uint64_t *h_gather = nullptr;                // pinned copy
    // staging for the host-buffer API; two copy streams + events to overlap H2D / kernel / D2H by chunks
    DevBuf<uint8_t> stage_in, stage_out;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    static constexpr int kPipeEvents = 64;
    cudaEvent_t ev_in[kPipeEvents] = {}, ev_k[kPipeEvents] = {};
    uint64_t *h_hist = nullptr;          // pinned 256 x u64
    unsigned long long *h_total_bits = nullptr;
};

namespace {

// Makes the ctx's device current for the duration of an API call and restores the caller's device afterwards
// (the host application -- e.g. torch -- keeps its own notion of the current device).
struct DeviceScope {
    int prev = -1;
    bool switched = false;
    hb_status enter(hb_ctx *ctx) {
        if (!ctx) return HB_ERR_INVALID_ARG;
        HB_CUDA(cudaGetDevice(&prev));
        if (prev != ctx->device) {
            HB_CUDA(cudaSetDevice(ctx->device));
            switched = true;
        }
        return HB_OK;
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};
#define HB_ENTER(ctx) DeviceScope hb_scope_; HB_TRY(hb_scope_.enter(ctx))

bool same_codes(const hb_tree &a, const hb_tree &b) {
    return std::memcmp(a.has_code, b.has_code, sizeof a.has_code) == 0 &&
           std::memcmp(a.code_len, b.code_len, sizeof a.code_len) == 0 &&
           std::memcmp(a.code, b.code, sizeof a.code) == 0;
}
bool same_nodes(const hb_tree &a, const hb_tree &b) {
    return a.n_nodes == b.n_nodes && a.root == b.root &&
           std::memcmp(a.nodes, b.nodes, sizeof(hb_node) * a.n_nodes) == 0;
}

// L if every code of the tree has length L in {1,2,4,8} (perfect tree or lone root), else 0
uint32_t tree_fixed_len(const hb_tree *t) {
    if (t->min_len != t->max_len) return 0;
    const uint32_t L = t->max_len;
    if (L != 1 && L != 2 && L != 4 && L != 8) return 0;
    const bool lone_root = t->nodes[t->root].left == HB_NO_CHILD;
    if (!lone_root && t->n_leaves != (1u << L)) return 0;
    return L;
}

__global__ void set_u64_kernel(unsigned long long *p, unsigned long long v) { *p = v; }

// ---------------------------------------------------------------- histogram
size_t region_size_for(const hb_ctx *ctx, size_t n) {
    // a region (one CTA of the region encoder) is 32 sub-regions (one per warp of the warp encoder); sub-regions are whole
    // tiles of 1024 letters, and -- once they are large -- whole iterations of the histogram kernel (32 KiB), so that a
    // sub-region never ends in the middle of one
    const size_t subs = static_cast<size_t>(ctx->sm_count) * 32;
    size_t sub = ((n + subs - 1) / subs + 1023) / 1024 * 1024;
    const size_t step = static_cast<size_t>(hb::kHistUnroll) * hb::kHistThreads * 16;
    if (sub >= 2 * step) sub = (sub + step - 1) / step * step;
    return std::max<size_t>(1024, sub) * 32;
}

hb_status launch_hist(hb_ctx *ctx, const uint8_t *d_data, size_t n, unsigned long long *d_hist) {
    HB_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), ctx->stream));
    ctx->region_valid = false;
    if (n == 0) return HB_OK;
    const size_t region = region_size_for(ctx, n);
    if ((reinterpret_cast<uintptr_t>(d_data) & 15) == 0 && region < (static_cast<size_t>(1) << 32)) {
        // sub-region variant: global bins + one 256 x u32 histogram per SUB-REGION (one per encoder warp; a region of the
        // region encoder is 32 of them).  One CTA per sub-region.
        const size_t sub = region / 32;
        const uint32_t n_sub = static_cast<uint32_t>((n + sub - 1) / sub);
        HB_CUDA(cudaMemsetAsync(ctx->d_sub_cum, 0, static_cast<size_t>(n_sub) * 256 * sizeof(uint32_t), ctx->stream));
        hb::hist_regions_kernel<<<n_sub, hb::kHistThreads, 0, ctx->stream>>>(d_data, n, sub, 1, d_hist, ctx->d_sub_cum);
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
        ctx->region_ptr = d_data;
        ctx->region_n = n;
        ctx->region_letters = region;
        ctx->region_valid = true;
        ctx->n_sub = n_sub;
        ctx->sub_letters = sub;
        return HB_OK;
    }
    // any alignment: per-CTA partials are u32, keep every launch below 2^32 bytes per CTA
    const size_t max_piece = static_cast<size_t>(1) << 36;
    for (size_t off = 0; off < n; off += max_piece) {
        const size_t len = std::min(max_piece, n - off);
        const size_t vecs = len / 16 + 1;
        int grid = static_cast<int>(std::min<size_t>(ctx->hist_grid, (vecs + hb::kHistThreads - 1) / hb::kHistThreads));
        if (grid < 1) grid = 1;
        hb::hist_lane_columns_kernel<<<grid, hb::kHistThreads, 0, ctx->stream>>>(d_data + off, len, d_hist);
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
    }
    return HB_OK;
}

// ---------------------------------------------------------------- encoder
hb_status upload_enc_table(hb_ctx *ctx, const hb_tree *tree) {
    if (ctx->enc_tree_valid && same_codes(ctx->enc_tree_cached, *tree)) return HB_OK;
    // the code list travels as a kernel parameter; the table is built on the device (hb_tables.cuh)
    hb::CodesParam cp;
    uint32_t max_len = 0;
    for (int b = 0; b < 256; b++) {
        uint32_t len = tree->has_code[b] ? tree->code_len[b] : 0;
        if (len > HB_MAX_ENCODE_BITS) len = 0;             // such letters are rejected per input (check_encodable)
        cp.len[b] = static_cast<uint8_t>(len);
        cp.code[b] = len ? tree->code[b] : 0;
        max_len = std::max(max_len, len);
    }
    hb::enc_table_kernel<<<1, 256, 0, ctx->stream>>>(cp, ctx->d_enc_table);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());
    ctx->enc_tree_cached = *tree;
    ctx->enc_tree_valid = true;
    ctx->enc_chunk = max_len <= 16 ? 4 : (max_len <= 32 ? 2 : 1);
    ctx->enc_max_len = max_len;
    ctx->enc_fixed_len = ctx->fastpath ? tree_fixed_len(tree) : 0;
    if (ctx->enc_fixed_len) {
        uint8_t codes[256];
        for (int b = 0; b < 256; b++) codes[b] = tree->has_code[b] ? static_cast<uint8_t>(tree->code[b]) : 0;
        HB_CUDA(cudaMemcpyAsync(ctx->d_fix_enc, codes, sizeof codes, cudaMemcpyHostToDevice, ctx->stream));
    }
    return HB_OK;
}

template <int S>
hb_status launch_encode_s(hb_ctx *ctx, const uint8_t *d_data, size_t n, uint32_t start_bit, uint8_t *d_out,
                          unsigned long long *d_total_bits) {
    // the encoder needs the per-region histograms of exactly this input
    if (!(ctx->region_valid && ctx->region_ptr == d_data && ctx->region_n == n)) {
        HB_TRY(launch_hist(ctx, d_data, n, ctx->d_hist));
        if (!ctx->region_valid) return HB_ERR_INVALID_ARG;
    }
    ctx->region_valid = false;        // consume-once: a different buffer may land on the same address later
    if (S == 4 && ctx->enc_warps && ctx->enc_max_len <= static_cast<uint32_t>(hb::kEwMaxBits)) {
        // one sub-region per warp: exact bit offsets from the sub-region histograms, then the barrier-free kernel
        HB_CUDA(cudaMemsetAsync(ctx->d_enc_err, 0, sizeof(uint32_t), ctx->stream));
        hb::enc_prepare_kernel<<<(ctx->n_sub * 32 + 255) / 256, 256, 0, ctx->stream>>>(
            ctx->d_sub_cum, ctx->n_sub, ctx->d_enc_table, ctx->d_sub_bits, ctx->d_enc_err);
        hb::encode_warps_kernel<<<(ctx->n_sub + hb::kEwWarps - 1) / hb::kEwWarps, hb::kEwThreads, hb::kEwSmemBytes, ctx->stream>>>(
            d_data, n, ctx->d_enc_table, start_bit, reinterpret_cast<uint32_t *>(d_out), ctx->d_sub_bits, ctx->n_sub,
            ctx->sub_letters, d_total_bits);
        ctx->launches += 2;
        HB_CUDA(cudaGetLastError());
        return HB_OK;
    }
    const int n_regions = static_cast<int>((n + ctx->region_letters - 1) / ctx->region_letters);
    hb::hist_fold_regions_kernel<<<n_regions, 256, 0, ctx->stream>>>(ctx->d_sub_cum, ctx->n_sub, ctx->d_region_hist);
    ctx->launches++;
    hb::encode_regions_kernel<S><<<n_regions, hb::kEncThreads, hb::enc_smem_bytes(S), ctx->stream>>>(
        d_data, n, ctx->d_enc_table, start_bit, reinterpret_cast<uint32_t *>(d_out), ctx->d_region_hist,
        ctx->region_letters, d_total_bits);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

hb_status launch_encode(hb_ctx *ctx, const uint8_t *d_data, size_t n, const hb_tree *tree, uint32_t start_bit,
                        uint8_t *d_out, unsigned long long *d_total_bits, bool all_coded) {
    if (n == 0) {
        if (d_total_bits) HB_CUDA(cudaMemsetAsync(d_total_bits, 0, sizeof(unsigned long long), ctx->stream));
        return HB_OK;
    }
    HB_TRY(upload_enc_table(ctx, tree));
    const uint32_t L = ctx->enc_fixed_len;
    if (L && all_coded && (start_bit % 8) == 0) {
        // fixed-length code set: table translation, no offsets to scan
        uint8_t *dst = d_out + start_bit / 8;
        if (start_bit) HB_CUDA(cudaMemsetAsync(d_out, 0, start_bit / 8, ctx->stream));
        if (L == 8 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(d_data)) & 15) == 0) {
            const size_t blocks = (n / 16 + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid, blocks)));
            hb::fixed8_translate_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(d_data, dst, n, ctx->d_fix_enc);
        } else {
            const size_t out_bytes = (n * L + 7) / 8;
            const size_t blocks = (out_bytes + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid * 4, blocks)));
            hb::fixed_pack_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(d_data, n, dst, out_bytes, L, ctx->d_fix_enc);
        }
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
        if (d_total_bits) {
            set_u64_kernel<<<1, 1, 0, ctx->stream>>>(d_total_bits, static_cast<unsigned long long>(n) * L);
            ctx->launches++;
        }
        return HB_OK;
    }
    switch (ctx->enc_chunk) {
        case 4: return launch_encode_s<4>(ctx, d_data, n, start_bit, d_out, d_total_bits);
        case 2: return launch_encode_s<2>(ctx, d_data, n, start_bit, d_out, d_total_bits);
        default: return launch_encode_s<1>(ctx, d_data, n, start_bit, d_out, d_total_bits);
    }
}

hb_status check_encodable(const uint64_t weights[256], const hb_tree *tree, uint64_t *bits, uint8_t *missing) {
    HB_TRY(hb_stream_bits(weights, tree, bits, missing));
    for (int b = 0; b < 256; b++)
        if (weights[b] && tree->code_len[b] > HB_MAX_ENCODE_BITS) return HB_ERR_CODE_TOO_LONG;
    return HB_OK;
}

// ---------------------------------------------------------------- decoder
// The decoder's view of a tree: remembers the last tree seen (fixed-length fast-path decision + its 256-byte table).
// The big tables are NOT built here: only the general kernels need them (ensure_dec_tables).
hb_status prepare_dec_tree(hb_ctx *ctx, const hb_tree *tree) {
    if (ctx->dec_tree_valid && same_nodes(ctx->dec_tree_cached, *tree)) return HB_OK;
    ctx->dec_tree_cached = *tree;
    ctx->dec_tree_valid = true;
    ctx->dec_tables_valid = false;
    ctx->dec_fixed_len = ctx->fastpath ? tree_fixed_len(tree) : 0;
    if (ctx->dec_fixed_len) {
        const uint32_t L = ctx->dec_fixed_len;
        uint8_t letters[256] = {0};
        for (uint32_t pat = 0; pat < (1u << L); pat++) {                  // walk the pattern down the tree
            uint32_t node = tree->root;
            for (uint32_t k = 0; k < L && tree->nodes[node].left != HB_NO_CHILD; k++)
                node = ((pat >> (L - 1 - k)) & 1) ? tree->nodes[node].right : tree->nodes[node].left;
            letters[pat] = tree->nodes[node].letter;
        }
        HB_CUDA(cudaMemcpyAsync(ctx->d_fix_dec, letters, sizeof letters, cudaMemcpyHostToDevice, ctx->stream));
    }
    return HB_OK;
}

// Tables of the general decoder kernels for the tree of the last prepare_dec_tree call: one ~10 us kernel on the ctx
// stream (hb_tables.cuh); the tree is its parameter.  No host-side table construction, no upload, no synchronisation.
hb_status ensure_dec_tables(hb_ctx *ctx) {
    if (ctx->dec_tables_valid) return HB_OK;
    const hb_tree &tree = ctx->dec_tree_cached;
    hb::TreeParam tp;
    std::memset(&tp, 0, sizeof tp);
    const uint32_t n = std::min<uint32_t>(tree.n_nodes, HB_MAX_NODES);
    for (uint32_t i = 0; i < n; i++) {
        const hb_node &nd = tree.nodes[i];
        tp.nodes[i] = nd.left == HB_NO_CHILD ? (0xFFFFu | (static_cast<uint32_t>(nd.letter) << 16))
                                             : (static_cast<uint32_t>(nd.left) | (static_cast<uint32_t>(nd.right) << 16));
    }
    tp.root = tree.root;
    tp.n_nodes = n;
    tp.emit_bits = tree.max_len <= static_cast<uint32_t>(hb::kEmitBits) ? hb::kEmitBits
                 : (tree.max_len <= static_cast<uint32_t>(hb::kEmitBitsWide) ? hb::kEmitBitsWide : 0);
    for (int b = 0; b < 256; b++) tp.code_len[b] = tree.has_code[b] ? static_cast<uint8_t>(std::min<uint32_t>(tree.code_len[b], 255)) : 0;
    hb::dec_tables_kernel<<<1, hb::kTabThreads, 0, ctx->stream>>>(tp, ctx->d_dec_tables, ctx->d_emit, ctx->d_code_len, hb::kCntBits);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());
    ctx->dec_tables_valid = true;
    return HB_OK;
}

__global__ void dec_collect_kernel(hb::DecParams p, const uint64_t *grand_total, const uint32_t *n_dirty, DecResult *r) {
    r->total_letters = *grand_total;
    r->entry0 = p.blk_entry[0];
    r->exit_last = p.blk_exit[p.n_blocks - 1];
    r->n_dirty = *n_dirty;
}

hb_status run_count_pass(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin, uint64_t own_end,
                         int64_t entry_bit, uint64_t stream_bit0, const hb_tree *tree, hb_shard_info *info) {
    ctx->last_dec_valid = false;
    if (reinterpret_cast<uintptr_t>(d_buf) & 3) return HB_ERR_INVALID_ARG;
    if (own_end > avail_bits) own_end = avail_bits;
    if (own_begin >= own_end) {
        // an empty range passes the chain through: with a known entry (the first code-word start at or after own_begin) that
        // start is also the first one at or after own_end; without one there is nothing to report but the range itself
        info->entry_bit = entry_bit;
        info->exit_bit = (entry_bit >= 0 && static_cast<uint64_t>(entry_bit) > own_begin) ? static_cast<uint64_t>(entry_bit) : own_begin;
        info->n_letters = 0;
        ctx->last_dec_total = 0;
        ctx->last_dec.n_blocks = 0;
        ctx->last_dec_valid = true;
        return HB_OK;
    }
    HB_TRY(prepare_dec_tree(ctx, tree));
    ctx->last_dec_fixed = false;
    if (ctx->dec_fixed_len && entry_bit >= 0 && (entry_bit % 8) == 0 && static_cast<uint64_t>(entry_bit) >= own_begin) {
        // fixed-length code set: every L-th bit after the entry starts a code word; nothing to count on the device
        const uint64_t L = ctx->dec_fixed_len, e = static_cast<uint64_t>(entry_bit);
        const uint64_t by_own = own_end > e ? (own_end - e + L - 1) / L : 0;
        const uint64_t by_avail = avail_bits > e ? (avail_bits - e) / L : 0;
        const uint64_t n = std::min(by_own, by_avail);
        info->entry_bit = entry_bit;
        info->exit_bit = (n == by_own) ? e + n * L : avail_bits;         // same convention as the general path
        info->n_letters = n;
        ctx->last_dec_fixed = true;
        ctx->last_fix_src = d_buf + e / 8;
        ctx->last_fix_len = static_cast<uint32_t>(L);
        ctx->last_dec_total = n;
        ctx->last_dec.n_blocks = 0;
        ctx->last_dec_valid = true;
        return HB_OK;
    }
    HB_TRY(ensure_dec_tables(ctx));
    const uint64_t chunk_bits = static_cast<uint64_t>(hb::kChunkWords) * 32;
    const uint64_t first_block = own_begin / chunk_bits;
    const uint64_t last_block = (own_end - 1) / chunk_bits;
    const uint64_t n_blocks = last_block - first_block + 1;
    if (n_blocks > 0x7FFFFFFFull) return HB_ERR_INVALID_ARG;
    const uint32_t n_groups = static_cast<uint32_t>((n_blocks + hb::kScanGroup - 1) / hb::kScanGroup);
    HB_TRY(ctx->sub_info.reserve(n_blocks * hb::kDecThreads));
    HB_TRY(ctx->blk_count.reserve(n_blocks));
    HB_TRY(ctx->blk_local.reserve(n_blocks));
    HB_TRY(ctx->dirty.reserve(n_blocks));
    HB_TRY(ctx->blk_entry.reserve(n_blocks));
    HB_TRY(ctx->blk_exit.reserve(n_blocks));
    HB_TRY(ctx->group_total.reserve(n_groups + 1));

    hb::DecParams p;
    p.words = reinterpret_cast<const uint32_t *>(d_buf);
    p.n_words_readable = (avail_bits + 31) / 32;
    p.avail_bits = avail_bits;
    p.own_begin = own_begin;
    p.own_end = own_end;
    p.entry_bit = entry_bit;
    p.stream_bit0 = stream_bit0;
    p.len_gcd = tree->len_gcd ? tree->len_gcd : 1;
    p.fixed_len = (tree->min_len == tree->max_len) ? tree->max_len : 0;
    p.max_len = tree->max_len ? tree->max_len : 1;
    p.cnt_bits = static_cast<uint32_t>(hb::kCntBits);
    p.spoil_speculation = ctx->spoil_speculation ? 1u : 0u;
    ctx->last_repairs = 0;
    if (tree->nodes[tree->root].left == HB_NO_CHILD) { p.fixed_len = 1; p.len_gcd = 1; }
    p.first_block = static_cast<uint32_t>(first_block);
    p.n_blocks = static_cast<uint32_t>(n_blocks);
    p.sub_info = ctx->sub_info.p;
    p.blk_entry = ctx->blk_entry.p;
    p.blk_exit = ctx->blk_exit.p;
    p.blk_count = ctx->blk_count.p;

    const int grid = static_cast<int>(std::min<uint64_t>(ctx->dec_count_grid, n_blocks));
    const size_t smem_count = hb::dec_smem_count(hb::kCntBits);
    // trees without codes beyond the count table's index take the instance without the zero-entry tests
    if (p.max_len > static_cast<uint32_t>(hb::kCntBits))
        hb::dec_count_kernel<true><<<grid, hb::kDecThreads, smem_count, ctx->stream>>>(p, ctx->d_dec_tables);
    else
        hb::dec_count_kernel<false><<<grid, hb::kDecThreads, smem_count, ctx->stream>>>(p, ctx->d_dec_tables);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());

    uint64_t *d_grand = ctx->group_total.p + n_groups;
    for (int round = 0; round < 2; round++) {
        HB_CUDA(cudaMemsetAsync(ctx->d_n_dirty, 0, sizeof(uint32_t), ctx->stream));
        hb::dec_verify_kernel<<<static_cast<unsigned>((n_blocks + 255) / 256), 256, 0, ctx->stream>>>(p, ctx->dirty.p, ctx->d_n_dirty);
        hb::dec_scan_groups_kernel<<<n_groups, hb::kScanThreads, 0, ctx->stream>>>(ctx->blk_count.p, p.n_blocks, ctx->blk_local.p, ctx->group_total.p);
        hb::dec_scan_totals_kernel<<<1, hb::kScanThreads, 0, ctx->stream>>>(ctx->group_total.p, n_groups, d_grand);
        dec_collect_kernel<<<1, 1, 0, ctx->stream>>>(p, d_grand, ctx->d_n_dirty, ctx->d_dec_result);
        ctx->launches += 4;
        HB_CUDA(cudaGetLastError());
        HB_CUDA(cudaMemcpyAsync(ctx->h_dec_result, ctx->d_dec_result, sizeof(DecResult), cudaMemcpyDeviceToHost, ctx->stream));
        HB_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->h_dec_result->n_dirty == 0) break;
        ctx->last_repairs += ctx->h_dec_result->n_dirty;
        if (round == 1) { g_last_error = "decoder chain repair did not converge"; return HB_ERR_CUDA; }
        // rare: a chunk whose speculative entry was wrong even after a 1024-bit look-back -> serial repair
        hb::dec_fix_kernel<<<1, hb::kDecThreads, smem_count, ctx->stream>>>(p, ctx->d_dec_tables, ctx->dirty.p);
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
    }
    const DecResult &r = *ctx->h_dec_result;
    info->entry_bit = r.entry0 == hb::kEnd64 ? static_cast<int64_t>(avail_bits) : static_cast<int64_t>(r.entry0);
    info->exit_bit = r.exit_last == hb::kEnd64 ? avail_bits : r.exit_last;
    info->n_letters = r.total_letters;
    ctx->last_dec = p;
    ctx->last_dec_total = r.total_letters;
    ctx->last_dec_valid = true;
    return HB_OK;
}

// ---------------------------------------------------------------- fused one-kernel decoder (hb_decode_fused.cuh)
// index width of the emit table for a tree: the regular 12-bit table, or the wide one for codes of 13..14 bits
int fused_emit_bits(const hb_tree *tree) {
    return tree->max_len <= static_cast<uint32_t>(hb::kEmitBits) ? hb::kEmitBits : hb::kEmitBitsWide;
}

// FNV-1a over the code lengths: what decides how fast a code set resynchronises
uint64_t code_set_hash(const hb_tree *tree) {
    uint64_t h = 1469598103934665603ull;
    for (int b = 0; b < 256; b++) {
        const uint64_t v = tree->has_code[b] ? tree->code_len[b] : 0xFFFFu;
        h = (h ^ (v & 0xFF)) * 1099511628211ull;
        h = (h ^ (v >> 8)) * 1099511628211ull;
    }
    return h ? h : 1;
}

// Number of teams per CTA for a tree the fused kernel can serve, 0 if it cannot.
int fused_teams(const hb_ctx *ctx, const hb_tree *tree) {
    if (!ctx->fused_enabled) return 0;
    if (tree->max_len > static_cast<uint32_t>(hb::kEmitBitsWide) || tree->n_leaves < 2) return 0;
    // near-fixed-length code sets (all lengths within one bit) resynchronise too slowly for the one speculative entry per
    // chunk the fused kernel cannot repair: leave them to the two-pass decoder and its repair kernel
    if (tree->max_len - tree->min_len < 2) return 0;
    uint32_t coded = 0;
    for (int b = 0; b < 256; b++) coded += tree->has_code[b] ? 1u : 0u;
    if (coded != tree->n_leaves) return 0;                     // duplicate letters (ByteWeights quirk): two-pass decoder
    const uint64_t h = code_set_hash(tree);
    for (uint64_t r : ctx->refuted_codes)
        if (r == h) return 0;                                  // refuted lately: this code set resynchronises too slowly
    const size_t budget = 232448 - hb::fused_shared_bytes(fused_emit_bits(tree));
    int teams = std::min<int>(hb::kFMaxTeams, static_cast<int>(budget / hb::fused_team_bytes()));
    if (ctx->fused_teams_forced > 0) teams = std::min(teams, ctx->fused_teams_forced);
    return teams;
}

// slowly resynchronising code set: the next streams of this code set skip the speculative attempt
void note_refuted(hb_ctx *ctx, const hb_tree *tree) {
    ctx->refuted_codes[ctx->refuted_next] = code_set_hash(tree);
    ctx->refuted_next = (ctx->refuted_next + 1) % 4;
}

// Runs the fused kernel.  *refuted = true: a chunk's speculative entry was wrong (or the kernel cannot serve this call) and
// the caller must run the two-pass decoder instead; the output buffer then holds garbage.
hb_status run_fused(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin, uint64_t own_end,
                    uint64_t entry_bit, uint64_t stream_bit0, const hb_tree *tree, int teams,
                    uint8_t *d_out, size_t out_cap, hb_shard_info *info, bool *refuted) {
    *refuted = false;
    HB_TRY(ensure_dec_tables(ctx));
    const uint64_t chunk_bits = static_cast<uint64_t>(hb::kFChunkWords) * 32;
    const uint64_t first_chunk = own_begin / chunk_bits;
    const uint64_t n_chunks = (own_end - 1) / chunk_bits - first_chunk + 1;
    if (n_chunks > 0x7FFFFFFFull) return HB_ERR_INVALID_ARG;
    HB_TRY(ctx->fused_desc.reserve(n_chunks));
    hb::FusedResult *d_result = reinterpret_cast<hb::FusedResult *>(ctx->d_fused_ctl + 4);
    HB_CUDA(cudaMemsetAsync(ctx->fused_desc.p, 0, n_chunks * sizeof(unsigned long long), ctx->stream));
    HB_CUDA(cudaMemsetAsync(ctx->d_fused_ctl, 0, 16 + sizeof(hb::FusedResult), ctx->stream));

    hb::FusedParams p;
    p.words = reinterpret_cast<const uint32_t *>(d_buf);
    p.n_words_readable = (avail_bits + 31) / 32;
    p.avail_bits = avail_bits;
    p.own_begin = own_begin;
    p.own_end = own_end;
    p.entry_bit = entry_bit;
    p.stream_bit0 = stream_bit0;
    p.len_gcd = tree->len_gcd ? tree->len_gcd : 1;
    p.fixed_len = (tree->min_len == tree->max_len) ? tree->max_len : 0;
    p.first_chunk = static_cast<uint32_t>(first_chunk);
    p.n_chunks = static_cast<uint32_t>(n_chunks);
    p.spoil_speculation = ctx->spoil_speculation ? 1u : 0u;
    p.k1 = 1;
    p.k4 = 4;
    p.emit = ctx->d_emit;
    p.code_len = ctx->d_code_len;
    p.desc = ctx->fused_desc.p;
    p.ticket = ctx->d_fused_ctl;
    p.result = d_result;
    p.out = d_out;
    p.out_cap = out_cap;
    const int grid = static_cast<int>(std::min<uint64_t>(ctx->sm_count, (n_chunks + teams - 1) / teams));
    const int eb = fused_emit_bits(tree);
    const size_t smem = hb::fused_shared_bytes(eb) + static_cast<size_t>(teams) * hb::fused_team_bytes();
    if (eb == hb::kEmitBits)
        hb::dec_fused_kernel<hb::kEmitBits><<<grid, teams * hb::kFTeam, smem, ctx->stream>>>(p);
    else
        hb::dec_fused_kernel<hb::kEmitBitsWide><<<grid, teams * hb::kFTeam, smem, ctx->stream>>>(p);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());
    HB_CUDA(cudaMemcpyAsync(ctx->h_fused_result, d_result, sizeof(hb::FusedResult), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    const hb::FusedResult &r = *ctx->h_fused_result;
    ctx->last_fused_slow_chunks = r.careful_threads;
    if (r.error) { *refuted = true; return HB_OK; }
    info->entry_bit = r.entry0 == hb::kEnd64 ? static_cast<int64_t>(avail_bits) : static_cast<int64_t>(r.entry0);
    info->exit_bit = r.exit_last == hb::kEnd64 ? avail_bits : r.exit_last;
    info->n_letters = r.total_letters;
    return HB_OK;
}

hb_status run_write_pass(hb_ctx *ctx, uint8_t *d_out, size_t out_cap) {
    if (!ctx->last_dec_valid) return HB_ERR_INVALID_ARG;
    if (ctx->last_dec_total > out_cap) return HB_ERR_CAPACITY;
    if (ctx->last_dec_total == 0) return HB_OK;
    if (ctx->last_dec_fixed) {
        const size_t n = static_cast<size_t>(ctx->last_dec_total);
        const uint32_t L = ctx->last_fix_len;
        if (L == 8 && ((reinterpret_cast<uintptr_t>(ctx->last_fix_src) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0) {
            const size_t blocks = (n / 16 + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid, blocks)));
            hb::fixed8_translate_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(ctx->last_fix_src, d_out, n, ctx->d_fix_dec);
        } else {
            const size_t n_bytes = (n * L + 7) / 8;
            const size_t blocks = (n_bytes + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid * 4, blocks)));
            hb::fixed_unpack_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(ctx->last_fix_src, d_out, n, L, ctx->d_fix_dec);
        }
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
        return HB_OK;
    }
    if (ctx->last_dec.n_blocks == 0) return HB_OK;
    const hb::DecParams &p = ctx->last_dec;
    const int grid = static_cast<int>(std::min<uint32_t>(ctx->dec_write_grid, p.n_blocks));
    hb::dec_write_kernel<<<grid, hb::kDecThreads, hb::kDecSmemWrite, ctx->stream>>>(
        p, ctx->d_dec_tables, ctx->blk_local.p, ctx->group_total.p, d_out, ctx->last_dec_total);
    ctx->launches++;
    HB_CUDA(cudaGetLastError());
    return HB_OK;
}

// decompress of the owned bit range [own_begin, own_end) of a buffer into d_out: the fused one-pass kernel when the
// tree and the call allow it, else (or when its speculation was refuted) the two-pass kernels.  With HB_ERR_CAPACITY
// the count-pass state is kept, so hb_decode_write_dev can still write into a larger buffer.
hb_status decode_range(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin, uint64_t own_end,
                       int64_t entry_bit, uint64_t stream_bit0, const hb_tree *tree, uint8_t *d_out, size_t out_cap,
                       hb_shard_info *info) {
    HB_TRY(prepare_dec_tree(ctx, tree));
    ctx->last_fused = 0;
    if (own_end > avail_bits) own_end = avail_bits;
    int teams = 0;
    const bool fixed_path = ctx->dec_fixed_len && entry_bit >= 0 && (entry_bit % 8) == 0;
    if (!fixed_path && entry_bit >= 0 && static_cast<uint64_t>(entry_bit) >= own_begin && own_begin < own_end && d_out &&
        out_cap && (reinterpret_cast<uintptr_t>(d_buf) & 15) == 0)
        teams = fused_teams(ctx, tree);
    if (teams > 0) {
        bool refuted = false;
        HB_TRY(run_fused(ctx, d_buf, avail_bits, own_begin, own_end, static_cast<uint64_t>(entry_bit), stream_bit0, tree,
                         teams, d_out, out_cap, info, &refuted));
        if (!refuted && info->n_letters <= out_cap) {
            ctx->last_fused = 1;
            ctx->last_dec_valid = false;
            return HB_OK;
        }
        ctx->last_fused = 2;                                   // fall through: exact two-pass decode (with repair)
        if (refuted) note_refuted(ctx, tree);
    }
    info->entry_bit = entry_bit;
    HB_TRY(run_count_pass(ctx, d_buf, avail_bits, own_begin, own_end, entry_bit, stream_bit0, tree, info));
    if (info->n_letters > out_cap) return HB_ERR_CAPACITY;
    if (info->n_letters && !d_out) return HB_ERR_INVALID_ARG;
    return run_write_pass(ctx, d_out, out_cap);
}

}  // namespace

// ================================================================ C ABI
extern "C" {

const char *hb_status_str(int status) {
    switch (status) {
        case HB_OK: return "ok";
        case HB_ERR_EMPTY_WEIGHTS: return "provided empty weights";
        case HB_ERR_MISSING_LETTER: return "letter not found in codes";
        case HB_ERR_EMPTY_COMP: return "provided comp_bytes are empty";
        case HB_ERR_BAD_PADDING: return "padding bits cannot be larger than 7";
        case HB_ERR_CAPACITY: return "buffer too small";
        case HB_ERR_BIN_TOO_SMALL: return "Provided BitVec is too small for an encoded HuffTree";
        case HB_ERR_BIN_TOO_BIG: return "Provided BitVec is too big for an encoded HuffTree";
        case HB_ERR_BYTES_SHORT: return "slice too short";
        case HB_ERR_TREE_LEN: return "stored tree length must be at least 2";
        case HB_ERR_INVALID_TREE: return "invalid tree in slice";
        case HB_ERR_CUDA: return "CUDA error";
        case HB_ERR_INVALID_ARG: return "invalid argument";
        case HB_ERR_CODE_TOO_LONG: return "code longer than 64 bits";
        case HB_ERR_NO_MEM: return "out of host memory";
        case HB_ERR_TREE_NODES: return "tree has more than 513 nodes (more than 257 leaves)";
        default: return "unknown status";
    }
}

const char *hb_last_error(void) { return g_last_error.c_str(); }
int hb_version(void) { return HB_VERSION_MAJOR * 100 + HB_VERSION_MINOR; }

hb_status hb_ctx_create(int device, hb_ctx **out) {
    if (!out) return HB_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    HB_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { g_last_error = "no such CUDA device"; return HB_ERR_CUDA; }
    int prev_dev = -1;
    HB_CUDA(cudaGetDevice(&prev_dev));
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    HB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        g_last_error = std::string("libhuffb200 is built for sm_100a only; device is ") + prop.name;
        return HB_ERR_CUDA;
    }
    hb_ctx *ctx = new (std::nothrow) hb_ctx();
    if (!ctx) return HB_ERR_NO_MEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    hb_status rc = [&]() -> hb_status {
        HB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        HB_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < hb_ctx::kPipeEvents; i++) {
            HB_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            HB_CUDA(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
        }
        HB_CUDA(cudaMalloc(&ctx->d_hist, 256 * sizeof(unsigned long long)));
        HB_CUDA(cudaMalloc(&ctx->d_region_hist, static_cast<size_t>(ctx->sm_count) * 256 * sizeof(uint32_t)));
        HB_CUDA(cudaMalloc(&ctx->d_enc_table, sizeof(hb::EncTable)));
        HB_CUDA(cudaMalloc(&ctx->d_sub_cum, static_cast<size_t>(ctx->sm_count) * 32 * 256 * sizeof(uint32_t)));
        HB_CUDA(cudaMalloc(&ctx->d_sub_bits, static_cast<size_t>(ctx->sm_count) * 32 * sizeof(unsigned long long)));
        HB_CUDA(cudaMalloc(&ctx->d_enc_err, sizeof(uint32_t)));
        HB_CUDA(cudaMemset(ctx->d_enc_err, 0, sizeof(uint32_t)));
        HB_CUDA(cudaFuncSetAttribute(hb::encode_warps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::kEwSmemBytes)));
        { const char *nw = std::getenv("HB_NO_ENCODE_WARPS"); ctx->enc_warps = !(nw && nw[0] == '1'); }
        HB_CUDA(cudaMalloc(&ctx->d_total_bits, sizeof(unsigned long long)));
        HB_CUDA(cudaMalloc(&ctx->d_dec_tables, sizeof(hb::DecTables)));
        HB_CUDA(cudaMalloc(&ctx->d_emit, sizeof(uint32_t) << hb::kEmitBitsWide));
        HB_CUDA(cudaMalloc(&ctx->d_code_len, 256));
        HB_CUDA(cudaMalloc(&ctx->d_fused_ctl, 16 + sizeof(hb::FusedResult)));
        HB_CUDA(cudaMallocHost(&ctx->h_fused_result, sizeof(hb::FusedResult)));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_fused_kernel<hb::kEmitBits>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_fused_kernel<hb::kEmitBitsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        { const char *nf = std::getenv("HB_NO_FUSED"); ctx->fused_enabled = !(nf && nf[0] == '1'); }
        { const char *tf = std::getenv("HB_FUSED_TEAMS"); ctx->fused_teams_forced = tf ? std::atoi(tf) : 0; }
        HB_CUDA(cudaMalloc(&ctx->d_fix_enc, 256));
        HB_CUDA(cudaMalloc(&ctx->d_fix_dec, 256));
        { const char *nf = std::getenv("HB_NO_FASTPATH"); ctx->fastpath = !(nf && nf[0] == '1'); }
        { const char *sp = std::getenv("HB_DEBUG_SPOIL_SPECULATION"); ctx->spoil_speculation = sp && sp[0] == '1'; }
        HB_CUDA(cudaMalloc(&ctx->d_dec_result, sizeof(DecResult)));
        HB_CUDA(cudaMalloc(&ctx->d_n_dirty, sizeof(uint32_t)));
        HB_CUDA(cudaMallocHost(&ctx->h_dec_result, sizeof(DecResult)));
        HB_CUDA(cudaMallocHost(&ctx->h_hist, 256 * sizeof(uint64_t)));
        HB_CUDA(cudaMallocHost(&ctx->h_total_bits, sizeof(unsigned long long)));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::dec_smem_count(hb::kCntBits))));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::dec_smem_count(hb::kCntBits))));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_fix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::dec_smem_count(hb::kCntBits))));
        HB_CUDA(cudaFuncSetAttribute(hb::dec_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::kDecSmemWrite)));
        int occ = 0;
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hb::hist_lane_columns_kernel, hb::kHistThreads, 0));
        ctx->hist_ctas_per_sm = std::max(occ, 1);
        ctx->hist_grid = ctx->sm_count * ctx->hist_ctas_per_sm;
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hb::fixed8_translate_kernel, hb::kFixThreads, 0));
        ctx->fix_grid = ctx->sm_count * std::max(occ, 1);
        HB_CUDA(cudaFuncSetAttribute(hb::encode_regions_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::enc_smem_bytes(4))));
        HB_CUDA(cudaFuncSetAttribute(hb::encode_regions_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::enc_smem_bytes(2))));
        HB_CUDA(cudaFuncSetAttribute(hb::encode_regions_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(hb::enc_smem_bytes(1))));
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hb::dec_count_kernel<true>, hb::kDecThreads, hb::dec_smem_count(hb::kCntBits)));
        ctx->dec_count_grid = ctx->sm_count * std::max(occ, 1);
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hb::dec_write_kernel, hb::kDecThreads, hb::kDecSmemWrite));
        ctx->dec_write_grid = ctx->sm_count * std::max(occ, 1);
        return HB_OK;
    }();
    if (rc != HB_OK) { hb_ctx_destroy(ctx); return rc; }
    *out = ctx;
    return HB_OK;
}

hb_status hb_ctx_destroy(hb_ctx *ctx) {
    if (!ctx) return HB_OK;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) { NcclApi *api = nccl_api(); if (api) api->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    if (ctx->d_gather) cudaFree(ctx->d_gather);
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    cudaFree(ctx->d_hist); cudaFree(ctx->d_region_hist); cudaFree(ctx->d_sub_cum); cudaFree(ctx->d_sub_bits); cudaFree(ctx->d_enc_err); cudaFree(ctx->d_enc_table); cudaFree(ctx->d_total_bits); cudaFree(ctx->d_dec_tables); cudaFree(ctx->d_emit); cudaFree(ctx->d_code_len); cudaFree(ctx->d_fused_ctl); cudaFree(ctx->d_fix_enc);
    if (ctx->h_fused_result) cudaFreeHost(ctx->h_fused_result);
    ctx->fused_desc.release(); cudaFree(ctx->d_fix_dec);
    cudaFree(ctx->d_dec_result); cudaFree(ctx->d_n_dirty);
    if (ctx->h_dec_result) cudaFreeHost(ctx->h_dec_result);
    if (ctx->h_hist) cudaFreeHost(ctx->h_hist);
    if (ctx->h_total_bits) cudaFreeHost(ctx->h_total_bits);
    ctx->sub_info.release(); ctx->blk_count.release(); ctx->blk_local.release(); ctx->dirty.release();
    ctx->blk_entry.release(); ctx->blk_exit.release(); ctx->group_total.release();
    ctx->stage_in.release(); ctx->stage_out.release();
    for (int i = 0; i < hb_ctx::kPipeEvents; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_k[i]) cudaEventDestroy(ctx->ev_k[i]);
    }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HB_OK;
}

hb_status hb_ctx_sync(hb_ctx *ctx) {
    HB_ENTER(ctx);
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

void *hb_ctx_stream(hb_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

hb_status hb_ctx_last_decode_repairs(hb_ctx *ctx, uint32_t *count) {
    if (!ctx || !count) return HB_ERR_INVALID_ARG;
    *count = ctx->last_repairs;
    return HB_OK;
}

hb_status hb_ctx_kernel_launches(hb_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) return HB_ERR_INVALID_ARG;
    *count = ctx->launches;
    return HB_OK;
}

void hb_free(void *p) { std::free(p); }

hb_status hb_host_alloc(size_t bytes, void **p) {
    if (!p) return HB_ERR_INVALID_ARG;
    HB_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    return HB_OK;
}
void hb_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------- device-buffer API
hb_status hb_histogram_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, uint64_t *d_hist256) {
    HB_ENTER(ctx);
    if (!d_hist256 || (n && !d_data)) return HB_ERR_INVALID_ARG;
    return launch_hist(ctx, d_data, n, reinterpret_cast<unsigned long long *>(d_hist256));
}

hb_status hb_encode_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, const hb_tree *tree, uint32_t start_bit,
                           uint8_t *d_out, size_t out_cap, uint64_t *d_total_bits) {
    HB_ENTER(ctx);
    if (!tree || !d_out || (n && !d_data) || start_bit > 31) return HB_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(d_data) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 3)) return HB_ERR_INVALID_ARG;
    (void)out_cap;   // capacity is the caller's contract (exact size comes from hb_stream_bits); checked in the *_u8 path
    bool all_coded = true;
    for (int b = 0; b < 256; b++) all_coded = all_coded && tree->has_code[b];
    return launch_encode(ctx, d_data, n, tree, start_bit, d_out, reinterpret_cast<unsigned long long *>(d_total_bits), all_coded);
}

hb_status hb_compress_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, int order_mode, hb_tree *tree_out,
                             uint8_t *d_out, size_t out_cap, size_t *comp_len, uint8_t *padding_bits) {
    HB_ENTER(ctx);
    if (!tree_out || !d_out || !comp_len || !padding_bits || (n && !d_data)) return HB_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(d_data) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 3)) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_WEIGHTS;                      // comp.rs:354 -> tree_inner.rs:283-285
    HB_TRY(launch_hist(ctx, d_data, n, ctx->d_hist));
    HB_CUDA(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    HB_TRY(hb_tree_from_weights(ctx->h_hist, order_mode, tree_out));
    uint64_t bits = 0;
    uint8_t missing = 0;
    HB_TRY(check_encodable(ctx->h_hist, tree_out, &bits, &missing));
    const size_t need = static_cast<size_t>((bits + 7) / 8);
    if (((need + 3) & ~static_cast<size_t>(3)) > out_cap) { *comp_len = need; return HB_ERR_CAPACITY; }
    HB_TRY(launch_encode(ctx, d_data, n, tree_out, 0, d_out, nullptr, true));
    *comp_len = need;
    *padding_bits = static_cast<uint8_t>((8 - bits % 8) % 8);    // comp.rs:446
    return HB_OK;
}

hb_status hb_decode_count_dev(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin,
                              uint64_t own_end, uint64_t stream_bit0, const hb_tree *tree, hb_shard_info *info) {
    HB_ENTER(ctx);
    if (!d_buf || !tree || !info) return HB_ERR_INVALID_ARG;
    return run_count_pass(ctx, d_buf, avail_bits, own_begin, own_end, info->entry_bit, stream_bit0, tree, info);
}

hb_status hb_decode_write_dev(hb_ctx *ctx, uint8_t *d_out, size_t out_cap) {
    HB_ENTER(ctx);
    if (!d_out && ctx->last_dec_total) return HB_ERR_INVALID_ARG;
    return run_write_pass(ctx, d_out, out_cap);
}

hb_status hb_decompress_u8_dev(hb_ctx *ctx, const uint8_t *d_comp, size_t comp_len, uint8_t padding_bits,
                               const hb_tree *tree, uint8_t *d_out, size_t out_cap, size_t *out_n) {
    HB_ENTER(ctx);
    if (!tree || !out_n) return HB_ERR_INVALID_ARG;
    if (comp_len == 0) return HB_ERR_EMPTY_COMP;                  // comp.rs:56-58
    if (padding_bits > 7) return HB_ERR_BAD_PADDING;              // comp.rs:59-61
    if (!d_comp) return HB_ERR_INVALID_ARG;
    const uint64_t total_bits = static_cast<uint64_t>(comp_len) * 8 - padding_bits;
    hb_shard_info info;
    info.entry_bit = 0;
    info.exit_bit = 0;
    info.n_letters = 0;
    const hb_status st = decode_range(ctx, d_comp, total_bits, 0, total_bits, 0, 0, tree, d_out, out_cap, &info);
    *out_n = static_cast<size_t>(info.n_letters);
    return st;
}

hb_status hb_decode_shard_dev(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin,
                              uint64_t own_end, uint64_t stream_bit0, const hb_tree *tree, hb_shard_info *info,
                              uint8_t *d_out, size_t out_cap) {
    HB_ENTER(ctx);
    if (!d_buf || !tree || !info) return HB_ERR_INVALID_ARG;
    return decode_range(ctx, d_buf, avail_bits, own_begin, own_end, info->entry_bit, stream_bit0, tree, d_out, out_cap, info);
}

hb_status hb_ctx_last_encode_error(hb_ctx *ctx, uint32_t *flag) {
    HB_ENTER(ctx);
    if (!flag) return HB_ERR_INVALID_ARG;
    HB_CUDA(cudaMemcpyAsync(ctx->h_total_bits, ctx->d_enc_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    *flag = *reinterpret_cast<uint32_t *>(ctx->h_total_bits);
    return HB_OK;
}

hb_status hb_ctx_fused_phase_cycles(hb_ctx *ctx, uint64_t out[8]) {
    if (!ctx || !out) return HB_ERR_INVALID_ARG;
    for (int k = 0; k < 8; k++) out[k] = ctx->h_fused_result->phase_cycles[k];
    return HB_OK;
}

hb_status hb_ctx_last_decode_path(hb_ctx *ctx, uint32_t *fused, uint32_t *slow_chunks) {
    if (!ctx) return HB_ERR_INVALID_ARG;
    if (fused) *fused = ctx->last_fused;
    if (slow_chunks) *slow_chunks = ctx->last_fused_slow_chunks;
    return HB_OK;
}

// ---------------------------------------------------------------- multi-GPU inside the library
// NCCL is resolved at run time (dlopen): libhuffb200.so has no link-time dependency on it, and inside a process that
// already carries an NCCL (torch) the same copy is used.
static_assert(HB_COMM_ID_BYTES == sizeof(ncclUniqueId), "HB_COMM_ID_BYTES must match ncclUniqueId");

hb_status hb_comm_get_unique_id(uint8_t id[HB_COMM_ID_BYTES]) {
    if (!id) return HB_ERR_INVALID_ARG;
    NcclApi *api = nccl_api();
    if (!api) return HB_ERR_CUDA;
    ncclUniqueId u;
    const ncclResult_t r = api->GetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail(api, r, "ncclGetUniqueId");
    std::memcpy(id, &u, sizeof u);
    return HB_OK;
}

hb_status hb_comm_init(hb_ctx *ctx, int n_ranks, int rank, const uint8_t id[HB_COMM_ID_BYTES]) {
    HB_ENTER(ctx);
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id) || ctx->comm) return HB_ERR_INVALID_ARG;
    ctx->comm_ranks = n_ranks;
    ctx->comm_rank = rank;
    HB_CUDA(cudaMalloc(&ctx->d_gather, static_cast<size_t>(n_ranks) * 256 * sizeof(unsigned long long)));
    HB_CUDA(cudaMallocHost(&ctx->h_gather, static_cast<size_t>(n_ranks) * 256 * sizeof(uint64_t)));
    if (n_ranks == 1) return HB_OK;
    NcclApi *api = nccl_api();
    if (!api) return HB_ERR_CUDA;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    const ncclResult_t r = api->CommInitRank(&ctx->comm, n_ranks, u, rank);
    if (r != ncclSuccess) { ctx->comm = nullptr; return nccl_fail(api, r, "ncclCommInitRank"); }
    return HB_OK;
}

hb_status hb_comm_finalize(hb_ctx *ctx) {
    HB_ENTER(ctx);
    if (ctx->comm) {
        cudaStreamSynchronize(ctx->stream);
        NcclApi *api = nccl_api();
        if (api) api->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    if (ctx->d_gather) { cudaFree(ctx->d_gather); ctx->d_gather = nullptr; }
    if (ctx->h_gather) { cudaFreeHost(ctx->h_gather); ctx->h_gather = nullptr; }
    ctx->comm_ranks = 1;
    ctx->comm_rank = 0;
    return HB_OK;
}

// compress() of one contiguous shard of a larger input (weights.rs:293-319 is the reference's only data-parallel piece:
// split, count, reduce -- here the reduce is ONE all-gather of the shard histograms on the ctx stream).  Every rank calls
// this with its shard; the concatenation of the shard streams (hb_shard_layout says where each goes; neighbours share at
// most one byte, to be OR-ed) is bit for bit the stream one GPU produces for the concatenated input.
hb_status hb_compress_shard_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, int order_mode, hb_tree *tree_out,
                                uint8_t *d_out, size_t out_cap, hb_shard_layout *layout) {
    HB_ENTER(ctx);
    if (!tree_out || !d_out || !layout || (n && !d_data) || !ctx->d_gather) return HB_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(d_data) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 3)) return HB_ERR_INVALID_ARG;
    const int G = ctx->comm_ranks;
    unsigned long long *mine = ctx->d_gather + static_cast<size_t>(ctx->comm_rank) * 256;
    HB_TRY(launch_hist(ctx, d_data, n, mine));                   // my bins straight into my row of the gather buffer
    if (G > 1) {
        NcclApi *api = nccl_api();
        if (!api || !ctx->comm) return HB_ERR_INVALID_ARG;
        const ncclResult_t r = api->AllGather(mine, ctx->d_gather, 256, ncclUint64, ctx->comm, ctx->stream);
        if (r != ncclSuccess) return nccl_fail(api, r, "ncclAllGather");
    }
    HB_CUDA(cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, static_cast<size_t>(G) * 256 * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));                 // the tree needs the whole histogram
    uint64_t bits[64];
    if (G > 64) return HB_ERR_INVALID_ARG;
    HB_TRY(hb_shard_plan(ctx->h_gather, static_cast<size_t>(G), order_mode, tree_out, bits));   // empty input -> HB_ERR_EMPTY_WEIGHTS
    for (int b = 0; b < 256; b++)
        if (tree_out->has_code[b] && tree_out->code_len[b] > HB_MAX_ENCODE_BITS) return HB_ERR_CODE_TOO_LONG;
    uint64_t offset = 0, total = 0;
    for (int g = 0; g < G; g++) { if (g < ctx->comm_rank) offset += bits[g]; total += bits[g]; }
    layout->bit_offset = offset;
    layout->bits = bits[ctx->comm_rank];
    layout->total_bits = total;
    layout->start_bit = static_cast<uint32_t>(offset % 8);
    layout->padding_bits = static_cast<uint8_t>((8 - total % 8) % 8);
    // an empty shard owns no byte of the stream (not even the one its neighbours share)
    layout->comp_len = layout->bits ? static_cast<size_t>((layout->start_bit + layout->bits + 7) / 8) : 0;
    if (((layout->comp_len + 3) & ~static_cast<size_t>(3)) > out_cap) return HB_ERR_CAPACITY;
    if (n == 0) return HB_OK;
    return launch_encode(ctx, d_data, n, tree_out, layout->start_bit, d_out, nullptr, true);
}

// decompress() of a shard produced by hb_compress_shard_dev: its first code word starts at layout->start_bit, so no
// exchange is needed; one-pass fused decoder when the tree allows it.
hb_status hb_decompress_shard_dev(hb_ctx *ctx, const uint8_t *d_comp, const hb_shard_layout *layout, const hb_tree *tree,
                                  uint8_t *d_out, size_t out_cap, size_t *out_n) {
    HB_ENTER(ctx);
    if (!d_comp || !layout || !tree || !out_n) return HB_ERR_INVALID_ARG;
    *out_n = 0;
    if (layout->bits == 0) return HB_OK;
    const uint64_t begin = layout->start_bit, end = begin + layout->bits;
    hb_shard_info info;
    info.entry_bit = static_cast<int64_t>(begin);
    info.exit_bit = 0;
    info.n_letters = 0;
    const hb_status st = decode_range(ctx, d_comp, end, begin, end, static_cast<int64_t>(begin), layout->bit_offset - begin,
                                      tree, d_out, out_cap, &info);
    *out_n = static_cast<size_t>(info.n_letters);
    return st;
}

// ---------------------------------------------------------------- host-buffer API
hb_status hb_histogram_u8(hb_ctx *ctx, const uint8_t *data, size_t n, uint64_t out[256]) {
    HB_ENTER(ctx);
    if (!out || (n && !data)) return HB_ERR_INVALID_ARG;
    HB_TRY(ctx->stage_in.reserve(n + 16));
    if (n) HB_CUDA(cudaMemcpyAsync(ctx->stage_in.p, data, n, cudaMemcpyHostToDevice, ctx->stream));
    HB_TRY(launch_hist(ctx, ctx->stage_in.p, n, ctx->d_hist));
    HB_CUDA(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::memcpy(out, ctx->h_hist, 256 * sizeof(uint64_t));
    return HB_OK;
}

// dst == nullptr: allocate the result with malloc (returned through *comp_bytes); else write into dst[0..cap)
static hb_status compress_host_common(hb_ctx *ctx, const uint8_t *data, size_t n, const hb_tree *tree, int order_mode,
                                      hb_tree *tree_out, uint8_t **comp_bytes, uint8_t *dst, size_t cap, size_t *comp_len,
                                      uint8_t *padding_bits, uint8_t *missing) {
    if (comp_bytes) *comp_bytes = nullptr;
    *comp_len = 0;
    HB_TRY(ctx->stage_in.reserve(n + 16));
    if (n) HB_CUDA(cudaMemcpyAsync(ctx->stage_in.p, data, n, cudaMemcpyHostToDevice, ctx->stream));
    HB_TRY(launch_hist(ctx, ctx->stage_in.p, n, ctx->d_hist));
    HB_CUDA(cudaMemcpyAsync(ctx->h_hist, ctx->d_hist, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!tree) {
        HB_TRY(hb_tree_from_weights(ctx->h_hist, order_mode, tree_out));   // n == 0 -> HB_ERR_EMPTY_WEIGHTS
        tree = tree_out;
    }
    uint64_t bits = 0;
    uint8_t miss = 0;
    hb_status rc = hb_stream_bits(ctx->h_hist, tree, &bits, &miss);
    if (rc == HB_ERR_MISSING_LETTER) {
        // comp.rs:424-432 reports the first offending letter in INPUT order: find it on the host copy
        for (size_t i = 0; i < n; i++)
            if (!tree->has_code[data[i]]) { miss = data[i]; break; }
        if (missing) *missing = miss;
        return rc;
    }
    HB_TRY(rc);
    for (int b = 0; b < 256; b++)
        if (ctx->h_hist[b] && tree->code_len[b] > HB_MAX_ENCODE_BITS) return HB_ERR_CODE_TOO_LONG;
    if (bits == 0) return HB_ERR_EMPTY_COMP;                      // comp.rs:450 -> :56-58 (n == 0 with a given tree)
    const size_t need = static_cast<size_t>((bits + 7) / 8);
    if (dst && need > cap) { *comp_len = need; return HB_ERR_CAPACITY; }
    HB_TRY(ctx->stage_out.reserve(need + 16));
    HB_TRY(launch_encode(ctx, ctx->stage_in.p, n, tree, 0, ctx->stage_out.p, nullptr, true));
    uint8_t *host = dst ? dst : static_cast<uint8_t *>(std::malloc(need));
    if (!host) return HB_ERR_NO_MEM;
    cudaError_t e = cudaMemcpyAsync(host, ctx->stage_out.p, need, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { if (!dst) std::free(host); return cuda_fail(e, "D2H of the compressed stream", __LINE__); }
    if (comp_bytes) *comp_bytes = host;
    *comp_len = need;
    *padding_bits = static_cast<uint8_t>((8 - bits % 8) % 8);
    return HB_OK;
}

hb_status hb_compress_u8(hb_ctx *ctx, const uint8_t *data, size_t n, int order_mode, hb_tree *tree_out,
                         uint8_t **comp_bytes, size_t *comp_len, uint8_t *padding_bits) {
    HB_ENTER(ctx);
    if (!tree_out || !comp_bytes || !comp_len || !padding_bits || (n && !data)) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_WEIGHTS;
    return compress_host_common(ctx, data, n, nullptr, order_mode, tree_out, comp_bytes, nullptr, 0, comp_len, padding_bits, nullptr);
}

hb_status hb_compress_with_tree_u8(hb_ctx *ctx, const uint8_t *data, size_t n, const hb_tree *tree,
                                   uint8_t **comp_bytes, size_t *comp_len, uint8_t *padding_bits, uint8_t *missing) {
    HB_ENTER(ctx);
    if (!tree || !comp_bytes || !comp_len || !padding_bits || (n && !data)) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_COMP;                         // empty letters -> CompressData::new panics (comp.rs:56-58)
    return compress_host_common(ctx, data, n, tree, 0, nullptr, comp_bytes, nullptr, 0, comp_len, padding_bits, missing);
}

// Fixed-length code set, host buffers: every chunk of the stream decodes independently, so the H2D copy of chunk
// i+1, the translation kernel of chunk i and the D2H copy of chunk i-1 run concurrently (full-duplex PCIe).
static hb_status decompress_host_fixed_pipelined(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint64_t total_bits,
                                                 uint8_t *host, size_t n_letters) {
    const uint32_t L = ctx->dec_fixed_len;
    const size_t per = 8 / L;                                        // letters per stream byte
    const size_t in_bytes = (n_letters * L + 7) / 8;
    (void)comp_len; (void)total_bits;
    HB_TRY(ctx->stage_in.reserve(in_bytes + 64));
    HB_TRY(ctx->stage_out.reserve(n_letters + 64));
    size_t chunk = std::max<size_t>((in_bytes + hb_ctx::kPipeEvents - 1) / hb_ctx::kPipeEvents, static_cast<size_t>(32) << 20);
    chunk = (chunk + 255) / 256 * 256;
    HB_CUDA(cudaStreamSynchronize(ctx->stream));                     // the table upload precedes every kernel below
    int i = 0;
    for (size_t off = 0; off < in_bytes; off += chunk, i++) {
        const size_t len = std::min(chunk, in_bytes - off);
        const size_t letters = std::min(len * per, n_letters - off * per);
        HB_CUDA(cudaMemcpyAsync(ctx->stage_in.p + off, comp + off, len, cudaMemcpyHostToDevice, ctx->s_h2d));
        HB_CUDA(cudaEventRecord(ctx->ev_in[i], ctx->s_h2d));
        HB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[i], 0));
        if (L == 8) {
            const size_t blocks = (letters / 16 + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid, blocks)));
            hb::fixed8_translate_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(ctx->stage_in.p + off, ctx->stage_out.p + off, letters, ctx->d_fix_dec);
        } else {
            const size_t blocks = (len + hb::kFixThreads - 1) / hb::kFixThreads;
            const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>(ctx->fix_grid * 4, blocks)));
            hb::fixed_unpack_kernel<<<grid, hb::kFixThreads, 0, ctx->stream>>>(ctx->stage_in.p + off, ctx->stage_out.p + off * per, letters, L, ctx->d_fix_dec);
        }
        ctx->launches++;
        HB_CUDA(cudaGetLastError());
        HB_CUDA(cudaEventRecord(ctx->ev_k[i], ctx->stream));
        HB_CUDA(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[i], 0));
        HB_CUDA(cudaMemcpyAsync(host + off * per, ctx->stage_out.p + off * per, letters, cudaMemcpyDeviceToHost, ctx->s_d2h));
    }
    HB_CUDA(cudaStreamSynchronize(ctx->s_d2h));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

// General (variable-length) code set, host buffers, long stream: the stream is cut into slabs; the entry of slab k is the
// exit of slab k - 1 (known when its fused kernel has finished), so the H2D copy of slab k + 1, the decode of slab k and
// the D2H copy of the letters of slab k - 1 run concurrently (full-duplex PCIe), like the fixed-length pipeline above.
// *done = false: nothing usable was produced (speculation refuted, or the caller's buffer is too small) -- the caller
// takes the unpipelined path, which handles both exactly.
static hb_status decompress_host_slabs(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint64_t total_bits,
                                       const hb_tree *tree, uint8_t *dst, size_t cap, size_t dev_cap, size_t *out_n,
                                       bool *done) {
    *done = false;
    const int teams = fused_teams(ctx, tree);
    size_t slab = std::max<size_t>((comp_len + hb_ctx::kPipeEvents - 2) / (hb_ctx::kPipeEvents - 1), static_cast<size_t>(32) << 20);
    slab = (slab + 255) / 256 * 256;
    const int n_slabs = static_cast<int>((comp_len + slab - 1) / slab);
    HB_CUDA(cudaStreamSynchronize(ctx->stream));                     // tables of an earlier call are complete
    for (int k = 0; k < n_slabs; k++) {
        const size_t off = static_cast<size_t>(k) * slab, len = std::min(slab, comp_len - off);
        HB_CUDA(cudaMemcpyAsync(ctx->stage_in.p + off, comp + off, len, cudaMemcpyHostToDevice, ctx->s_h2d));
        HB_CUDA(cudaEventRecord(ctx->ev_in[k], ctx->s_h2d));
    }
    uint64_t entry = 0;
    size_t base = 0;
    hb_status rc = HB_OK;
    for (int k = 0; k < n_slabs && rc == HB_OK; k++) {
        const uint64_t own_begin = static_cast<uint64_t>(k) * slab * 8;
        const uint64_t own_end = std::min<uint64_t>(static_cast<uint64_t>(k + 1) * slab * 8, total_bits);
        if (own_begin >= own_end) break;
        // the last chunk of a slab straddles into the next slab: its bytes must have arrived too
        HB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[std::min(k + 1, n_slabs - 1)], 0));
        hb_shard_info info;
        bool refuted = false;
        if (entry < own_end) {
            rc = run_fused(ctx, ctx->stage_in.p, total_bits, std::max<uint64_t>(own_begin, std::min(entry, own_end)), own_end,
                           entry, 0, tree, teams, ctx->stage_out.p + base, dev_cap - base, &info, &refuted);
            if (rc != HB_OK) break;
            if (refuted) note_refuted(ctx, tree);
            if (refuted || base + info.n_letters > cap) { rc = HB_ERR_CAPACITY; break; }     // -> unpipelined path
            const size_t n_k = static_cast<size_t>(info.n_letters);
            if (n_k) HB_CUDA(cudaMemcpyAsync(dst + base, ctx->stage_out.p + base, n_k, cudaMemcpyDeviceToHost, ctx->s_d2h));
            base += n_k;
            entry = info.exit_bit;
        }
    }
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->s_d2h);
    if (rc == HB_ERR_CAPACITY) return HB_OK;                             // *done stays false
    HB_TRY(rc);
    *out_n = base;
    *done = true;
    ctx->last_fused = 1;
    ctx->last_dec_valid = false;
    return HB_OK;
}

static hb_status decompress_host_common(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint8_t padding_bits,
                                        const hb_tree *tree, uint8_t **out, uint8_t *dst, size_t cap, size_t *out_n) {
    if (out) *out = nullptr;
    *out_n = 0;
    if (comp_len == 0) return HB_ERR_EMPTY_COMP;
    if (padding_bits > 7) return HB_ERR_BAD_PADDING;
    if (!comp) return HB_ERR_INVALID_ARG;
    const uint64_t total_bits = static_cast<uint64_t>(comp_len) * 8 - padding_bits;
    HB_TRY(prepare_dec_tree(ctx, tree));
    if (ctx->dec_fixed_len) {
        const size_t n = static_cast<size_t>(total_bits / ctx->dec_fixed_len);   // trailing bits that complete no code are dropped
        if (dst && n > cap) { *out_n = n; return HB_ERR_CAPACITY; }
        uint8_t *host = dst ? dst : static_cast<uint8_t *>(std::malloc(n ? n : 1));
        if (!host) return HB_ERR_NO_MEM;
        if (n) {
            hb_status rc = decompress_host_fixed_pipelined(ctx, comp, comp_len, total_bits, host, n);
            if (rc != HB_OK) { if (!dst) std::free(host); return rc; }
        }
        if (out) *out = host;
        *out_n = n;
        return HB_OK;
    }
    HB_TRY(ctx->stage_in.reserve(comp_len + 16));
    // the letter count is not known before decoding: size the device staging for the most letters this stream can hold
    // (every code the tree's shortest), capped by what the caller can take
    const uint64_t bound = total_bits / std::max<uint32_t>(tree->min_len, 1) + 1;
    const size_t dev_cap = static_cast<size_t>(dst ? std::min<uint64_t>(bound, cap) : bound);
    HB_TRY(ctx->stage_out.reserve(dev_cap + 64));
    // long streams a fused-decodable tree: slab pipeline (H2D of slab k+1 | decode of slab k | D2H of slab k-1)
    if (dst && comp_len >= (static_cast<size_t>(64) << 20) && fused_teams(ctx, tree) > 0) {
        bool done = false;
        HB_TRY(decompress_host_slabs(ctx, comp, comp_len, total_bits, tree, dst, cap, dev_cap, out_n, &done));
        if (done) return HB_OK;
        // (refuted speculation or a buffer that is too small: the plain path below decides and reports)
    }
    HB_CUDA(cudaMemcpyAsync(ctx->stage_in.p, comp, comp_len, cudaMemcpyHostToDevice, ctx->stream));
    hb_shard_info info;
    info.entry_bit = 0;
    info.exit_bit = 0;
    info.n_letters = 0;
    hb_status st = decode_range(ctx, ctx->stage_in.p, total_bits, 0, total_bits, 0, 0, tree, ctx->stage_out.p, dev_cap, &info);
    const size_t n = static_cast<size_t>(info.n_letters);
    if (st == HB_ERR_CAPACITY) { *out_n = n; return st; }
    HB_TRY(st);
    if (dst && n > cap) { *out_n = n; return HB_ERR_CAPACITY; }
    uint8_t *host = dst ? dst : static_cast<uint8_t *>(std::malloc(n ? n : 1));
    if (!host) return HB_ERR_NO_MEM;
    if (n) {
        cudaError_t e = cudaMemcpyAsync(host, ctx->stage_out.p, n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { if (!dst) std::free(host); return cuda_fail(e, "D2H of the decoded letters", __LINE__); }
    }
    if (out) *out = host;
    *out_n = n;
    return HB_OK;
}

hb_status hb_decompress_u8(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint8_t padding_bits,
                           const hb_tree *tree, uint8_t **out, size_t *out_n) {
    HB_ENTER(ctx);
    if (!tree || !out || !out_n) return HB_ERR_INVALID_ARG;
    return decompress_host_common(ctx, comp, comp_len, padding_bits, tree, out, nullptr, 0, out_n);
}

hb_status hb_decompress_u8_into(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint8_t padding_bits,
                                const hb_tree *tree, uint8_t *out, size_t out_cap, size_t *out_n) {
    HB_ENTER(ctx);
    if (!tree || !out || !out_n) return HB_ERR_INVALID_ARG;
    return decompress_host_common(ctx, comp, comp_len, padding_bits, tree, nullptr, out, out_cap, out_n);
}

hb_status hb_compress_u8_into(hb_ctx *ctx, const uint8_t *data, size_t n, int order_mode, hb_tree *tree_out,
                              uint8_t *comp_bytes, size_t comp_cap, size_t *comp_len, uint8_t *padding_bits) {
    HB_ENTER(ctx);
    if (!tree_out || !comp_bytes || !comp_len || !padding_bits || (n && !data)) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_WEIGHTS;
    return compress_host_common(ctx, data, n, nullptr, order_mode, tree_out, nullptr, comp_bytes, comp_cap, comp_len, padding_bits, nullptr);
}

hb_status hb_compress_with_tree_u8_into(hb_ctx *ctx, const uint8_t *data, size_t n, const hb_tree *tree,
                                        uint8_t *comp_bytes, size_t comp_cap, size_t *comp_len, uint8_t *padding_bits,
                                        uint8_t *missing) {
    HB_ENTER(ctx);
    if (!tree || !comp_bytes || !comp_len || !padding_bits || (n && !data)) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_COMP;
    return compress_host_common(ctx, data, n, tree, 0, nullptr, nullptr, comp_bytes, comp_cap, comp_len, padding_bits, missing);
}

}  // extern "C"
