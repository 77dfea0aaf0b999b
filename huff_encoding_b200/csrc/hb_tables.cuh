// hb_tables.cuh -- device-side construction of every lookup table the codec kernels use.
//
// The tree is built on the host (hb_tree.cpp: at most 513 nodes, reference tie-breaks), but the TABLES derived from it
// -- 4 096 + 8 192 + up to 65 536 decoder entries, the multi-letter emit table, the lane-independent encoder table --
// are built here by one small kernel each.  The tree travels as a KERNEL PARAMETER (2.3 KiB, copied by the launch
// itself): no host table construction (it cost 0.7-0.8 ms per new tree), no 158 KiB upload, no stream synchronisation,
// no pageable-memory copy.  A decompress() of a tree the context has not seen costs one ~10 us kernel.
//
// Reference semantics reproduced by the tables: comp.rs:496-509 (0 -> left, 1 -> right, emit at a leaf, restart at the
// root; a lone root emits its letter for every bit) and tree_inner.rs:422-440 (left appends 0, right appends 1).
#pragma once

#include "hb_common.cuh"
#include "hb_decode.cuh"
#include "hb_encode.cuh"

namespace hb {

// the flat tree as a kernel parameter: nodes[i] = left | right << 16 ; leaf: left = 0xFFFF, right = letter
struct TreeParam {
    uint32_t nodes[HB_MAX_NODES];
    uint32_t root;
    uint32_t n_nodes;
    uint32_t emit_bits;                // index width of the emit table to build (0 = do not build it)
    uint32_t pad;
    uint8_t code_len[256];             // code length per letter (0 = none); copied to the device for the fused decoder
};
static_assert(sizeof(TreeParam) <= 4000, "TreeParam must fit the kernel parameter space");

struct CodesParam {                    // read_codes() (tree_inner.rs:388-419) as a kernel parameter
    unsigned long long code[256];      // right-aligned
    uint8_t len[256];                  // 0 = no code (or longer than HB_MAX_ENCODE_BITS)
};
static_assert(sizeof(CodesParam) <= 4000, "CodesParam must fit the kernel parameter space");

constexpr int kTabThreads = 1024;

__device__ __forceinline__ bool tab_is_leaf(uint32_t nd) { return (nd & 0xFFFFu) == 0xFFFFu; }
__device__ __forceinline__ uint32_t tab_child(uint32_t nd, uint32_t bit) { return bit ? (nd >> 16) : (nd & 0xFFFFu); }

// One CTA.  Fills DecTables (first-level table, multi-letter count table, second-level tables + slot list, node copy)
// and, when tp.emit_bits != 0, the multi-letter emit table `emit` (1 << emit_bits entries) used by the fused decoder.
__global__ void __launch_bounds__(kTabThreads)
dec_tables_kernel(const TreeParam tp, DecTables *__restrict__ t, uint32_t *__restrict__ emit,
                  uint8_t *__restrict__ lens_out, int cnt_bits) {
    __shared__ uint32_t s_nodes[HB_MAX_NODES];
    __shared__ uint32_t s_scan[kTabThreads / 32];
    __shared__ uint16_t s_slot_node[256];
    __shared__ uint32_t s_n_slots;
    const int tid = threadIdx.x;
    for (int i = tid; i < HB_MAX_NODES; i += kTabThreads) {
        const uint32_t nd = i < static_cast<int>(tp.n_nodes) ? tp.nodes[i] : 0xFFFFu;
        s_nodes[i] = nd;
        t->nodes[i] = nd;
    }
    if (tid == 0) t->root = tp.root;
    if (tid < 256 && lens_out) lens_out[tid] = tp.code_len[tid];
    __syncthreads();
    const uint32_t root = tp.root;
    const bool lone = tab_is_leaf(s_nodes[root]);
    constexpr int K = kLutBits;

    // ---- first level: thread tid owns prefixes 4 tid .. 4 tid + 3 (so a block scan numbers the long ones in order)
    uint32_t n_long = 0;
    uint32_t deep[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t p = 4u * tid + j;
        deep[j] = 0xFFFFu;
        if (lone) {
            t->lut[p] = static_cast<uint16_t>((1u << kLutLenShift) | (s_nodes[root] >> 16));
            continue;
        }
        uint32_t node = root, used = 0;
        while (used < K && !tab_is_leaf(s_nodes[node])) {
            node = tab_child(s_nodes[node], (p >> (K - 1 - used)) & 1u);
            used++;
        }
        if (tab_is_leaf(s_nodes[node])) {
            t->lut[p] = static_cast<uint16_t>((used << kLutLenShift) | (s_nodes[node] >> 16));
        } else {
            deep[j] = node;                    // every K-bit prefix reaches a different depth-K node
            n_long++;
        }
    }
    // exclusive block scan of n_long -> slot numbers in ascending prefix order (= the order the codes are met)
    const uint32_t incl = warp_incl_scan(n_long);
    if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    for (int k = 0; k < kTabThreads / 32; k++) { if (k < (tid >> 5)) before += s_scan[k]; total += s_scan[k]; }
    uint32_t slot = before + incl - n_long;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (deep[j] != 0xFFFFu) {
            t->lut[4u * tid + j] = static_cast<uint16_t>(kLutLongFlag | slot);
            s_slot_node[slot] = static_cast<uint16_t>(deep[j]);
            t->slot_node[slot] = static_cast<uint16_t>(deep[j]);
            slot++;
        }
    }
    if (tid == 0) s_n_slots = total;
    __syncthreads();

    // ---- second level: 256 entries per slot, stream bits K .. K+7 below the slot's node
    for (uint32_t idx = tid; idx < s_n_slots * 256u; idx += kTabThreads) {
        const uint32_t sl = idx >> 8, b8 = idx & 255u;
        uint32_t nd = s_slot_node[sl], extra = 0;
        while (extra < 8 && !tab_is_leaf(s_nodes[nd])) {
            nd = tab_child(s_nodes[nd], (b8 >> (7 - extra)) & 1u);
            extra++;
        }
        t->lut2[idx] = tab_is_leaf(s_nodes[nd])
            ? static_cast<uint16_t>(((K + extra) << kLutLenShift) | (s_nodes[nd] >> 16))
            : static_cast<uint16_t>(kLutLongFlag | sl);
    }

    // ---- multi-letter count table over CB bits: greedy run of complete code words
    const uint32_t CB = static_cast<uint32_t>(cnt_bits);
    for (uint32_t p = tid; p < (1u << CB); p += kTabThreads) {
        if (lone) { t->cnt[p] = static_cast<uint8_t>((CB << 4) | CB); continue; }
        uint32_t pos = 0, letters = 0;
        for (;;) {
            uint32_t nd = root, q = pos;
            while (q < CB && !tab_is_leaf(s_nodes[nd])) {
                nd = tab_child(s_nodes[nd], (p >> (CB - 1 - q)) & 1u);
                q++;
            }
            if (!tab_is_leaf(s_nodes[nd])) break;
            pos = q;
            letters++;
            if (pos >= CB) break;
        }
        t->cnt[p] = static_cast<uint8_t>((pos << 4) | letters);
    }

    // ---- multi-letter emit table: up to three complete code words of the next EB bits.
    //      entry = letter0 | letter1 << 8 | letter2 << 16 | count << 24 | bits consumed << 28   (count 0: the first code
    //      word is longer than EB bits)
    const uint32_t EB = tp.emit_bits;
    if (EB && emit && !lone) {
        for (uint32_t p = tid; p < (1u << EB); p += kTabThreads) {
            uint32_t pos = 0, letters = 0, e = 0;
            while (letters < 3) {
                uint32_t nd = root, q = pos;
                while (q < EB && !tab_is_leaf(s_nodes[nd])) {
                    nd = tab_child(s_nodes[nd], (p >> (EB - 1 - q)) & 1u);
                    q++;
                }
                if (!tab_is_leaf(s_nodes[nd])) break;
                e |= (s_nodes[nd] >> 16) << (8 * letters);
                pos = q;
                letters++;
                if (pos >= EB) break;
            }
            emit[p] = e | (letters << 24) | (pos << 28);
        }
    }
}

// Encoder table (hb_encode.cuh EncTable) from the code list: 256 threads.
__global__ void __launch_bounds__(256)
enc_table_kernel(const CodesParam cp, EncTable *__restrict__ t) {
    const int b = threadIdx.x;
    const uint32_t len = cp.len[b];
    const unsigned long long left = len ? cp.code[b] << (64 - len) : 0ull;     // code left-aligned in 64 bits
    t->lo[b] = make_uint2(static_cast<uint32_t>(left >> 32), len);
    t->hi[b] = static_cast<uint32_t>(left);
    t->packed[b] = (len && len <= 16) ? ((static_cast<uint32_t>(left >> 32) & 0xFFFF0000u) | len) : 0u;
}

}  // namespace hb
