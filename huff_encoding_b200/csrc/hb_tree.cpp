// hb_tree.cpp -- host side of libhuffb200: Huffman tree + code table + (de)serialisation.
//
// Tree construction stays on the host (at most 513 nodes) and reproduces the reference's merge order
// bit for bit, including the tie-breaks it inherits from Rust's std::collections::BinaryHeap:
//   tree/branch_heap.rs:18-83   min-heap through a reversed comparator, built by sequential pushes
//   tree/tree_inner.rs:281-320  merge loop: first pop -> left (bit 0), second pop -> right (bit 1)
//   tree/tree_inner.rs:388-440  code assignment; duplicate letters keep the code of the last DFS visit
//   tree/tree_inner.rs:522-668  try_from_bin / as_bin
//   comp.rs:128-184, 279-300    CompressData::try_from_bytes / to_bytes
// (paths relative to /root/reference/huff_coding/src).
#include "../../include/huffb200.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

namespace {

// ---- std::collections::BinaryHeap, specialised to "smaller weight = greater in heap order"
class BranchHeap {
public:
    BranchHeap() = default;
    size_t size() const { return n_; }

    void push(uint16_t id, uint64_t weight) {
        slots_[n_] = Slot{weight, id};
        bubble_up(n_++);
    }

    uint16_t pop_min() {
        const Slot last = slots_[--n_];
        if (n_ == 0) return last.id;
        const uint16_t top = slots_[0].id;
        // std: put the former last element at the root, walk the hole down to a leaf position taking the
        // child that is >= its sibling in heap order (the right one on equal weights), then bubble up.
        size_t hole = 0;
        const size_t n = n_;
        for (;;) {
            const size_t l = 2 * hole + 1, r = l + 1;
            if (r < n) {
                const size_t pick = l + static_cast<size_t>(slots_[l].w >= slots_[r].w);   // branch-free: data-random
                slots_[hole] = slots_[pick];
                hole = pick;
            } else if (l < n) {
                slots_[hole] = slots_[l];
                hole = l;
                break;
            } else {
                break;
            }
        }
        slots_[hole] = last;
        bubble_up(hole);
        return top;
    }

private:
    struct Slot { uint64_t w; uint16_t id; };        // the weight travels with the id: no indirection per comparison
    void bubble_up(size_t pos) {
        const Slot moving = slots_[pos];
        while (pos > 0) {
            const size_t parent = (pos - 1) / 2;
            if (moving.w >= slots_[parent].w) break;     // std sift_up: stop when element <= parent in heap order
            slots_[pos] = slots_[parent];
            pos = parent;
        }
        slots_[pos] = moving;
    }
    Slot slots_[HB_MAX_LEAVES + 1];                      // at most one slot per leaf is ever live
    size_t n_ = 0;
};

bool is_leaf(const hb_node &n) { return n.left == HB_NO_CHILD; }

// tree_inner.rs:422-440 + 388-419.  Iterative DFS, left before right; code kept as a bit vector so depth > 64 is fine.
void fill_code_table(hb_tree *t) {
    std::memset(t->has_code, 0, sizeof t->has_code);
    std::memset(t->code_len, 0, sizeof t->code_len);
    std::memset(t->code, 0, sizeof t->code);
    t->n_leaves = 0;

    const hb_node &root = t->nodes[t->root];
    if (is_leaf(root)) {                       // tree_inner.rs:313-315: lone root letter -> code [0]
        t->has_code[root.letter] = 1;
        t->code_len[root.letter] = 1;
        t->code[root.letter] = 0;
        t->n_leaves = 1;
    } else {
        // the frame carries its code (valid while depth <= 64; deeper codes are stored as 0, only their length counts)
        struct Frame { uint16_t node; uint16_t depth; uint64_t code; };
        Frame todo[HB_MAX_LEAVES + 2];                   // pending right siblings along one root-to-leaf path (+ 1)
        int top = 0;
        todo[top++] = Frame{static_cast<uint16_t>(t->root), 0, 0};
        while (top > 0) {
            const Frame f = todo[--top];
            const hb_node &nd = t->nodes[f.node];
            if (is_leaf(nd)) {
                t->n_leaves++;
                t->has_code[nd.letter] = 1;
                t->code_len[nd.letter] = f.depth;
                t->code[nd.letter] = f.depth <= 64 ? f.code : 0;   // later visits overwrite earlier ones (HashMap::insert)
                continue;
            }
            const uint16_t d = static_cast<uint16_t>(f.depth + 1);
            todo[top++] = Frame{nd.right, d, (f.code << 1) | 1u};     // popped second
            todo[top++] = Frame{nd.left, d, f.code << 1};             // popped first
        }
    }
    // longest / shortest code and the gcd of all lengths (gcd over the DISTINCT lengths: at most a few dozen)
    uint32_t mx = 0, mn = 0xFFFFFFFFu, g = 0;
    uint64_t seen[5] = {0, 0, 0, 0, 0};                    // lengths 0..319 (a tree has at most 257 leaves)
    for (int b = 0; b < 256; b++)
        if (t->has_code[b]) {
            const uint32_t len = t->code_len[b];
            mx = std::max<uint32_t>(mx, len);
            mn = std::min<uint32_t>(mn, len);
            if (len < 320) seen[len >> 6] |= 1ull << (len & 63); else g = std::gcd(g, len);
        }
    for (uint32_t len = 1; len < 320 && len <= mx; len++)
        if (seen[len >> 6] >> (len & 63) & 1) g = std::gcd(g, len);
    t->max_len = mx;
    t->min_len = mn;
    t->len_gcd = g;
}

struct BitSink {
    uint8_t *buf; size_t cap_bits; size_t n = 0; bool overflow = false;
    void put(int bit) {
        if (n >= cap_bits) { overflow = true; return; }
        if ((n & 7) == 0) buf[n >> 3] = 0;
        if (bit) buf[n >> 3] |= static_cast<uint8_t>(0x80u >> (n & 7));
        n++;
    }
};

}  // namespace

extern "C" {

hb_status hb_tree_from_pairs(const uint8_t *letters, const uint64_t *weights, size_t n, hb_tree *tree) {
    if (!tree || (n && (!letters || !weights))) return HB_ERR_INVALID_ARG;
    if (n == 0) return HB_ERR_EMPTY_WEIGHTS;              // tree_inner.rs:283-285
    if (n > HB_MAX_LEAVES) return HB_ERR_INVALID_ARG;
    std::memset(tree, 0, sizeof *tree);
    BranchHeap heap;
    for (size_t i = 0; i < n; i++) {                      // branch_heap.rs:52-58
        hb_node &nd = tree->nodes[tree->n_nodes];
        nd.left = nd.right = HB_NO_CHILD;
        nd.letter = letters[i];
        nd.weight = weights[i];
        heap.push(static_cast<uint16_t>(tree->n_nodes++), nd.weight);
    }
    while (heap.size() > 1) {                             // tree_inner.rs:289-303
        uint16_t lo = heap.pop_min();
        uint16_t next = heap.pop_min();
        hb_node &nd = tree->nodes[tree->n_nodes];
        nd.left = lo;
        nd.right = next;
        nd.letter = 0;
        nd.weight = tree->nodes[lo].weight + tree->nodes[next].weight;
        heap.push(static_cast<uint16_t>(tree->n_nodes++), nd.weight);
    }
    tree->root = heap.pop_min();                          // tree_inner.rs:306
    fill_code_table(tree);
    return HB_OK;
}

hb_status hb_tree_from_weights(const uint64_t weights[256], int order_mode, hb_tree *tree) {
    if (!weights || !tree) return HB_ERR_INVALID_ARG;
    if (order_mode != HB_ORDER_ASC && order_mode != HB_ORDER_BYTEWEIGHTS) return HB_ERR_INVALID_ARG;
    uint8_t letters[HB_MAX_LEAVES];
    uint64_t w[HB_MAX_LEAVES];
    size_t n = 0;
    for (int b = 0; b < 256; b++)
        if (weights[b]) { letters[n] = static_cast<uint8_t>(b); w[n] = weights[b]; n++; }
    // weights.rs:396-415: ByteWeights' iterator re-yields byte 0 when bin 255 is empty (index 256 wraps to 0)
    if (order_mode == HB_ORDER_BYTEWEIGHTS && n && weights[0] && !weights[255]) {
        letters[n] = 0; w[n] = weights[0]; n++;
    }
    return hb_tree_from_pairs(letters, w, n, tree);
}

hb_status hb_stream_bits(const uint64_t weights[256], const hb_tree *tree, uint64_t *bits, uint8_t *missing) {
    if (!weights || !tree || !bits) return HB_ERR_INVALID_ARG;
    uint64_t total = 0;
    int first_missing = -1;
    for (int b = 0; b < 256; b++) {
        if (!weights[b]) continue;
        if (!tree->has_code[b]) { if (first_missing < 0) first_missing = b; continue; }
        total += weights[b] * tree->code_len[b];
    }
    *bits = total;
    if (first_missing >= 0) { if (missing) *missing = static_cast<uint8_t>(first_missing); return HB_ERR_MISSING_LETTER; }
    return HB_OK;
}

hb_status hb_shard_plan(const uint64_t *hists, size_t n_shards, int order_mode, hb_tree *tree_out, uint64_t *shard_bits) {
    if (!hists || !n_shards || !tree_out || !shard_bits) return HB_ERR_INVALID_ARG;
    uint64_t total[256] = {0};
    for (size_t g = 0; g < n_shards; g++)
        for (int b = 0; b < 256; b++) total[b] += hists[g * 256 + b];
    const hb_status st = hb_tree_from_weights(total, order_mode, tree_out);
    if (st != HB_OK) return st;
    for (size_t g = 0; g < n_shards; g++) {
        uint64_t bits = 0;
        for (int b = 0; b < 256; b++) bits += hists[g * 256 + b] * tree_out->code_len[b];
        shard_bits[g] = bits;
    }
    return HB_OK;
}

hb_status hb_tree_as_bin(const hb_tree *tree, uint8_t *out, size_t cap_bytes, size_t *n_bits) {
    if (!tree || !out || !n_bits) return HB_ERR_INVALID_ARG;
    BitSink sink{out, cap_bytes * 8};
    std::vector<uint16_t> todo{static_cast<uint16_t>(tree->root)};
    while (!todo.empty()) {                               // tree_inner.rs:637-663, preorder
        const hb_node &nd = tree->nodes[todo.back()];
        todo.pop_back();
        if (!is_leaf(nd)) {
            sink.put(1);
            todo.push_back(nd.right);
            todo.push_back(nd.left);
        } else {
            sink.put(0);
            for (int k = 7; k >= 0; k--) sink.put((nd.letter >> k) & 1);
        }
    }
    if (sink.overflow) return HB_ERR_CAPACITY;
    *n_bits = sink.n;
    return HB_OK;
}

hb_status hb_tree_from_bin(const uint8_t *bin, size_t n_bits, hb_tree *tree) {
    if (!tree || (n_bits && !bin)) return HB_ERR_INVALID_ARG;
    std::memset(tree, 0, sizeof *tree);
    size_t pos = 0;
    auto next_bit = [&](int &bit) -> bool {
        if (pos >= n_bits) return false;
        bit = (bin[pos >> 3] >> (7 - (pos & 7))) & 1;
        pos++;
        return true;
    };
    // tree_inner.rs:526-578 without recursion: joints wait on a stack until both children are read
    struct Waiting { uint16_t node; bool has_left; };
    std::vector<Waiting> open;
    int root = -1;
    while (root < 0) {
        int bit;
        if (!next_bit(bit)) return HB_ERR_BIN_TOO_SMALL;
        if (tree->n_nodes >= HB_MAX_NODES) return HB_ERR_TREE_NODES;     // documented cap (huffb200.h)
        uint16_t id = static_cast<uint16_t>(tree->n_nodes++);
        hb_node &nd = tree->nodes[id];
        nd.left = nd.right = HB_NO_CHILD;
        if (bit) { open.push_back({id, false}); continue; }
        if (n_bits - pos < 8) return HB_ERR_BIN_TOO_SMALL;
        uint8_t letter = 0;
        for (int k = 0; k < 8; k++) { int b; next_bit(b); letter = static_cast<uint8_t>((letter << 1) | b); }
        nd.letter = letter;
        uint16_t done = id;
        for (;;) {
            if (open.empty()) { root = done; break; }
            Waiting &top = open.back();
            if (!top.has_left) { tree->nodes[top.node].left = done; top.has_left = true; break; }
            tree->nodes[top.node].right = done;
            done = top.node;
            open.pop_back();
        }
    }
    if (pos != n_bits) return HB_ERR_BIN_TOO_BIG;         // tree_inner.rs:586-590
    tree->root = static_cast<uint32_t>(root);
    fill_code_table(tree);
    return HB_OK;
}

hb_status hb_to_bytes(const uint8_t *comp, size_t comp_len, uint8_t padding_bits, const hb_tree *tree,
                      uint8_t *out, size_t cap, size_t *out_len) {
    if (!comp || !tree || !out || !out_len) return HB_ERR_INVALID_ARG;
    uint8_t tree_bin[(HB_MAX_LEAVES * 10) / 8 + 8];
    size_t n_bits = 0;
    hb_status rc = hb_tree_as_bin(tree, tree_bin, sizeof tree_bin, &n_bits);
    if (rc != HB_OK) return rc;
    uint8_t tree_pad = static_cast<uint8_t>((8 - n_bits % 8) % 8);          // utils.rs:37-40
    uint32_t tree_bytes = static_cast<uint32_t>((n_bits + tree_pad) / 8);   // comp.rs:285
    size_t total = 5 + static_cast<size_t>(tree_bytes) + comp_len;
    if (total > cap) { *out_len = total; return HB_ERR_CAPACITY; }
    out[0] = static_cast<uint8_t>((tree_pad << 4) + padding_bits);          // comp.rs:289
    for (int k = 0; k < 4; k++) out[1 + k] = static_cast<uint8_t>(tree_bytes >> (24 - 8 * k));
    std::memcpy(out + 5, tree_bin, tree_bytes);
    std::memcpy(out + 5 + tree_bytes, comp, comp_len);
    *out_len = total;
    return HB_OK;
}

hb_status hb_try_from_bytes(const uint8_t *bytes, size_t n, hb_tree *tree,
                            size_t *data_off, size_t *data_len, uint8_t *padding_bits) {
    if (!tree || !data_off || !data_len || !padding_bits || (n && !bytes)) return HB_ERR_INVALID_ARG;
    if (n < 1) return HB_ERR_BYTES_SHORT;                 // comp.rs:143
    uint8_t tree_pad = bytes[0] >> 4, data_pad = bytes[0] & 0x0F;
    if (n < 5) return HB_ERR_BYTES_SHORT;                 // comp.rs:149
    size_t tree_len = (size_t(bytes[1]) << 24) | (size_t(bytes[2]) << 16) | (size_t(bytes[3]) << 8) | bytes[4];
    if (tree_len < 2) return HB_ERR_TREE_LEN;             // comp.rs:153-155
    if (n < 5 + tree_len) return HB_ERR_BYTES_SHORT;      // comp.rs:161
    size_t tree_bits = tree_len * 8;
    tree_bits = tree_bits >= tree_pad ? tree_bits - tree_pad : 0;           // comp.rs:164
    const hb_status tst = hb_tree_from_bin(bytes + 5, tree_bits, tree);
    if (tst == HB_ERR_TREE_NODES) return tst;
    if (tst != HB_OK) return HB_ERR_INVALID_TREE;
    size_t dlen = n - 5 - tree_len;
    if (dlen == 0) return HB_ERR_EMPTY_COMP;              // comp.rs:179-183 -> :56-58
    if (data_pad > 7) return HB_ERR_BAD_PADDING;          // comp.rs:59-61
    *data_off = 5 + tree_len;
    *data_len = dlen;
    *padding_bits = data_pad;
    return HB_OK;
}

}  // extern "C"
