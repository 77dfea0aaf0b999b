/*
 * huff_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See huff_oracle.h.
 *
 * Plain-C restatement of huff_coding's u8 path.  The loops are kept in the reference's
 * shape on purpose (per-letter table lookup + per-bit packing, bit-serial tree walk), so
 * that timing this file is a fair "port" CPU baseline of the reference algorithm.
 */
#include "huff_oracle.h"

#include <string.h>

/* ------------------------------------------------------------------ histogram */

/* weights.rs:116-123 and weights.rs:265-279: one increment per letter, 64-bit counts. */
void ho_histogram(const uint8_t *data, size_t n, uint64_t w[256])
{
    memset(w, 0, 256 * sizeof(uint64_t));
    for (size_t i = 0; i < n; i++)
        w[data[i]] += 1;
}

/* ------------------------------------------------------------------ heap
 * tree/branch_heap.rs:18-83 wraps std::collections::BinaryHeap with a reversed comparator
 * (`other.leaf().cmp(self.leaf())`, branch_heap.rs:67-71; leaves compare by weight only,
 * leaf.rs:31-47), i.e. a min-heap on weight.  The moves below are std's:
 *   push  = append, sift_up(0, last)           (stop when element <= parent in heap order,
 *                                               i.e. when x.w >= parent.w)
 *   pop   = take last; if heap not empty swap it with data[0], sift_down_to_bottom(0):
 *           walk the hole to the bottom always choosing the "greater" child in heap order,
 *           right child on ties (`child += (left <= right)`), then sift_up from there.
 */
typedef struct {
    uint16_t d[HO_MAX_NODES];      /* node indices */
    size_t   len;
} ho_heap;

static void heap_sift_up(ho_heap *h, const ho_node *nodes, size_t start, size_t pos)
{
    uint16_t elem = h->d[pos];
    uint64_t ew = nodes[elem].weight;
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        /* `if hole.element() <= hole.get(parent) { break }` ; a <= b in heap order <=> a.w >= b.w */
        if (ew >= nodes[h->d[parent]].weight)
            break;
        h->d[pos] = h->d[parent];
        pos = parent;
    }
    h->d[pos] = elem;
}

static void heap_push(ho_heap *h, const ho_node *nodes, uint16_t item)
{
    size_t old_len = h->len;
    h->d[h->len++] = item;
    heap_sift_up(h, nodes, 0, old_len);
}

static uint16_t heap_pop(ho_heap *h, const ho_node *nodes)
{
    uint16_t item = h->d[--h->len];
    if (h->len == 0)
        return item;
    /* swap(&mut item, &mut self.data[0]) */
    uint16_t top = h->d[0];
    h->d[0] = item;
    item = top;
    /* sift_down_to_bottom(0) */
    size_t end = h->len;
    size_t pos = 0;
    uint16_t elem = h->d[0];
    size_t child = 1;
    while (end >= 2 && child <= end - 2) {
        /* child += (hole.get(child) <= hole.get(child + 1)); left <= right <=> left.w >= right.w */
        if (nodes[h->d[child]].weight >= nodes[h->d[child + 1]].weight)
            child += 1;
        h->d[pos] = h->d[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) {
        h->d[pos] = h->d[child];
        pos = child;
    }
    h->d[pos] = elem;
    heap_sift_up(h, nodes, 0, pos);
    return item;
}

/* ------------------------------------------------------------------ codes */

/* tree_inner.rs:422-440 (set_codes_in_child_branches: left appends 0, right appends 1) combined with
 * tree_inner.rs:388-419 (read_codes: DFS left then right, `codes.insert` => a duplicate letter keeps the
 * code of its LAST visit).  Iterative preorder with an explicit stack; depth <= HO_MAX_LEAVES. */
static void assign_codes(ho_tree *t)
{
    memset(t->has_code, 0, sizeof t->has_code);
    memset(t->code_len, 0, sizeof t->code_len);
    memset(t->code_bits, 0, sizeof t->code_bits);

    const ho_node *root = &t->nodes[t->root];
    if (root->left == HO_NONE) {
        /* tree_inner.rs:313-315 / :415-417: a lone root letter gets the code [0] */
        t->has_code[root->letter] = 1;
        t->code_len[root->letter] = 1;
        return;
    }
    struct { uint16_t node; uint16_t depth; uint8_t state; } stack[HO_MAX_LEAVES + 2];
    uint8_t path[HO_CODE_BYTES];
    memset(path, 0, sizeof path);
    int sp = 0;
    stack[0].node = (uint16_t)t->root; stack[0].depth = 0; stack[0].state = 0;
    while (sp >= 0) {
        uint16_t n = stack[sp].node;
        uint16_t depth = stack[sp].depth;
        const ho_node *nd = &t->nodes[n];
        if (nd->left == HO_NONE) {
            t->has_code[nd->letter] = 1;
            t->code_len[nd->letter] = depth;
            memcpy(t->code_bits[nd->letter], path, HO_CODE_BYTES);
            /* clear bits beyond depth so equal codes compare equal bytewise */
            for (unsigned k = depth; k < HO_CODE_BYTES * 8; k++)
                t->code_bits[nd->letter][k >> 3] &= (uint8_t)~(0x80u >> (k & 7));
            sp--;
            continue;
        }
        if (stack[sp].state == 0) {            /* descend left with bit 0 */
            stack[sp].state = 1;
            path[depth >> 3] &= (uint8_t)~(0x80u >> (depth & 7));
            sp++;
            stack[sp].node = nd->left; stack[sp].depth = (uint16_t)(depth + 1); stack[sp].state = 0;
        } else if (stack[sp].state == 1) {     /* descend right with bit 1 */
            stack[sp].state = 2;
            path[depth >> 3] |= (uint8_t)(0x80u >> (depth & 7));
            sp++;
            stack[sp].node = nd->right; stack[sp].depth = (uint16_t)(depth + 1); stack[sp].state = 0;
        } else {
            path[depth >> 3] &= (uint8_t)~(0x80u >> (depth & 7));
            sp--;
        }
    }
}

/* ------------------------------------------------------------------ tree */

/* tree_inner.rs:281-320 over branch_heap.rs:24-58 */
int ho_tree_from_pairs(const uint8_t *letters, const uint64_t *weights, size_t n, ho_tree *t)
{
    if (n == 0)
        return HO_ERR_EMPTY_WEIGHTS;            /* tree_inner.rs:283-285 */
    if (n > HO_MAX_LEAVES)
        return HO_ERR_CAPACITY;
    memset(t, 0, sizeof *t);
    ho_heap h; h.len = 0;
    /* branch_heap.rs:52-58: one push per (letter, weight) in iteration order */
    for (size_t i = 0; i < n; i++) {
        ho_node *nd = &t->nodes[t->n_nodes];
        nd->left = nd->right = HO_NONE;
        nd->letter = letters[i];
        nd->weight = weights[i];
        heap_push(&h, t->nodes, (uint16_t)t->n_nodes);
        t->n_nodes++;
    }
    /* tree_inner.rs:289-303: min = first pop (left, bit 0), next_min = second pop (right, bit 1) */
    while (h.len > 1) {
        uint16_t mn = heap_pop(&h, t->nodes);
        uint16_t nx = heap_pop(&h, t->nodes);
        ho_node *nd = &t->nodes[t->n_nodes];
        nd->left = mn; nd->right = nx; nd->letter = 0;
        nd->weight = t->nodes[mn].weight + t->nodes[nx].weight;
        heap_push(&h, t->nodes, (uint16_t)t->n_nodes);
        t->n_nodes++;
    }
    t->root = heap_pop(&h, t->nodes);           /* tree_inner.rs:306 */
    assign_codes(t);
    return HO_OK;
}

int ho_tree_from_weights(const uint64_t w[256], int order_mode, ho_tree *t)
{
    uint8_t letters[HO_MAX_LEAVES];
    uint64_t weights[HO_MAX_LEAVES];
    size_t n = 0;
    for (int b = 0; b < 256; b++)
        if (w[b] != 0) { letters[n] = (uint8_t)b; weights[n] = w[b]; n++; }
    if (order_mode == HO_ORDER_BYTEWEIGHTS) {
        /* weights.rs:396-415: the iterator indexes with `current_index as u8`; when index 255 is empty the
         * scan reaches 256, which wraps to byte 0; if byte 0 is present it is yielded a second time. */
        if (n > 0 && w[0] != 0 && w[255] == 0) { letters[n] = 0; weights[n] = w[0]; n++; }
    }
    return ho_tree_from_pairs(letters, weights, n, t);
}

/* ------------------------------------------------------------------ compress */

uint8_t ho_calc_padding_bits(uint64_t bit_count)
{
    /* utils.rs:37-40 */
    uint8_t n = (uint8_t)(8 - bit_count % 8);
    return n == 8 ? 0 : n;
}

int ho_compressed_bits(const uint8_t *data, size_t n, const ho_tree *t, uint64_t *bits, uint8_t *missing)
{
    uint64_t total = 0;
    for (size_t i = 0; i < n; i++) {
        if (!t->has_code[data[i]]) { if (missing) *missing = data[i]; return HO_ERR_MISSING_LETTER; }
        total += t->code_len[data[i]];
    }
    *bits = total;
    return HO_OK;
}

/* comp.rs:419-451 */
int ho_compress_with_tree(const uint8_t *data, size_t n, const ho_tree *t,
                          uint8_t *out, size_t cap, size_t *out_len, uint8_t *padding_bits, uint8_t *missing)
{
    size_t o = 0;
    uint8_t comp_byte = 0;
    int bit_ptr = 7;
    for (size_t i = 0; i < n; i++) {
        uint8_t letter = data[i];
        if (!t->has_code[letter]) {              /* comp.rs:426-432 */
            if (missing) *missing = letter;
            return HO_ERR_MISSING_LETTER;
        }
        const uint8_t *code = t->code_bits[letter];
        unsigned len = t->code_len[letter];
        for (unsigned k = 0; k < len; k++) {     /* comp.rs:433-443 */
            unsigned bit = (code[k >> 3] >> (7 - (k & 7))) & 1u;
            comp_byte |= (uint8_t)(bit << bit_ptr);
            if (bit_ptr == 0) {
                if (o >= cap) return HO_ERR_CAPACITY;
                out[o++] = comp_byte;
                comp_byte = 0;
                bit_ptr = 7;
            } else {
                bit_ptr -= 1;
            }
        }
    }
    uint8_t pad = (bit_ptr == 7) ? 0 : (uint8_t)(bit_ptr + 1);   /* comp.rs:446 */
    if (pad != 0) {                                              /* comp.rs:447 */
        if (o >= cap) return HO_ERR_CAPACITY;
        out[o++] = comp_byte;
    }
    if (o == 0)
        return HO_ERR_EMPTY_COMP;                /* comp.rs:450 -> :56-58 (only reachable for n == 0) */
    *out_len = o;
    *padding_bits = pad;
    return HO_OK;
}

/* comp.rs:353-356 */
int ho_compress(const uint8_t *data, size_t n, int order_mode, ho_tree *t,
                uint8_t *out, size_t cap, size_t *out_len, uint8_t *padding_bits)
{
    uint64_t w[256];
    ho_histogram(data, n, w);
    int rc = ho_tree_from_weights(w, order_mode, t);
    if (rc != HO_OK) return rc;
    uint8_t missing;
    return ho_compress_with_tree(data, n, t, out, cap, out_len, padding_bits, &missing);
}

/* ------------------------------------------------------------------ decompress */

/* comp.rs:487-519: bit-serial walk; 0 -> left, 1 -> right; at a letter branch emit and return to the root.
 * A lone-root tree skips the descent, so every bit emits the root letter (comp.rs:496,506-509). */
static int decompress_impl(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t,
                           uint8_t *out, size_t cap, size_t *out_n)
{
    if (len == 0) return HO_ERR_EMPTY_COMP;      /* comp.rs:56-58 */
    if (padding_bits > 7) return HO_ERR_BAD_PADDING;
    size_t o = 0;
    uint32_t cur = t->root;
    for (size_t i = 0; i < len; i++) {
        uint8_t byte = comp[i];
        int nbits = (i == len - 1) ? 8 - padding_bits : 8;       /* comp.rs:513-516 */
        for (int bit_ptr = 0; bit_ptr < nbits; bit_ptr++) {
            const ho_node *nd = &t->nodes[cur];
            if (nd->left != HO_NONE)
                cur = ((byte >> (7 - bit_ptr)) & 1) ? nd->right : nd->left;
            nd = &t->nodes[cur];
            if (nd->left == HO_NONE) {
                if (out) {
                    if (o >= cap) return HO_ERR_CAPACITY;
                    out[o] = nd->letter;
                }
                o++;
                cur = t->root;
            }
        }
    }
    *out_n = o;
    return HO_OK;
}

int ho_decompress(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t,
                  uint8_t *out, size_t cap, size_t *out_n)
{
    static uint8_t dummy;
    return decompress_impl(comp, len, padding_bits, t, out ? out : &dummy, out ? cap : 0, out_n);
}

int ho_decompress_count(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t, size_t *out_n)
{
    return decompress_impl(comp, len, padding_bits, t, NULL, 0, out_n);
}

/* ------------------------------------------------------------------ tree <-> bits */

typedef struct { uint8_t *buf; size_t cap_bits; size_t n; int overflow; } bitw;

static void bw_push(bitw *w, int bit)
{
    if (w->n >= w->cap_bits) { w->overflow = 1; return; }
    if ((w->n & 7) == 0) w->buf[w->n >> 3] = 0;
    if (bit) w->buf[w->n >> 3] |= (uint8_t)(0x80u >> (w->n & 7));
    w->n++;
}

/* tree_inner.rs:632-668 */
int ho_tree_as_bin(const ho_tree *t, uint8_t *out, size_t cap, size_t *n_bits)
{
    bitw w = { out, cap * 8, 0, 0 };
    uint16_t stack[HO_MAX_NODES];
    int sp = 0;
    stack[sp++] = (uint16_t)t->root;
    while (sp > 0) {
        const ho_node *nd = &t->nodes[stack[--sp]];
        if (nd->left != HO_NONE) {
            bw_push(&w, 1);                       /* joint branch */
            stack[sp++] = nd->right;              /* left is visited first */
            stack[sp++] = nd->left;
        } else {
            bw_push(&w, 0);                       /* letter branch + the letter, big-endian, MSB first */
            for (int k = 0; k < 8; k++) bw_push(&w, (nd->letter >> (7 - k)) & 1);
        }
    }
    if (w.overflow) return HO_ERR_CAPACITY;
    *n_bits = w.n;
    return HO_OK;
}

typedef struct { const uint8_t *bin; size_t n_bits; size_t pos; } bitr;

/* tree_inner.rs:526-578, iterative.  Returns node index or -err. */
static int read_branches(bitr *r, ho_tree *t)
{
    /* explicit stack of joints waiting for children */
    struct { uint16_t node; uint8_t filled; } stack[HO_MAX_NODES];
    int sp = 0;
    int result = -1;
    for (;;) {
        if (r->pos >= r->n_bits) return -HO_ERR_BIN_TOO_SMALL;
        int bit = (r->bin[r->pos >> 3] >> (7 - (r->pos & 7))) & 1; r->pos++;
        uint16_t me;
        if (t->n_nodes >= HO_MAX_NODES) return -HO_ERR_CAPACITY;
        me = (uint16_t)t->n_nodes++;
        ho_node *nd = &t->nodes[me];
        nd->weight = 0;                            /* tree_inner.rs:538,573: weights are not stored */
        if (bit) {
            nd->left = nd->right = HO_NONE; nd->letter = 0;
            stack[sp].node = me; stack[sp].filled = 0; sp++;
            continue;
        }
        if (r->n_bits - r->pos < 8) return -HO_ERR_BIN_TOO_SMALL;
        uint8_t letter = 0;
        for (int k = 0; k < 8; k++) {
            letter = (uint8_t)((letter << 1) | ((r->bin[r->pos >> 3] >> (7 - (r->pos & 7))) & 1));
            r->pos++;
        }
        nd->left = nd->right = HO_NONE; nd->letter = letter;
        /* attach completed subtree to waiting joints */
        uint16_t done = me;
        for (;;) {
            if (sp == 0) { result = done; break; }
            ho_node *p = &t->nodes[stack[sp - 1].node];
            if (stack[sp - 1].filled == 0) { p->left = done; stack[sp - 1].filled = 1; break; }
            p->right = done; done = stack[sp - 1].node; sp--;
        }
        if (result >= 0) return result;
    }
}

/* tree_inner.rs:522-604 */
int ho_tree_from_bin(const uint8_t *bin, size_t n_bits, ho_tree *t)
{
    memset(t, 0, sizeof *t);
    bitr r = { bin, n_bits, 0 };
    int root = read_branches(&r, t);
    if (root < 0) return -root;
    if (r.pos != n_bits) return HO_ERR_BIN_TOO_BIG;       /* tree_inner.rs:586-590 */
    t->root = (uint32_t)root;
    assign_codes(t);
    return HO_OK;
}

/* ------------------------------------------------------------------ container */

/* comp.rs:279-300 */
int ho_to_bytes(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t,
                uint8_t *out, size_t cap, size_t *out_len)
{
    uint8_t tree_bin[(HO_MAX_LEAVES * 10 + 7) / 8 + 8];
    size_t n_bits;
    int rc = ho_tree_as_bin(t, tree_bin, sizeof tree_bin, &n_bits);
    if (rc != HO_OK) return rc;
    uint8_t tree_pad = ho_calc_padding_bits(n_bits);
    uint32_t tree_bytes_len = (uint32_t)((n_bits + tree_pad) / 8);
    size_t total = 1 + 4 + (size_t)tree_bytes_len + len;
    if (total > cap) return HO_ERR_CAPACITY;
    out[0] = (uint8_t)((tree_pad << 4) + padding_bits);
    out[1] = (uint8_t)(tree_bytes_len >> 24); out[2] = (uint8_t)(tree_bytes_len >> 16);
    out[3] = (uint8_t)(tree_bytes_len >> 8);  out[4] = (uint8_t)tree_bytes_len;
    memcpy(out + 5, tree_bin, tree_bytes_len);
    memcpy(out + 5 + tree_bytes_len, comp, len);
    *out_len = total;
    return HO_OK;
}

/* comp.rs:128-184 */
int ho_try_from_bytes(const uint8_t *bytes, size_t n, ho_tree *t,
                      size_t *data_off, size_t *data_len, uint8_t *padding_bits)
{
    if (n < 1) return HO_ERR_BYTES_SHORT;                       /* "slice is empty" */
    uint8_t tree_pad = bytes[0] >> 4;
    uint8_t data_pad = bytes[0] & 0x0F;
    if (n < 5) return HO_ERR_BYTES_SHORT;                       /* "slice too short to read tree length" */
    size_t tree_len = ((size_t)bytes[1] << 24) | ((size_t)bytes[2] << 16) | ((size_t)bytes[3] << 8) | bytes[4];
    if (tree_len < 2) return HO_ERR_TREE_LEN;                   /* comp.rs:153-155 */
    if (n < 5 + tree_len) return HO_ERR_BYTES_SHORT;            /* "slice too short to read tree" */
    size_t tree_bits = tree_len * 8;
    /* comp.rs:164: `for _ in 0..tree_padding_bits { b.pop(); }` */
    tree_bits = tree_bits >= tree_pad ? tree_bits - tree_pad : 0;
    if (ho_tree_from_bin(bytes + 5, tree_bits, t) != HO_OK)
        return HO_ERR_INVALID_TREE;                             /* comp.rs:172-176 */
    /* comp.rs:180: `bytes.get(5 + tree_len..)` succeeds with an empty slice when n == 5 + tree_len;
     * CompressData::new then panics on empty comp_bytes (comp.rs:56-58). */
    size_t dlen = n - 5 - tree_len;
    if (dlen == 0) return HO_ERR_EMPTY_COMP;
    if (data_pad > 7) return HO_ERR_BAD_PADDING;                /* comp.rs:59-61 */
    *data_off = 5 + tree_len; *data_len = dlen; *padding_bits = data_pad;
    return HO_OK;
}
