/*
 * huff_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the u8 hot path of the Rust crate `huff_coding`
 * (k-xlsx/huff-encoding): frequency count -> Huffman tree -> code table ->
 * compress_with_tree / decompress, plus the tree / container serialisation.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may link or call this.  The product (huff_encoding_b200/) never
 * does; it fails loudly when its CUDA library is missing.
 *
 * PINNING STATUS.  The reference cannot be compiled here (no rustc/cargo in the
 * image, no network), so there is no oracle/_ref.  The oracle is pinned against
 * every golden vector the reference's own tests and doctests hold for this path
 * (see tests/test_oracle_golden.py).  One
 * sub-behaviour stays "parity unpinned": the tie-break order of equal weights in
 * a heap of more than three live elements.  It is inherited from Rust std's
 * `BinaryHeap` (push = sift_up, pop = swap-last-to-root + sift_down_to_bottom +
 * sift_up), which lives outside /root/reference; it is restated here from std's
 * published algorithm and no reference test exercises it.
 */
#ifndef HUFF_ORACLE_H
#define HUFF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HO_MAX_LEAVES 257          /* 256 bytes + the duplicate byte-0 leaf of ByteWeights (weights.rs:396-415) */
#define HO_MAX_NODES  (2 * HO_MAX_LEAVES - 1)
#define HO_NONE       0xFFFFu
#define HO_CODE_BYTES 33           /* up to 256+ bits per code, MSB-first */

/* insertion order of the leaves into the heap (branch_heap.rs:52-58) */
#define HO_ORDER_ASC          0    /* ascending byte value, every non-zero bin once (canonical order) */
#define HO_ORDER_BYTEWEIGHTS  1    /* ByteWeights iterator incl. its wrap-around quirk (weights.rs:396-415) */

#define HO_OK                 0
#define HO_ERR_EMPTY_WEIGHTS  1    /* tree_inner.rs:283-285 panic "provided empty weights" */
#define HO_ERR_MISSING_LETTER 2    /* comp.rs:426-432 CompressError */
#define HO_ERR_EMPTY_COMP     3    /* comp.rs:56-58 panic "provided comp_bytes are empty" */
#define HO_ERR_BAD_PADDING    4    /* comp.rs:59-61 panic "padding bits cannot be larger than 7" */
#define HO_ERR_CAPACITY       5
#define HO_ERR_BIN_TOO_SMALL  6    /* tree_inner.rs:532-534,556-558 */
#define HO_ERR_BIN_TOO_BIG    7    /* tree_inner.rs:586-590 */
#define HO_ERR_BYTES_SHORT    8    /* comp.rs:143,149,161,180 */
#define HO_ERR_TREE_LEN       9    /* comp.rs:153-155 panic */
#define HO_ERR_INVALID_TREE   10   /* comp.rs:172-176 */

typedef struct {
    uint16_t left, right;   /* child node indices, HO_NONE for a letter branch (branch.rs:158-162) */
    uint8_t  letter;        /* valid when left == HO_NONE (leaf.rs:25-29) */
    uint64_t weight;        /* leaf.rs:27 (usize) */
} ho_node;

typedef struct {
    uint32_t n_nodes;
    uint32_t root;
    ho_node  nodes[HO_MAX_NODES];
    /* code table as read_codes() returns it (tree_inner.rs:388-419): duplicate letters -> last DFS visit wins */
    uint8_t  has_code[256];
    uint16_t code_len[256];
    uint8_t  code_bits[256][HO_CODE_BYTES];   /* bit k of the code = (code_bits[k/8] >> (7-k%8)) & 1 */
} ho_tree;

/* weights.rs:116-123 (build_weights_map) / weights.rs:265-279 (ByteWeights::from_bytes) */
void ho_histogram(const uint8_t *data, size_t n, uint64_t w[256]);

/* branch_heap.rs:24-58 + tree_inner.rs:281-320,422-440 with an explicit insertion order:
 * leaf i is (letters[i], weights[i]); duplicates allowed. */
int ho_tree_from_pairs(const uint8_t *letters, const uint64_t *weights, size_t n, ho_tree *t);

/* Same, leaves taken from a 256-bin histogram in the given order mode. */
int ho_tree_from_weights(const uint64_t w[256], int order_mode, ho_tree *t);

/* comp.rs:419-451.  out must hold ho_compress_bound() bytes.  *missing = first letter without a code. */
int ho_compress_with_tree(const uint8_t *data, size_t n, const ho_tree *t,
                          uint8_t *out, size_t cap, size_t *out_len, uint8_t *padding_bits, uint8_t *missing);
/* total bits of the stream, and first missing letter check, without writing */
int ho_compressed_bits(const uint8_t *data, size_t n, const ho_tree *t, uint64_t *bits, uint8_t *missing);

/* comp.rs:353-356 with the given leaf order */
int ho_compress(const uint8_t *data, size_t n, int order_mode, ho_tree *t,
                uint8_t *out, size_t cap, size_t *out_len, uint8_t *padding_bits);

/* comp.rs:487-519.  *out_n = number of letters decoded; fails with HO_ERR_CAPACITY if cap is too small. */
int ho_decompress(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t,
                  uint8_t *out, size_t cap, size_t *out_n);
/* count only (no output buffer) */
int ho_decompress_count(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t, size_t *out_n);

/* tree_inner.rs:632-668: preorder, joint = 1, letter = 0 + 8 bits.  bits out MSB-first, dead bits zero. */
int ho_tree_as_bin(const ho_tree *t, uint8_t *out, size_t cap, size_t *n_bits);
/* tree_inner.rs:522-604 */
int ho_tree_from_bin(const uint8_t *bin, size_t n_bits, ho_tree *t);

/* comp.rs:279-300 / comp.rs:128-184 */
int ho_to_bytes(const uint8_t *comp, size_t len, uint8_t padding_bits, const ho_tree *t,
                uint8_t *out, size_t cap, size_t *out_len);
int ho_try_from_bytes(const uint8_t *bytes, size_t n, ho_tree *t,
                      size_t *data_off, size_t *data_len, uint8_t *padding_bits);

/* utils.rs:37-40 */
uint8_t ho_calc_padding_bits(uint64_t bit_count);

#ifdef __cplusplus
}
#endif
#endif
