/*
 * huffb200.h -- C ABI of libhuffb200.so: the B200-native (sm_100a) implementation of huff_coding's
 * byte-alphabet hot path.  This is the drop-in boundary a Rust `-sys` crate (or any FFI) binds; see
 * INTEGRATION.md for the reference-side binding.  Plain pointers and sizes only, no torch types.
 *
 * Each entry point cites the reference interface it replaces (paths relative to /root/reference/).
 *
 * Conventions
 *   - every function returns an hb_status; HB_OK == 0.  Reference panics / Err values map to status
 *     codes (table below); nothing here calls abort().
 *   - *_u8 functions take HOST pointers, are synchronous, and own the host<->device copies.
 *   - *_dev functions take DEVICE pointers (cudaMalloc'ed on the ctx's device, 16-byte aligned),
 *     run on the ctx's stream and synchronise only where the algorithm needs a host step
 *     (tree construction, output sizing); they are what bench.py times with inputs resident in HBM.
 *   - an hb_ctx is bound to one CUDA device and is not thread-safe: one ctx per calling thread / rank.
 *   - there is NO CPU fallback: without a usable CUDA device hb_ctx_create fails with HB_ERR_CUDA.
 *   - buffers returned through `uint8_t **` are allocated by the library and released with hb_free().
 */
#ifndef HUFFB200_H
#define HUFFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_VERSION_MAJOR 0
#define HB_VERSION_MINOR 1

#define HB_MAX_LEAVES 257              /* 256 bytes + ByteWeights' duplicate byte-0 leaf (weights.rs:396-415) */
#define HB_MAX_NODES  (2 * HB_MAX_LEAVES - 1)
#define HB_NO_CHILD   0xFFFFu
#define HB_MAX_ENCODE_BITS 64          /* longest code the GPU encoder packs (needs N > 2^44 letters to exceed) */

typedef enum hb_status {
    HB_OK                 = 0,
    HB_ERR_EMPTY_WEIGHTS  = 1,   /* panic "provided empty weights"            tree_inner.rs:283-285 */
    HB_ERR_MISSING_LETTER = 2,   /* Err(CompressError{missing_letter})        comp.rs:426-432 */
    HB_ERR_EMPTY_COMP     = 3,   /* panic "provided comp_bytes are empty"     comp.rs:56-58 */
    HB_ERR_BAD_PADDING    = 4,   /* panic "padding bits cannot be larger than 7"  comp.rs:59-61 */
    HB_ERR_CAPACITY       = 5,   /* caller-provided buffer too small (needed size is reported) */
    HB_ERR_BIN_TOO_SMALL  = 6,   /* FromBinError "...too small..."            tree_inner.rs:532-534,556-558 */
    HB_ERR_BIN_TOO_BIG    = 7,   /* FromBinError "...too big..."              tree_inner.rs:586-590 */
    HB_ERR_BYTES_SHORT    = 8,   /* CompressedDataFromBytesError              comp.rs:143,149,161 */
    HB_ERR_TREE_LEN       = 9,   /* panic "stored tree length must be at least 2"  comp.rs:153-155 */
    HB_ERR_INVALID_TREE   = 10,  /* CompressedDataFromBytesError "invalid tree in slice"  comp.rs:172-176 */
    HB_ERR_CUDA           = 11,  /* CUDA runtime failure; hb_last_error() has the text */
    HB_ERR_INVALID_ARG    = 12,
    HB_ERR_CODE_TOO_LONG  = 13,  /* a used letter's code exceeds HB_MAX_ENCODE_BITS */
    HB_ERR_NO_MEM         = 14,
    HB_ERR_TREE_NODES     = 15   /* hb_tree_from_bin: more than HB_MAX_NODES nodes.  The reference builds a boxed tree of
                                    any size (tree_inner.rs:522-604); hb_tree is a flat array sized for a byte alphabet
                                    (256 letters + ByteWeights' duplicate byte-0 leaf).  A foreign tree with more than
                                    257 leaves (duplicates of duplicates) is refused with this status, never truncated. */
} hb_status;

/* leaf insertion order into the heap (tree/branch_heap.rs:52-58) */
#define HB_ORDER_ASC         0   /* ascending byte value: the canonical order for compress() */
#define HB_ORDER_BYTEWEIGHTS 1   /* ByteWeights' iterator, wrap-around quirk included (weights.rs:396-415) */

/* tree/branch.rs:158-162 + tree/leaf.rs:25-29, flattened */
typedef struct hb_node {
    uint16_t left, right;      /* HB_NO_CHILD for a letter branch */
    uint8_t  letter;
    uint8_t  reserved[3];
    uint64_t weight;           /* 0 for trees read with hb_tree_from_bin (tree_inner.rs:538,573) */
} hb_node;

/* tree/tree_inner.rs:193-196 HuffTree<u8> + its read_codes() table (tree_inner.rs:388-419) */
typedef struct hb_tree {
    uint32_t n_nodes;
    uint32_t root;
    uint32_t n_leaves;
    uint32_t max_len;          /* longest / shortest code among letters with a code */
    uint32_t min_len;
    uint32_t len_gcd;          /* gcd of all code lengths (decoder speculation alignment) */
    hb_node  nodes[HB_MAX_NODES];
    uint8_t  has_code[256];
    uint16_t code_len[256];
    uint64_t code[256];        /* code bits right-aligned (first bit = most significant of the len bits); 0 if len > 64 */
} hb_tree;

typedef struct hb_ctx hb_ctx;

/* ---------------------------------------------------------------- library / context */
const char *hb_status_str(int status);
const char *hb_last_error(void);                 /* thread-local text of the last HB_ERR_CUDA */
int  hb_version(void);                           /* major * 100 + minor */
hb_status hb_ctx_create(int device, hb_ctx **ctx);
hb_status hb_ctx_destroy(hb_ctx *ctx);
hb_status hb_ctx_sync(hb_ctx *ctx);
void     *hb_ctx_stream(hb_ctx *ctx);            /* the cudaStream_t every *_dev call is enqueued on */
hb_status hb_ctx_kernel_launches(hb_ctx *ctx, uint64_t *count);   /* kernels launched so far (bench.py's gpu_launches) */
/* number of 32 KiB chunks whose speculative entry was refuted and repaired in the last decode count pass (normally 0;
 * HB_DEBUG_SPOIL_SPECULATION=1 at ctx creation makes the speculation deliberately bad so tests can reach that path) */
hb_status hb_ctx_last_decode_repairs(hb_ctx *ctx, uint32_t *count);
void      hb_free(void *p);                      /* frees buffers returned by *_u8 calls */
/* pinned host staging for callers that want full PCIe speed on the *_u8 path */
hb_status hb_host_alloc(size_t bytes, void **p);
void      hb_host_free(void *p);

/* ---------------------------------------------------------------- host-only: tree + code table
 * replaces HuffTree::<u8>::from_weights (tree_inner.rs:281-320) over HuffBranchHeap (branch_heap.rs:18-83),
 * reproducing std::collections::BinaryHeap's tie-breaks, and read_codes() (tree_inner.rs:388-440). */
hb_status hb_tree_from_weights(const uint64_t weights[256], int order_mode, hb_tree *tree);
hb_status hb_tree_from_pairs(const uint8_t *letters, const uint64_t *weights, size_t n, hb_tree *tree);
/* HuffTree::as_bin / try_from_bin (tree_inner.rs:632-668 / 522-604); bits MSB-first, dead bits zero */
hb_status hb_tree_as_bin(const hb_tree *tree, uint8_t *out, size_t cap_bytes, size_t *n_bits);
hb_status hb_tree_from_bin(const uint8_t *bin, size_t n_bits, hb_tree *tree);
/* CompressData::to_bytes / try_from_bytes (comp.rs:279-300 / 128-184) */
hb_status hb_to_bytes(const uint8_t *comp, size_t comp_len, uint8_t padding_bits, const hb_tree *tree,
                      uint8_t *out, size_t cap, size_t *out_len);
hb_status hb_try_from_bytes(const uint8_t *bytes, size_t n, hb_tree *tree,
                            size_t *data_off, size_t *data_len, uint8_t *padding_bits);

/* ---------------------------------------------------------------- host-buffer API (synchronous) */
/* build_weights_map(&[u8]) (weights.rs:82-84,116-123) / ByteWeights::from_bytes (weights.rs:265-279) */
hb_status hb_histogram_u8(hb_ctx *ctx, const uint8_t *data, size_t n, uint64_t out[256]);
/* compress(&[u8]) (comp.rs:353-356): histogram + tree (order_mode) + encode.  n == 0 -> HB_ERR_EMPTY_WEIGHTS. */
hb_status hb_compress_u8(hb_ctx *ctx, const uint8_t *data, size_t n, int order_mode, hb_tree *tree_out,
                         uint8_t **comp_bytes, size_t *comp_len, uint8_t *padding_bits);
/* compress_with_tree(&[u8], HuffTree<u8>) (comp.rs:419-451).  HB_ERR_MISSING_LETTER sets *missing to the
 * first letter (in input order) the tree has no code for. */
hb_status hb_compress_with_tree_u8(hb_ctx *ctx, const uint8_t *data, size_t n, const hb_tree *tree,
                                   uint8_t **comp_bytes, size_t *comp_len, uint8_t *padding_bits, uint8_t *missing);
/* decompress(&CompressData<u8>) (comp.rs:487-519).  comp_len == 0 -> HB_ERR_EMPTY_COMP, padding > 7 ->
 * HB_ERR_BAD_PADDING (the CompressData::new invariants, comp.rs:55-61). */
hb_status hb_decompress_u8(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint8_t padding_bits,
                           const hb_tree *tree, uint8_t **out, size_t *out_n);

/* Same three calls writing into CALLER-OWNED host buffers (e.g. pinned memory that is reused across calls, which
 * is what makes the host path run at PCIe speed).  If the buffer is too small they return HB_ERR_CAPACITY and
 * report the needed size in *comp_len / *out_n. */
hb_status hb_compress_u8_into(hb_ctx *ctx, const uint8_t *data, size_t n, int order_mode, hb_tree *tree_out,
                              uint8_t *comp_bytes, size_t comp_cap, size_t *comp_len, uint8_t *padding_bits);
hb_status hb_compress_with_tree_u8_into(hb_ctx *ctx, const uint8_t *data, size_t n, const hb_tree *tree,
                                        uint8_t *comp_bytes, size_t comp_cap, size_t *comp_len, uint8_t *padding_bits,
                                        uint8_t *missing);
hb_status hb_decompress_u8_into(hb_ctx *ctx, const uint8_t *comp, size_t comp_len, uint8_t padding_bits,
                                const hb_tree *tree, uint8_t *out, size_t out_cap, size_t *out_n);

/* ---------------------------------------------------------------- device-buffer API (ctx stream) */
/* d_hist256: 256 x u64 on the device; overwritten.  No host sync. */
hb_status hb_histogram_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, uint64_t *d_hist256);
/* Exact stream size for (histogram, tree): sum w[b] * len[b].  Host arithmetic. */
hb_status hb_stream_bits(const uint64_t weights[256], const hb_tree *tree, uint64_t *bits, uint8_t *missing);
/* Multi-GPU plan, host arithmetic: from the G gathered shard histograms (hists[g*256 + b]) build the tree of
 * their sum (= the tree one GPU would build for the whole input) and every shard's exact bit total
 * shard_bits[g] = sum_b hists[g][b] * len[b]; the exclusive scan of shard_bits is each shard's global bit offset. */
hb_status hb_shard_plan(const uint64_t *hists, size_t n_shards, int order_mode, hb_tree *tree_out, uint64_t *shard_bits);
/* Encoder proper = compress_with_tree's packing loop (comp.rs:422-447).
 * The encoder needs the per-region histograms hb_histogram_u8_dev produces as a by-product.  If the immediately
 * preceding histogram call on this ctx was for the same (d_data, n) they are reused; otherwise the histogram kernel is
 * run again internally.  Callers that overwrite d_data IN PLACE between hb_histogram_u8_dev and hb_encode_u8_dev must
 * call hb_histogram_u8_dev again first (the library cannot see device-side writes).  Writes the stream as if it started at
 * bit `start_bit` (0..31) of d_out[0]: the first start_bit bits are left 0 so a neighbouring shard can be OR-ed in
 * (multi-GPU concatenation).  d_out must hold out_cap >= ceil((start_bit + bits)/8) rounded up to 4 bytes.
 * d_total_bits (device u64, optional) receives the number of code bits written.  If the input holds a letter the tree
 * has no code of 1..64 bits for, the stream is UNDEFINED (callers check with hb_stream_bits first); the warp encoder
 * records that case and hb_ctx_last_encode_error reports it.  No host sync. */
hb_status hb_encode_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, const hb_tree *tree, uint32_t start_bit,
                           uint8_t *d_out, size_t out_cap, uint64_t *d_total_bits);
/* synchronises the ctx stream; *flag != 0: the last hb_encode_u8_dev met a letter without a usable code */
hb_status hb_ctx_last_encode_error(hb_ctx *ctx, uint32_t *flag);
/* compress() on device buffers: histogram -> (sync) host tree -> encode.  *comp_len / *padding_bits are exact. */
hb_status hb_compress_u8_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, int order_mode, hb_tree *tree_out,
                             uint8_t *d_out, size_t out_cap, size_t *comp_len, uint8_t *padding_bits);
/* decompress() on device buffers.  Decodes the bits [first_bit, total_bits) of d_comp where
 * total_bits = 8 * comp_len - padding_bits; d_comp must be readable up to comp_len rounded up to 16 bytes.
 * If out_cap is too small returns HB_ERR_CAPACITY with *out_n = needed letters.  d_out[0 .. out_cap) is the decoder's to
 * use: on success the first *out_n bytes are the letters and the bytes behind them are UNDEFINED (the one-pass decoder
 * writes whole 32-byte rows and, when its speculation is refuted, a discarded attempt may have touched any of them);
 * with HB_ERR_CAPACITY the whole buffer is undefined.  Nothing outside [d_out, d_out + out_cap) is ever written. */
hb_status hb_decompress_u8_dev(hb_ctx *ctx, const uint8_t *d_comp, size_t comp_len, uint8_t padding_bits,
                               const hb_tree *tree, uint8_t *d_out, size_t out_cap, size_t *out_n);

/* ---------------------------------------------------------------- sharded decode building blocks (multi-GPU)
 * A rank owning stream bits [own_begin, own_end) of a buffer that also holds a halo on both sides
 * (buffer bit 0 .. avail_bits) runs the count phase with entry_bit < 0 (speculative: entry found by
 * self-synchronisation from the halo) or with a known entry; it reports where its last code word ends
 * (*exit_bit, relative to the buffer) and how many letters start in its range.  stream_bit0 is the position of
 * buffer bit 0 in the whole stream (it fixes the phase for code sets whose lengths share a factor). */
typedef struct hb_shard_info {
    int64_t  entry_bit;     /* in: >= 0 known first code-word start, < 0 speculative.  out: entry used */
    uint64_t exit_bit;      /* out: first code-word start >= own_end (may be > avail_bits at the stream end) */
    uint64_t n_letters;     /* out: letters whose code word starts in [entry, own_end) and ends <= avail_bits */
} hb_shard_info;
hb_status hb_decode_count_dev(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin,
                              uint64_t own_end, uint64_t stream_bit0, const hb_tree *tree, hb_shard_info *info);
/* writes the letters counted by the matching hb_decode_count_dev call (same ctx, same arguments) */
hb_status hb_decode_write_dev(hb_ctx *ctx, uint8_t *d_out, size_t out_cap);
/* Count + write in one call for a shard whose first code-word start is KNOWN (info->entry_bit >= 0 on entry), e.g. the
 * shards hb_encode_u8_dev produced (entry = start_bit): takes the one-pass fused decoder (every code word decoded once,
 * stream read once) when the tree allows it, else the two calls above.  HB_ERR_CAPACITY reports info->n_letters and
 * leaves the count state for hb_decode_write_dev.  d_out[0 .. out_cap): as for hb_decompress_u8_dev (bytes behind the
 * letters are undefined; nothing outside the buffer is written). */
hb_status hb_decode_shard_dev(hb_ctx *ctx, const uint8_t *d_buf, uint64_t avail_bits, uint64_t own_begin,
                              uint64_t own_end, uint64_t stream_bit0, const hb_tree *tree, hb_shard_info *info,
                              uint8_t *d_out, size_t out_cap);
/* which decoder the last decompress / decode_shard call ran: *fused = 0 two-pass (or fixed-length translation),
 * 1 fused single kernel, 2 fused refuted and redone two-pass; *slow_chunks = threads of the fused kernel that wrote their
 * rows letter by letter (end of the stream, ragged head).  HB_NO_FUSED=1 at ctx creation disables the fused kernel. */
hb_status hb_ctx_last_decode_path(hb_ctx *ctx, uint32_t *fused, uint32_t *slow_chunks);
/* profiling aid: SM clock cycles the teams of the last fused decode spent per phase, summed over chunks
 * (stage, decode, verify, scan, look-back, compaction), out[6] = number of chunks */
hb_status hb_ctx_fused_phase_cycles(hb_ctx *ctx, uint64_t out[8]);

/* ---------------------------------------------------------------- multi-GPU inside the library (one rank per ctx)
 * The reference's only data-parallel piece is ByteWeights::threaded_from_bytes (weights.rs:293-319): split the input,
 * count the parts, reduce.  Here every rank (one process or thread per GPU) owns a contiguous shard of the input and a
 * ctx; the library owns the communicator (NCCL over NVLink, loaded at run time) and does the one exchange a compress
 * needs -- an all-gather of the G shard histograms (G x 2 KiB) on the ctx stream -- between its own kernels.
 *   rank 0:  hb_comm_get_unique_id(id), hand `id` to the other ranks (the host application's transport)
 *   all:     hb_comm_init(ctx, G, rank, id)   ...   hb_compress_shard_dev / hb_decompress_shard_dev   ...   hb_comm_finalize
 * hb_comm_init(ctx, 1, 0, NULL) is valid (single GPU, no NCCL needed). */
#define HB_COMM_ID_BYTES 128
typedef struct hb_shard_layout {
    uint64_t bit_offset;     /* where this shard's first bit sits in the whole stream */
    uint64_t bits;           /* code bits of this shard */
    uint64_t total_bits;     /* code bits of all shards */
    size_t   comp_len;       /* bytes of d_out in use: ceil((start_bit + bits) / 8); 0 for an empty shard (bits == 0) */
    uint32_t start_bit;      /* bit_offset % 8: the first start_bit bits of d_out[0] are zero (the neighbour's) */
    uint8_t  padding_bits;   /* of the whole stream (comp.rs:446) */
} hb_shard_layout;
hb_status hb_comm_get_unique_id(uint8_t id[HB_COMM_ID_BYTES]);
hb_status hb_comm_init(hb_ctx *ctx, int n_ranks, int rank, const uint8_t id[HB_COMM_ID_BYTES]);
hb_status hb_comm_finalize(hb_ctx *ctx);
/* compress() over all ranks' shards: histogram kernel -> all-gather -> identical host tree on every rank -> encode at
 * the shard's global bit offset.  Byte (bit_offset / 8) of the whole stream is d_out[0]; concatenating the shard
 * buffers (OR-ing the byte two neighbours share) gives exactly the single-GPU stream.  Collective: every rank calls it. */
hb_status hb_compress_shard_dev(hb_ctx *ctx, const uint8_t *d_data, size_t n, int order_mode, hb_tree *tree_out,
                                uint8_t *d_out, size_t out_cap, hb_shard_layout *layout);
/* decompress() of the shard hb_compress_shard_dev produced (d_comp readable to comp_len rounded up to 16 bytes). */
hb_status hb_decompress_shard_dev(hb_ctx *ctx, const uint8_t *d_comp, const hb_shard_layout *layout, const hb_tree *tree,
                                  uint8_t *d_out, size_t out_cap, size_t *out_n);

#ifdef __cplusplus
}
#endif
#endif /* HUFFB200_H */
