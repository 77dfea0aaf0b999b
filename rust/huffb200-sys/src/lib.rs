//! Raw bindings of include/huffb200.h (hand-written to mirror the header 1:1).
//! NOTE: never compiled in this repository (no rustc in the build image); kept next to the header it mirrors.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const HB_MAX_LEAVES: usize = 257;
pub const HB_MAX_NODES: usize = 2 * HB_MAX_LEAVES - 1;
pub const HB_NO_CHILD: u16 = 0xFFFF;

pub const HB_OK: c_int = 0;
pub const HB_ERR_EMPTY_WEIGHTS: c_int = 1;
pub const HB_ERR_MISSING_LETTER: c_int = 2;
pub const HB_ERR_EMPTY_COMP: c_int = 3;
pub const HB_ERR_BAD_PADDING: c_int = 4;
pub const HB_ERR_CAPACITY: c_int = 5;
pub const HB_ERR_BIN_TOO_SMALL: c_int = 6;
pub const HB_ERR_BIN_TOO_BIG: c_int = 7;
pub const HB_ERR_BYTES_SHORT: c_int = 8;
pub const HB_ERR_TREE_LEN: c_int = 9;
pub const HB_ERR_INVALID_TREE: c_int = 10;
pub const HB_ERR_CUDA: c_int = 11;
pub const HB_ERR_INVALID_ARG: c_int = 12;
pub const HB_ERR_CODE_TOO_LONG: c_int = 13;
pub const HB_ERR_NO_MEM: c_int = 14;
pub const HB_ERR_TREE_NODES: c_int = 15;
pub const HB_COMM_ID_BYTES: usize = 128;

pub const HB_ORDER_ASC: c_int = 0;
pub const HB_ORDER_BYTEWEIGHTS: c_int = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct hb_node {
    pub left: u16,
    pub right: u16,
    pub letter: u8,
    pub reserved: [u8; 3],
    pub weight: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct hb_tree {
    pub n_nodes: u32,
    pub root: u32,
    pub n_leaves: u32,
    pub max_len: u32,
    pub min_len: u32,
    pub len_gcd: u32,
    pub nodes: [hb_node; HB_MAX_NODES],
    pub has_code: [u8; 256],
    pub code_len: [u16; 256],
    pub code: [u64; 256],
}

#[repr(C)]
pub struct hb_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn hb_status_str(status: c_int) -> *const c_char;
    pub fn hb_last_error() -> *const c_char;
    pub fn hb_ctx_create(device: c_int, ctx: *mut *mut hb_ctx) -> c_int;
    pub fn hb_ctx_destroy(ctx: *mut hb_ctx) -> c_int;
    pub fn hb_free(p: *mut c_void);

    pub fn hb_tree_from_weights(weights: *const u64, order_mode: c_int, tree: *mut hb_tree) -> c_int;
    pub fn hb_tree_from_pairs(letters: *const u8, weights: *const u64, n: usize, tree: *mut hb_tree) -> c_int;
    pub fn hb_tree_as_bin(tree: *const hb_tree, out: *mut u8, cap_bytes: usize, n_bits: *mut usize) -> c_int;
    pub fn hb_tree_from_bin(bin: *const u8, n_bits: usize, tree: *mut hb_tree) -> c_int;
    pub fn hb_to_bytes(comp: *const u8, comp_len: usize, padding_bits: u8, tree: *const hb_tree,
                       out: *mut u8, cap: usize, out_len: *mut usize) -> c_int;
    pub fn hb_try_from_bytes(bytes: *const u8, n: usize, tree: *mut hb_tree, data_off: *mut usize,
                             data_len: *mut usize, padding_bits: *mut u8) -> c_int;

    pub fn hb_histogram_u8(ctx: *mut hb_ctx, data: *const u8, n: usize, out: *mut u64) -> c_int;
    pub fn hb_compress_u8(ctx: *mut hb_ctx, data: *const u8, n: usize, order_mode: c_int, tree_out: *mut hb_tree,
                          comp_bytes: *mut *mut u8, comp_len: *mut usize, padding_bits: *mut u8) -> c_int;
    pub fn hb_compress_with_tree_u8(ctx: *mut hb_ctx, data: *const u8, n: usize, tree: *const hb_tree,
                                    comp_bytes: *mut *mut u8, comp_len: *mut usize, padding_bits: *mut u8,
                                    missing: *mut u8) -> c_int;
    pub fn hb_decompress_u8(ctx: *mut hb_ctx, comp: *const u8, comp_len: usize, padding_bits: u8,
                            tree: *const hb_tree, out: *mut *mut u8, out_n: *mut usize) -> c_int;

    // caller-owned (ideally pinned, hb_host_alloc) buffers: no allocation inside the call
    pub fn hb_host_alloc(bytes: usize, p: *mut *mut c_void) -> c_int;
    pub fn hb_host_free(p: *mut c_void);
    pub fn hb_compress_u8_into(ctx: *mut hb_ctx, data: *const u8, n: usize, order_mode: c_int, tree_out: *mut hb_tree,
                               comp_bytes: *mut u8, comp_cap: usize, comp_len: *mut usize, padding_bits: *mut u8) -> c_int;
    pub fn hb_compress_with_tree_u8_into(ctx: *mut hb_ctx, data: *const u8, n: usize, tree: *const hb_tree,
                                         comp_bytes: *mut u8, comp_cap: usize, comp_len: *mut usize,
                                         padding_bits: *mut u8, missing: *mut u8) -> c_int;
    pub fn hb_decompress_u8_into(ctx: *mut hb_ctx, comp: *const u8, comp_len: usize, padding_bits: u8,
                                 tree: *const hb_tree, out: *mut u8, out_cap: usize, out_n: *mut usize) -> c_int;

    // context helpers
    pub fn hb_version() -> c_int;
    pub fn hb_ctx_sync(ctx: *mut hb_ctx) -> c_int;
    pub fn hb_ctx_stream(ctx: *mut hb_ctx) -> *mut c_void;
    pub fn hb_ctx_kernel_launches(ctx: *mut hb_ctx, count: *mut u64) -> c_int;
    pub fn hb_ctx_last_decode_repairs(ctx: *mut hb_ctx, count: *mut u32) -> c_int;

    // device-buffer API (pointers are CUDA device pointers; everything is enqueued on hb_ctx_stream)
    pub fn hb_histogram_u8_dev(ctx: *mut hb_ctx, d_data: *const u8, n: usize, d_hist256: *mut u64) -> c_int;
    pub fn hb_stream_bits(weights: *const u64, tree: *const hb_tree, bits: *mut u64, missing: *mut u8) -> c_int;
    pub fn hb_shard_plan(hists: *const u64, n_shards: usize, order_mode: c_int, tree_out: *mut hb_tree,
                         shard_bits: *mut u64) -> c_int;
    pub fn hb_encode_u8_dev(ctx: *mut hb_ctx, d_data: *const u8, n: usize, tree: *const hb_tree, start_bit: u32,
                            d_out: *mut u8, out_cap: usize, d_total_bits: *mut u64) -> c_int;
    pub fn hb_compress_u8_dev(ctx: *mut hb_ctx, d_data: *const u8, n: usize, order_mode: c_int, tree_out: *mut hb_tree,
                              d_out: *mut u8, out_cap: usize, comp_len: *mut usize, padding_bits: *mut u8) -> c_int;
    pub fn hb_decompress_u8_dev(ctx: *mut hb_ctx, d_comp: *const u8, comp_len: usize, padding_bits: u8,
                                tree: *const hb_tree, d_out: *mut u8, out_cap: usize, out_n: *mut usize) -> c_int;
    pub fn hb_decode_count_dev(ctx: *mut hb_ctx, d_buf: *const u8, avail_bits: u64, own_begin: u64, own_end: u64,
                               stream_bit0: u64, tree: *const hb_tree, info: *mut hb_shard_info) -> c_int;
    pub fn hb_decode_write_dev(ctx: *mut hb_ctx, d_out: *mut u8, out_cap: usize) -> c_int;
    pub fn hb_decode_shard_dev(ctx: *mut hb_ctx, d_buf: *const u8, avail_bits: u64, own_begin: u64, own_end: u64,
                               stream_bit0: u64, tree: *const hb_tree, info: *mut hb_shard_info, d_out: *mut u8,
                               out_cap: usize) -> c_int;
    pub fn hb_ctx_last_decode_path(ctx: *mut hb_ctx, fused: *mut u32, slow_chunks: *mut u32) -> c_int;
    pub fn hb_ctx_last_encode_error(ctx: *mut hb_ctx, flag: *mut u32) -> c_int;
    pub fn hb_ctx_fused_phase_cycles(ctx: *mut hb_ctx, out: *mut u64) -> c_int;

    // multi-GPU inside the library: one rank per ctx, NCCL communicator owned by the library (loaded at run time)
    pub fn hb_comm_get_unique_id(id: *mut u8) -> c_int;
    pub fn hb_comm_init(ctx: *mut hb_ctx, n_ranks: c_int, rank: c_int, id: *const u8) -> c_int;
    pub fn hb_comm_finalize(ctx: *mut hb_ctx) -> c_int;
    pub fn hb_compress_shard_dev(ctx: *mut hb_ctx, d_data: *const u8, n: usize, order_mode: c_int, tree_out: *mut hb_tree,
                                 d_out: *mut u8, out_cap: usize, layout: *mut hb_shard_layout) -> c_int;
    pub fn hb_decompress_shard_dev(ctx: *mut hb_ctx, d_comp: *const u8, layout: *const hb_shard_layout,
                                   tree: *const hb_tree, d_out: *mut u8, out_cap: usize, out_n: *mut usize) -> c_int;
}

/// Where a shard sits in the whole stream (include/huffb200.h: hb_shard_layout)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hb_shard_layout {
    pub bit_offset: u64,
    pub bits: u64,
    pub total_bits: u64,
    pub comp_len: usize,
    pub start_bit: u32,
    pub padding_bits: u8,
}

/// Sharded decode result (include/huffb200.h: hb_shard_info)
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct hb_shard_info {
    pub entry_bit: i64,
    pub exit_bit: u64,
    pub n_letters: u64,
}
