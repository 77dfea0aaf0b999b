"""ctypes binding of libhuffb200.so (the C ABI declared in include/huffb200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("HUFFB200_SO") or os.path.join(HERE, "libhuffb200.so")   # override: experiment builds

HB_MAX_LEAVES = 257
HB_MAX_NODES = 2 * HB_MAX_LEAVES - 1
HB_NO_CHILD = 0xFFFF

HB_OK = 0
HB_ERR_EMPTY_WEIGHTS = 1
HB_ERR_MISSING_LETTER = 2
HB_ERR_EMPTY_COMP = 3
HB_ERR_BAD_PADDING = 4
HB_ERR_CAPACITY = 5
HB_ERR_BIN_TOO_SMALL = 6
HB_ERR_BIN_TOO_BIG = 7
HB_ERR_BYTES_SHORT = 8
HB_ERR_TREE_LEN = 9
HB_ERR_INVALID_TREE = 10
HB_ERR_CUDA = 11
HB_ERR_INVALID_ARG = 12
HB_ERR_CODE_TOO_LONG = 13
HB_ERR_NO_MEM = 14
HB_ERR_TREE_NODES = 15

HB_ORDER_ASC = 0
HB_ORDER_BYTEWEIGHTS = 1


class HbNode(C.Structure):
    _fields_ = [("left", C.c_uint16), ("right", C.c_uint16), ("letter", C.c_uint8), ("reserved", C.c_uint8 * 3),
                ("weight", C.c_uint64)]


class HbTree(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32), ("root", C.c_uint32), ("n_leaves", C.c_uint32),
        ("max_len", C.c_uint32), ("min_len", C.c_uint32), ("len_gcd", C.c_uint32),
        ("nodes", HbNode * HB_MAX_NODES),
        ("has_code", C.c_uint8 * 256),
        ("code_len", C.c_uint16 * 256),
        ("code", C.c_uint64 * 256),
    ]


class HbShardLayout(C.Structure):
    _fields_ = [("bit_offset", C.c_uint64), ("bits", C.c_uint64), ("total_bits", C.c_uint64), ("comp_len", C.c_size_t),
                ("start_bit", C.c_uint32), ("padding_bits", C.c_uint8)]


class HbShardInfo(C.Structure):
    _fields_ = [("entry_bit", C.c_int64), ("exit_bit", C.c_uint64), ("n_letters", C.c_uint64)]


# every symbol include/huffb200.h declares: (name, restype, argtypes)
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)
_treep = C.POINTER(HbTree)
_vp = C.c_void_p
SYMBOLS = [
    ("hb_status_str", C.c_char_p, [C.c_int]),
    ("hb_last_error", C.c_char_p, []),
    ("hb_version", C.c_int, []),
    ("hb_ctx_create", C.c_int, [C.c_int, C.POINTER(_vp)]),
    ("hb_ctx_destroy", C.c_int, [_vp]),
    ("hb_ctx_sync", C.c_int, [_vp]),
    ("hb_ctx_stream", _vp, [_vp]),
    ("hb_ctx_kernel_launches", C.c_int, [_vp, _u64p]),
    ("hb_ctx_last_decode_repairs", C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    ("hb_free", None, [_vp]),
    ("hb_host_alloc", C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    ("hb_host_free", None, [_vp]),
    ("hb_tree_from_weights", C.c_int, [_u64p, C.c_int, _treep]),
    ("hb_tree_from_pairs", C.c_int, [_u8p, _u64p, C.c_size_t, _treep]),
    ("hb_tree_as_bin", C.c_int, [_treep, _vp, C.c_size_t, _szp]),
    ("hb_tree_from_bin", C.c_int, [_vp, C.c_size_t, _treep]),
    ("hb_to_bytes", C.c_int, [_vp, C.c_size_t, C.c_uint8, _treep, _vp, C.c_size_t, _szp]),
    ("hb_try_from_bytes", C.c_int, [_vp, C.c_size_t, _treep, _szp, _szp, _u8p]),
    ("hb_histogram_u8", C.c_int, [_vp, _vp, C.c_size_t, _u64p]),
    ("hb_compress_u8", C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _treep, C.POINTER(_vp), _szp, _u8p]),
    ("hb_compress_with_tree_u8", C.c_int, [_vp, _vp, C.c_size_t, _treep, C.POINTER(_vp), _szp, _u8p, _u8p]),
    ("hb_decompress_u8", C.c_int, [_vp, _vp, C.c_size_t, C.c_uint8, _treep, C.POINTER(_vp), _szp]),
    ("hb_compress_u8_into", C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _treep, _vp, C.c_size_t, _szp, _u8p]),
    ("hb_compress_with_tree_u8_into", C.c_int, [_vp, _vp, C.c_size_t, _treep, _vp, C.c_size_t, _szp, _u8p, _u8p]),
    ("hb_decompress_u8_into", C.c_int, [_vp, _vp, C.c_size_t, C.c_uint8, _treep, _vp, C.c_size_t, _szp]),
    ("hb_histogram_u8_dev", C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    ("hb_stream_bits", C.c_int, [_u64p, _treep, _u64p, _u8p]),
    ("hb_shard_plan", C.c_int, [_u64p, C.c_size_t, C.c_int, _treep, _u64p]),
    ("hb_encode_u8_dev", C.c_int, [_vp, _vp, C.c_size_t, _treep, C.c_uint32, _vp, C.c_size_t, _vp]),
    ("hb_compress_u8_dev", C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _treep, _vp, C.c_size_t, _szp, _u8p]),
    ("hb_decompress_u8_dev", C.c_int, [_vp, _vp, C.c_size_t, C.c_uint8, _treep, _vp, C.c_size_t, _szp]),
    ("hb_decode_count_dev", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _treep,
                                      C.POINTER(HbShardInfo)]),
    ("hb_decode_write_dev", C.c_int, [_vp, _vp, C.c_size_t]),
    ("hb_decode_shard_dev", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _treep,
                                      C.POINTER(HbShardInfo), _vp, C.c_size_t]),
    ("hb_ctx_last_decode_path", C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("hb_ctx_fused_phase_cycles", C.c_int, [_vp, _u64p]),
    ("hb_ctx_last_encode_error", C.c_int, [_vp, C.POINTER(C.c_uint32)]),
    ("hb_comm_get_unique_id", C.c_int, [_vp]),
    ("hb_comm_init", C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    ("hb_comm_finalize", C.c_int, [_vp]),
    ("hb_compress_shard_dev", C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _treep, _vp, C.c_size_t, C.POINTER(HbShardLayout)]),
    ("hb_decompress_shard_dev", C.c_int, [_vp, _vp, C.POINTER(HbShardLayout), _treep, _vp, C.c_size_t, _szp]),
]

_lib = None


def load() -> C.CDLL:
    """Load libhuffb200.so.  Raises if it has not been built (python -m huff_encoding_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `python -m huff_encoding_b200.build` "
                              "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        if hasattr(L, "hb_emu_is_model") and os.environ.get("HB_EMU") != "1":
            # tests/emu builds a CPU model of this library for the test suite; it must never stand in for the product
            raise ImportError(f"{SO_PATH} is the CPU test model of libhuffb200, not the library (there is no CPU fallback)")
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(L, name)          # AttributeError here = the .so does not export what the header declares
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def status_str(code: int) -> str:
    return load().hb_status_str(code).decode()


def last_error() -> str:
    return load().hb_last_error().decode()
