"""ctypes front-end of the C oracle (oracle/huff_oracle.c) -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (huff_encoding_b200) never does.

Reference anchors: see huff_oracle.h (each entry point cites huff_coding/src/... file:line).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhuff_oracle.so")

HO_MAX_LEAVES = 257
HO_MAX_NODES = 2 * HO_MAX_LEAVES - 1
HO_NONE = 0xFFFF
HO_CODE_BYTES = 33

ORDER_ASC = 0
ORDER_BYTEWEIGHTS = 1

OK = 0
ERR_EMPTY_WEIGHTS = 1
ERR_MISSING_LETTER = 2
ERR_EMPTY_COMP = 3
ERR_BAD_PADDING = 4
ERR_CAPACITY = 5
ERR_BIN_TOO_SMALL = 6
ERR_BIN_TOO_BIG = 7
ERR_BYTES_SHORT = 8
ERR_TREE_LEN = 9
ERR_INVALID_TREE = 10


class HoNode(C.Structure):
    _fields_ = [("left", C.c_uint16), ("right", C.c_uint16), ("letter", C.c_uint8), ("weight", C.c_uint64)]


class HoTree(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32),
        ("root", C.c_uint32),
        ("nodes", HoNode * HO_MAX_NODES),
        ("has_code", C.c_uint8 * 256),
        ("code_len", C.c_uint16 * 256),
        ("code_bits", (C.c_uint8 * HO_CODE_BYTES) * 256),
    ]

    # ---- convenience views used by the tests
    def code_str(self, letter: int) -> str | None:
        """Code of `letter` as a '0'/'1' string (None when the tree has no such letter)."""
        if not self.has_code[letter]:
            return None
        n = self.code_len[letter]
        bits = self.code_bits[letter]
        return "".join("1" if (bits[k >> 3] >> (7 - (k & 7))) & 1 else "0" for k in range(n))

    def codes(self) -> dict[int, str]:
        return {b: self.code_str(b) for b in range(256) if self.has_code[b]}

    def lens(self) -> np.ndarray:
        return np.array([self.code_len[b] if self.has_code[b] else 0 for b in range(256)], dtype=np.int64)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms).  Building the checker is not using it."""
    src = os.path.join(_HERE, "huff_oracle.c")
    hdr = os.path.join(_HERE, "huff_oracle.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libhuff_oracle.so"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u64p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_size_t)
        tp = C.POINTER(HoTree)
        L.ho_histogram.argtypes = [C.c_void_p, C.c_size_t, u64p]
        L.ho_histogram.restype = None
        L.ho_tree_from_pairs.argtypes = [u8p, u64p, C.c_size_t, tp]
        L.ho_tree_from_weights.argtypes = [u64p, C.c_int, tp]
        L.ho_compress_with_tree.argtypes = [C.c_void_p, C.c_size_t, tp, C.c_void_p, C.c_size_t, szp, u8p, u8p]
        L.ho_compressed_bits.argtypes = [C.c_void_p, C.c_size_t, tp, u64p, u8p]
        L.ho_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, tp, C.c_void_p, C.c_size_t, szp, u8p]
        L.ho_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, tp, C.c_void_p, C.c_size_t, szp]
        L.ho_decompress_count.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, tp, szp]
        L.ho_tree_as_bin.argtypes = [tp, C.c_void_p, C.c_size_t, szp]
        L.ho_tree_from_bin.argtypes = [C.c_void_p, C.c_size_t, tp]
        L.ho_to_bytes.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, tp, C.c_void_p, C.c_size_t, szp]
        L.ho_try_from_bytes.argtypes = [C.c_void_p, C.c_size_t, tp, szp, szp, u8p]
        L.ho_calc_padding_bits.argtypes = [C.c_uint64]
        L.ho_calc_padding_bits.restype = C.c_uint8
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int, missing: int | None = None):
        super().__init__(f"oracle status {code}" + (f" (missing letter {missing})" if missing is not None else ""))
        self.code = code
        self.missing = missing


def _as_u8(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.asarray(data)
    if a.dtype != np.uint8:
        raise TypeError("u8 data expected")
    return np.ascontiguousarray(a)


def histogram(data) -> np.ndarray:
    a = _as_u8(data)
    w = np.zeros(256, dtype=np.uint64)
    lib().ho_histogram(a.ctypes.data, a.size, w.ctypes.data_as(C.POINTER(C.c_uint64)))
    return w


def tree_from_weights(w, order: int = ORDER_ASC) -> HoTree:
    w = np.ascontiguousarray(np.asarray(w, dtype=np.uint64))
    assert w.shape == (256,)
    t = HoTree()
    rc = lib().ho_tree_from_weights(w.ctypes.data_as(C.POINTER(C.c_uint64)), order, C.byref(t))
    if rc:
        raise OracleError(rc)
    return t


def tree_from_pairs(letters, weights) -> HoTree:
    le = np.ascontiguousarray(np.asarray(letters, dtype=np.uint8))
    we = np.ascontiguousarray(np.asarray(weights, dtype=np.uint64))
    t = HoTree()
    rc = lib().ho_tree_from_pairs(le.ctypes.data_as(C.POINTER(C.c_uint8)), we.ctypes.data_as(C.POINTER(C.c_uint64)),
                                  le.size, C.byref(t))
    if rc:
        raise OracleError(rc)
    return t


def compress_bound(n: int, tree: HoTree) -> int:
    mx = max([tree.code_len[b] for b in range(256) if tree.has_code[b]] + [1])
    return (n * mx + 7) // 8 + 8


def compress_with_tree(data, tree: HoTree) -> tuple[np.ndarray, int]:
    """-> (comp_bytes, padding_bits); raises OracleError(ERR_MISSING_LETTER, letter)."""
    a = _as_u8(data)
    bits = C.c_uint64(0)
    missing = C.c_uint8(0)
    rc = lib().ho_compressed_bits(a.ctypes.data, a.size, C.byref(tree), C.byref(bits), C.byref(missing))
    if rc:
        raise OracleError(rc, missing.value)
    cap = (bits.value + 7) // 8 + 8
    out = np.empty(cap, dtype=np.uint8)
    out_len, pad = C.c_size_t(0), C.c_uint8(0)
    rc = lib().ho_compress_with_tree(a.ctypes.data, a.size, C.byref(tree), out.ctypes.data, cap,
                                     C.byref(out_len), C.byref(pad), C.byref(missing))
    if rc:
        raise OracleError(rc, missing.value)
    return out[: out_len.value].copy(), pad.value


def compress(data, order: int = ORDER_ASC) -> tuple[np.ndarray, int, HoTree]:
    """-> (comp_bytes, padding_bits, tree) as huff_coding::compress would (comp.rs:353-356)."""
    a = _as_u8(data)
    t = tree_from_weights(histogram(a), order)
    comp, pad = compress_with_tree(a, t)
    return comp, pad, t


def decompress(comp, padding_bits: int, tree: HoTree) -> np.ndarray:
    """One bit-serial walk (comp.rs:487-519).  The reference grows a Vec (comp.rs:491); here the output is sized for
    the worst case up front (every letter = the tree's shortest code), which costs no time: untouched pages are never
    committed."""
    a = _as_u8(comp)
    if a.size == 0 or padding_bits > 7:
        n0 = C.c_size_t(0)
        raise OracleError(lib().ho_decompress_count(a.ctypes.data if a.size else None, a.size, padding_bits,
                                                    C.byref(tree), C.byref(n0)) or ERR_EMPTY_COMP)
    lens = [tree.code_len[b] for b in range(256) if tree.has_code[b]]
    shortest = max(1, min(lens)) if lens else 1
    cap = (a.size * 8 - padding_bits) // shortest + 1
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = lib().ho_decompress(a.ctypes.data, a.size, padding_bits, C.byref(tree), out.ctypes.data, out.size, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].copy()


def tree_as_bin(tree: HoTree) -> tuple[np.ndarray, int]:
    """-> (bytes with dead bits zero, n_bits)  (tree_inner.rs:632-668)"""
    out = np.zeros(HO_MAX_LEAVES * 10 // 8 + 16, dtype=np.uint8)
    nb = C.c_size_t(0)
    rc = lib().ho_tree_as_bin(C.byref(tree), out.ctypes.data, out.size, C.byref(nb))
    if rc:
        raise OracleError(rc)
    return out[: (nb.value + 7) // 8].copy(), nb.value


def bin_to_string(bin_bytes, n_bits: int) -> str:
    """bitvec 0.20 `BitVec<Msb0,u8>::to_string()` layout: '[10011000, 11100110, 00010]'."""
    b = _as_u8(bin_bytes)
    s = "".join(f"{x:08b}" for x in b)[:n_bits]
    return "[" + ", ".join(s[i:i + 8] for i in range(0, n_bits, 8)) + "]"


def tree_from_bin(bin_bytes, n_bits: int) -> HoTree:
    b = _as_u8(bin_bytes)
    t = HoTree()
    rc = lib().ho_tree_from_bin(b.ctypes.data if b.size else None, n_bits, C.byref(t))
    if rc:
        raise OracleError(rc)
    return t


def to_bytes(comp, padding_bits: int, tree: HoTree) -> np.ndarray:
    a = _as_u8(comp)
    out = np.empty(a.size + 512, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = lib().ho_to_bytes(a.ctypes.data, a.size, padding_bits, C.byref(tree), out.ctypes.data, out.size, C.byref(n))
    if rc:
        raise OracleError(rc)
    return out[: n.value].copy()


def try_from_bytes(blob) -> tuple[np.ndarray, int, HoTree]:
    a = _as_u8(blob)
    t = HoTree()
    off, ln, pad = C.c_size_t(0), C.c_size_t(0), C.c_uint8(0)
    rc = lib().ho_try_from_bytes(a.ctypes.data if a.size else None, a.size, C.byref(t), C.byref(off), C.byref(ln), C.byref(pad))
    if rc:
        raise OracleError(rc)
    return a[off.value: off.value + ln.value].copy(), pad.value, t
