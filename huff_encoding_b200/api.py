"""Host-side mirror of `huff_coding::prelude` for the u8 alphabet, over the C ABI of libhuffb200.so.

Same names, argument meaning and error behaviour as the reference (paths relative to
/root/reference/huff_coding/src):

    build_weights_map            weights.rs:82-84, 116-123
    ByteWeights                  weights.rs:175-443   (iterator quirk of :396-415 reproduced, see `compat`)
    HuffTree.from_weights        tree/tree_inner.rs:281-320
    HuffTree.read_codes          tree/tree_inner.rs:356-419
    HuffTree.as_bin/try_from_bin tree/tree_inner.rs:632-668 / 522-604
    compress / compress_with_tree / decompress      comp.rs:353-356 / 419-451 / 487-519
    CompressData                 comp.rs:41-89, to_bytes :279-300, try_from_bytes :128-184

Reference panics become `HuffPanic` with the reference's message; `Err` values become the error classes below.
All counting / packing / decoding runs in the CUDA kernels; there is no CPU fallback (tree construction is host
work in the reference too and stays on the host, inside the library).
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L

# ---------------------------------------------------------------- errors


class HuffPanic(RuntimeError):
    """A reference `panic!` (message identical to the reference's)."""


class CompressError(Exception):
    """comp.rs:561-590"""

    def __init__(self, message: str, missing_letter: int):
        super().__init__(f"{message} ({missing_letter})")
        self._message, self._missing = message, missing_letter

    def message(self) -> str:
        return self._message

    def missing_letter(self) -> int:
        return self._missing


class CompressedDataFromBytesError(Exception):
    """comp.rs:531-554"""

    def __init__(self, message: str):
        super().__init__(message)
        self._message = message

    def message(self) -> str:
        return self._message


class FromBinError(Exception):
    """tree_inner.rs:673-700"""


class HuffCudaError(RuntimeError):
    pass


class TreeTooLargeError(ValueError):
    """hb_tree_from_bin met a tree with more than HB_MAX_NODES nodes (HB_ERR_TREE_NODES).  The reference's try_from_bin
    (tree_inner.rs:522-604) accepts any size; a byte alphabet never produces such a tree, only hand-made input does."""


def _raise(status: int, missing: int | None = None):
    if status == L.HB_OK:
        return
    if status == L.HB_ERR_EMPTY_WEIGHTS:
        raise HuffPanic("provided empty weights")
    if status == L.HB_ERR_EMPTY_COMP:
        raise HuffPanic("provided comp_bytes are empty")
    if status == L.HB_ERR_BAD_PADDING:
        raise HuffPanic("padding bits cannot be larger than 7")
    if status == L.HB_ERR_TREE_LEN:
        raise HuffPanic("stored tree length must be at least 2")
    if status == L.HB_ERR_MISSING_LETTER:
        raise CompressError("letter not found in codes", int(missing if missing is not None else -1))
    if status == L.HB_ERR_BIN_TOO_SMALL:
        raise FromBinError("Provided BitVec is too small for an encoded HuffTree<u8>")
    if status == L.HB_ERR_BIN_TOO_BIG:
        raise FromBinError("Provided BitVec is too big for an encoded HuffTree<u8>")
    if status == L.HB_ERR_INVALID_TREE:
        raise CompressedDataFromBytesError("invalid tree in slice")
    if status == L.HB_ERR_TREE_NODES:
        raise TreeTooLargeError("foreign tree with more than 513 nodes (more than 257 leaves): beyond hb_tree's capacity")
    if status == L.HB_ERR_BYTES_SHORT:
        raise CompressedDataFromBytesError("slice too short")
    if status == L.HB_ERR_CUDA:
        raise HuffCudaError(L.last_error() or "CUDA error")
    raise RuntimeError(f"libhuffb200: {L.status_str(status)} (status {status})")


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        a = data
    elif isinstance(data, (bytes, bytearray, memoryview)):
        a = np.frombuffer(data, dtype=np.uint8)
    else:
        a = np.asarray(list(data), dtype=np.uint8)
    if a.dtype != np.uint8:
        raise TypeError("letters must be u8")
    return np.ascontiguousarray(a.reshape(-1))


# ---------------------------------------------------------------- context (one per thread, created lazily)
_tls = threading.local()


class Context:
    """An hb_ctx bound to one CUDA device."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        self._h = C.c_void_p()
        _raise(self._lib.hb_ctx_create(device, C.byref(self._h)))
        self.device = device

    @property
    def handle(self):
        return self._h

    def kernel_launches(self) -> int:
        n = C.c_uint64(0)
        _raise(self._lib.hb_ctx_kernel_launches(self._h, C.byref(n)))
        return n.value

    def sync(self):
        _raise(self._lib.hb_ctx_sync(self._h))

    def last_decode_repairs(self) -> int:
        n = C.c_uint32(0)
        _raise(self._lib.hb_ctx_last_decode_repairs(self._h, C.byref(n)))
        return n.value

    def last_decode_path(self) -> tuple[int, int]:
        """(path, slow_chunks) of the last decompress: path 0 = two-pass / fixed-length translation, 1 = fused one-pass
        kernel, 2 = fused kernel refuted and redone two-pass; slow_chunks = fused chunks that overflowed their slots."""
        f, sl = C.c_uint32(0), C.c_uint32(0)
        _raise(self._lib.hb_ctx_last_decode_path(self._h, C.byref(f), C.byref(sl)))
        return f.value, sl.value

    def stream(self) -> int:
        return int(self._lib.hb_ctx_stream(self._h) or 0)

    def close(self):
        if self._h:
            self._lib.hb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_context() -> Context:
    ctx = getattr(_tls, "ctx", None)
    if ctx is None:
        ctx = _tls.ctx = Context(0)
    return ctx


# ---------------------------------------------------------------- weights
def build_weights_map(letters, ctx: Context | None = None) -> dict[int, int]:
    """weights.rs:82-84: occurrences of every distinct byte.  Keys come out in ascending order (the canonical leaf
    order; Rust's HashMap order is random per process)."""
    a = _u8(letters)
    ctx = ctx or default_context()
    w = np.zeros(256, dtype=np.uint64)
    _raise(L.load().hb_histogram_u8(ctx.handle, a.ctypes.data, a.size, w.ctypes.data_as(C.POINTER(C.c_uint64))))
    return {b: int(w[b]) for b in range(256) if w[b]}


class ByteWeights:
    """weights.rs:175-443.  `compat=True` (default) reproduces the reference iterator bit for bit, including its
    wrap-around: when bin 255 is empty, byte 0 (if present) is yielded a second time (weights.rs:404,410)."""

    def __init__(self, compat: bool = True):
        self.weights = np.zeros(256, dtype=np.uint64)
        self._len = 0
        self.compat = compat

    @classmethod
    def from_bytes(cls, data, compat: bool = True, ctx: Context | None = None) -> "ByteWeights":
        a = _u8(data)
        bw = cls(compat)
        ctx = ctx or default_context()
        _raise(L.load().hb_histogram_u8(ctx.handle, a.ctypes.data, a.size,
                                        bw.weights.ctypes.data_as(C.POINTER(C.c_uint64))))
        bw._len = int(np.count_nonzero(bw.weights))
        return bw

    @classmethod
    def threaded_from_bytes(cls, data, thread_num: int, compat: bool = True, ctx: Context | None = None) -> "ByteWeights":
        """weights.rs:293-319.  The reference splits the input into `thread_num` rations, counts each on its own
        thread and folds the partial ByteWeights with `+=` (which, through the iterator quirk, double-counts byte 0
        of a partial when that partial has no byte 255).  With compat=True the same folds are applied to per-ration
        GPU histograms; with compat=False it equals from_bytes."""
        a = _u8(data)
        if not compat or thread_num <= 0:
            return cls.from_bytes(a, compat, ctx)
        per = a.size // thread_num                          # utils.rs:6-28 ration_vec
        if per == 0:
            rations = [a]
        else:
            rations = [a[i * per:(i + 1) * per] for i in range(thread_num - 1)] + [a[(thread_num - 1) * per:]]
        parts = [cls.from_bytes(r, compat, ctx) for r in rations]
        acc = parts.pop()                                   # weights.rs:313 `weights_vec.pop()`
        for other in parts:
            acc += other
        return acc

    def get(self, byte: int):
        w = int(self.weights[byte])
        return None if w == 0 else w

    def __len__(self):
        return self._len

    def len(self):
        return self._len

    def is_empty(self):
        return self._len == 0

    def __iter__(self):
        for b in range(256):
            if self.weights[b]:
                yield b, int(self.weights[b])
        if self.compat and self.weights[0] and not self.weights[255]:
            yield 0, int(self.weights[0])

    def iter(self):
        return iter(self)

    def add_byte_weights(self, other: "ByteWeights"):
        """weights.rs:374-387"""
        for b, f in other:
            if self.weights[b]:
                self.weights[b] += np.uint64(f)
            else:
                self.weights[b] = np.uint64(f)
                self._len += 1

    def __iadd__(self, other):
        self.add_byte_weights(other)
        return self

    def __add__(self, other):
        r = ByteWeights(self.compat)
        r.weights = self.weights.copy()
        r._len = self._len
        r.add_byte_weights(other)
        return r

    def __eq__(self, other):
        return isinstance(other, ByteWeights) and bool(np.array_equal(self.weights, other.weights))


# ---------------------------------------------------------------- tree
class HuffTree:
    """tree/tree_inner.rs:193-196 (u8 letters).  Wraps an hb_tree."""

    def __init__(self, raw: L.HbTree):
        self._t = raw

    @property
    def raw(self) -> L.HbTree:
        return self._t

    @classmethod
    def from_weights(cls, weights) -> "HuffTree":
        """tree_inner.rs:281-320.  `weights`: ByteWeights, or a mapping letter -> weight (inserted in ascending
        letter order, the canonical order), or an iterable of (letter, weight) pairs (inserted as given)."""
        lib = L.load()
        t = L.HbTree()
        if isinstance(weights, ByteWeights):
            mode = L.HB_ORDER_BYTEWEIGHTS if weights.compat else L.HB_ORDER_ASC
            w = np.ascontiguousarray(weights.weights)
            _raise(lib.hb_tree_from_weights(w.ctypes.data_as(C.POINTER(C.c_uint64)), mode, C.byref(t)))
            return cls(t)
        pairs = sorted(weights.items()) if hasattr(weights, "items") else list(weights)
        letters = np.array([p[0] for p in pairs], dtype=np.uint8)
        ws = np.array([p[1] for p in pairs], dtype=np.uint64)
        _raise(lib.hb_tree_from_pairs(letters.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      ws.ctypes.data_as(C.POINTER(C.c_uint64)), len(pairs), C.byref(t)))
        return cls(t)

    def code_str(self, letter: int):
        if not self._t.has_code[letter]:
            return None
        n = self._t.code_len[letter]
        if n > 64:
            return self._walk_codes()[letter]
        return format(self._t.code[letter], "b").zfill(n)

    def _walk_codes(self) -> dict[int, str]:
        out: dict[int, str] = {}
        t = self._t
        if t.nodes[t.root].left == L.HB_NO_CHILD:
            return {t.nodes[t.root].letter: "0"}
        stack = [(t.root, "")]
        while stack:
            n, code = stack.pop()
            nd = t.nodes[n]
            if nd.left == L.HB_NO_CHILD:
                out[nd.letter] = code
            else:
                stack.append((nd.right, code + "1"))
                stack.append((nd.left, code + "0"))
        return out

    def read_codes(self) -> dict[int, str]:
        """tree_inner.rs:356-419: letter -> code as a '0'/'1' string (bitvec Msb0 order)."""
        return {b: self.code_str(b) for b in range(256) if self._t.has_code[b]}

    def root_letter(self):
        nd = self._t.nodes[self._t.root]
        return nd.letter if nd.left == L.HB_NO_CHILD else None

    def as_bin(self) -> tuple[bytes, int]:
        """tree_inner.rs:632-668 -> (bytes, n_bits); bits MSB-first, dead bits zero."""
        out = np.zeros(L.HB_MAX_LEAVES * 10 // 8 + 16, dtype=np.uint8)
        nb = C.c_size_t(0)
        _raise(L.load().hb_tree_as_bin(C.byref(self._t), out.ctypes.data, out.size, C.byref(nb)))
        return out[: (nb.value + 7) // 8].tobytes(), nb.value

    def as_bin_string(self) -> str:
        """What bitvec 0.20's `as_bin().to_string()` prints."""
        b, n = self.as_bin()
        s = "".join(f"{x:08b}" for x in b)[:n]
        return "[" + ", ".join(s[i:i + 8] for i in range(0, n, 8)) + "]"

    @classmethod
    def try_from_bin(cls, bin_bytes, n_bits: int) -> "HuffTree":
        """tree_inner.rs:522-604"""
        a = _u8(bin_bytes)
        t = L.HbTree()
        _raise(L.load().hb_tree_from_bin(a.ctypes.data if a.size else None, n_bits, C.byref(t)))
        return cls(t)


# ---------------------------------------------------------------- CompressData
class CompressData:
    """comp.rs:41-89"""

    def __init__(self, comp_bytes, padding_bits: int, huff_tree: HuffTree):
        cb = _u8(comp_bytes)
        if cb.size == 0:
            raise HuffPanic("provided comp_bytes are empty")            # comp.rs:56-58
        if padding_bits > 7:
            raise HuffPanic("padding bits cannot be larger than 7")     # comp.rs:59-61
        self._comp, self._pad, self._tree = cb, int(padding_bits), huff_tree

    def comp_bytes(self) -> np.ndarray:
        return self._comp

    def padding_bits(self) -> int:
        return self._pad

    def huff_tree(self) -> HuffTree:
        return self._tree

    def into_inner(self):
        return self._comp, self._pad, self._tree

    def to_bytes(self) -> bytes:
        """comp.rs:279-300"""
        out = np.empty(self._comp.size + 512, dtype=np.uint8)
        n = C.c_size_t(0)
        _raise(L.load().hb_to_bytes(self._comp.ctypes.data, self._comp.size, self._pad, C.byref(self._tree.raw),
                                    out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value].tobytes()

    @classmethod
    def try_from_bytes(cls, blob, copy: bool = True) -> "CompressData":
        """comp.rs:128-184.  copy=False keeps comp_bytes a view of `blob` (e.g. a pinned file image)."""
        a = _u8(blob)
        t = L.HbTree()
        off, ln, pad = C.c_size_t(0), C.c_size_t(0), C.c_uint8(0)
        st = L.load().hb_try_from_bytes(a.ctypes.data if a.size else None, a.size, C.byref(t),
                                        C.byref(off), C.byref(ln), C.byref(pad))
        if st == L.HB_ERR_BYTES_SHORT:
            msg = ("slice is empty" if a.size == 0 else
                   "slice too short to read tree length" if a.size < 5 else "slice too short to read tree")
            raise CompressedDataFromBytesError(msg)
        _raise(st)
        data = a[off.value: off.value + ln.value]
        return cls(data.copy() if copy else data, pad.value, HuffTree(t))


# ---------------------------------------------------------------- compress / decompress (host buffers)
def _take(ptr: C.c_void_p, n: int) -> np.ndarray:
    out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(max(n, 1),))[:n].copy()
    L.load().hb_free(ptr)
    return out


def compress(letters, ctx: Context | None = None, out: np.ndarray | None = None) -> CompressData:
    """comp.rs:353-356.  `out`: optional caller-owned u8 buffer (e.g. pinned memory reused across calls) that
    receives comp_bytes; the returned CompressData then views it."""
    a = _u8(letters)
    ctx = ctx or default_context()
    t = L.HbTree()
    n, pad = C.c_size_t(0), C.c_uint8(0)
    if out is not None:
        _raise(L.load().hb_compress_u8_into(ctx.handle, a.ctypes.data, a.size, L.HB_ORDER_ASC, C.byref(t),
                                            out.ctypes.data, out.size, C.byref(n), C.byref(pad)))
        return CompressData(out[: n.value], pad.value, HuffTree(t))
    ptr = C.c_void_p()
    _raise(L.load().hb_compress_u8(ctx.handle, a.ctypes.data, a.size, L.HB_ORDER_ASC, C.byref(t),
                                   C.byref(ptr), C.byref(n), C.byref(pad)))
    return CompressData(_take(ptr, n.value), pad.value, HuffTree(t))


class PinnedBuffer:
    """Page-locked host memory from the library (hb_host_alloc): file and network staging that moves over PCIe at full
    speed and asynchronously.  `.array` is a numpy u8 view; release with close() or use as a context manager."""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        _raise(L.load().hb_host_alloc(max(int(nbytes), 1), C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(int(nbytes), 1),))[: int(nbytes)]

    def close(self):
        if self._p:
            self.array = None
            L.load().hb_host_free(self._p)
            self._p = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def compress_with_tree(letters, huff_tree: HuffTree, ctx: Context | None = None, out: np.ndarray | None = None) -> CompressData:
    """comp.rs:419-451; raises CompressError(missing_letter) when the tree lacks a letter.  `out`: optional
    caller-owned u8 buffer (e.g. a PinnedBuffer) that receives comp_bytes."""
    a = _u8(letters)
    ctx = ctx or default_context()
    if out is not None:
        n, pad, missing = C.c_size_t(0), C.c_uint8(0), C.c_uint8(0)
        st = L.load().hb_compress_with_tree_u8_into(ctx.handle, a.ctypes.data, a.size, C.byref(huff_tree.raw),
                                                    out.ctypes.data, out.size, C.byref(n), C.byref(pad), C.byref(missing))
        _raise(st, missing.value)
        return CompressData(out[: n.value], pad.value, huff_tree)
    ptr, n, pad, missing = C.c_void_p(), C.c_size_t(0), C.c_uint8(0), C.c_uint8(0)
    st = L.load().hb_compress_with_tree_u8(ctx.handle, a.ctypes.data, a.size, C.byref(huff_tree.raw),
                                           C.byref(ptr), C.byref(n), C.byref(pad), C.byref(missing))
    _raise(st, missing.value)
    return CompressData(_take(ptr, n.value), pad.value, huff_tree)


def decompress(comp_data: CompressData, ctx: Context | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """comp.rs:487-519.  `out`: optional caller-owned u8 buffer that receives the letters."""
    ctx = ctx or default_context()
    cb = comp_data.comp_bytes()
    n = C.c_size_t(0)
    if out is not None:
        _raise(L.load().hb_decompress_u8_into(ctx.handle, cb.ctypes.data, cb.size, comp_data.padding_bits(),
                                              C.byref(comp_data.huff_tree().raw), out.ctypes.data, out.size, C.byref(n)))
        return out[: n.value]
    ptr = C.c_void_p()
    _raise(L.load().hb_decompress_u8(ctx.handle, cb.ctypes.data, cb.size, comp_data.padding_bits(),
                                     C.byref(comp_data.huff_tree().raw), C.byref(ptr), C.byref(n)))
    return _take(ptr, n.value)
