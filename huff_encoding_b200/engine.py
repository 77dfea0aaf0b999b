"""Device-resident front end: the `*_dev` entry points of the C ABI driven with torch CUDA tensors.

torch is plumbing here (device memory, streams); all compute is libhuffb200's kernels.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .api import Context, HuffTree, _raise


class Engine:
    """One CUDA device, one hb_ctx.  Tensors passed in must live on that device and be uint8 + contiguous."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("huff_encoding_b200.Engine needs a CUDA device (there is no CPU fallback)")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.ctx = Context(self.device_index)
        self.lib = L.load()
        self._hist = torch.zeros(256, dtype=torch.int64, device=self.device)
        self._stream = torch.cuda.ExternalStream(self.ctx.stream(), device=self.device)
        self.comm_world = None

    # -- helpers
    @property
    def stream(self) -> torch.cuda.Stream:
        """The stream every kernel of this engine runs on (time with events recorded on THIS stream)."""
        return self._stream

    def _check(self, t: torch.Tensor):
        if t.dtype != torch.uint8 or not t.is_contiguous() or t.device != self.device:
            raise ValueError("expected a contiguous uint8 tensor on " + str(self.device))

    def sync(self):
        self.ctx.sync()

    @contextlib.contextmanager
    def _ordered(self):
        """Order the ctx stream after the caller's current torch stream and back (no-op when they are the same)."""
        cur = torch.cuda.current_stream(self.device)
        same = cur.cuda_stream == self._stream.cuda_stream
        if not same:
            self._stream.wait_stream(cur)
        try:
            yield
        finally:
            if not same:
                cur.wait_stream(self._stream)

    def kernel_launches(self) -> int:
        return self.ctx.kernel_launches()

    # -- K1
    def histogram(self, data: torch.Tensor) -> torch.Tensor:
        """256 x int64 counts on the device (no host sync)."""
        self._check(data)
        with self._ordered():
            _raise(self.lib.hb_histogram_u8_dev(self.ctx.handle, data.data_ptr(), data.numel(), self._hist.data_ptr()))
        return self._hist

    # -- K2
    def encode(self, data: torch.Tensor, tree: HuffTree, out: torch.Tensor, start_bit: int = 0,
               total_bits: torch.Tensor | None = None):
        """compress_with_tree's packing loop: writes the stream into `out` (no host sync)."""
        self._check(data)
        self._check(out)
        tb = total_bits.data_ptr() if total_bits is not None else None
        with self._ordered():
            _raise(self.lib.hb_encode_u8_dev(self.ctx.handle, data.data_ptr(), data.numel(), C.byref(tree.raw),
                                             start_bit, out.data_ptr(), out.numel(), tb))

    def compress(self, data: torch.Tensor, out: torch.Tensor | None = None, order: int = L.HB_ORDER_ASC):
        """compress(): histogram -> host tree -> encode.  Returns (out, comp_len, padding_bits, HuffTree)."""
        self._check(data)
        if out is None:
            out = torch.empty(data.numel() + data.numel() // 8 + 64, dtype=torch.uint8, device=self.device)
        self._check(out)
        t = L.HbTree()
        n, pad = C.c_size_t(0), C.c_uint8(0)
        with self._ordered():
            st = self.lib.hb_compress_u8_dev(self.ctx.handle, data.data_ptr(), data.numel(), order, C.byref(t),
                                             out.data_ptr(), out.numel(), C.byref(n), C.byref(pad))
            if st == L.HB_ERR_CAPACITY:
                out = torch.empty(((n.value + 3) // 4) * 4 + 64, dtype=torch.uint8, device=self.device)
                st = self.lib.hb_compress_u8_dev(self.ctx.handle, data.data_ptr(), data.numel(), order, C.byref(t),
                                                 out.data_ptr(), out.numel(), C.byref(n), C.byref(pad))
        _raise(st)
        return out, n.value, pad.value, HuffTree(t)

    # -- K3
    def decompress(self, comp: torch.Tensor, comp_len: int, padding_bits: int, tree: HuffTree,
                   out: torch.Tensor | None = None):
        """decompress(): returns (out, n_letters).  `comp` must be readable up to comp_len rounded up to 4 bytes."""
        self._check(comp)
        if comp.numel() < ((comp_len + 3) // 4) * 4:
            raise ValueError("comp tensor must be padded to a multiple of 4 bytes")
        n = C.c_size_t(0)
        cap = out.numel() if out is not None else 0
        ptr = out.data_ptr() if out is not None else None
        with self._ordered():
            st = self.lib.hb_decompress_u8_dev(self.ctx.handle, comp.data_ptr(), comp_len, padding_bits,
                                               C.byref(tree.raw), ptr, cap, C.byref(n))
            if st == L.HB_ERR_CAPACITY:
                out = torch.empty(n.value + 64, dtype=torch.uint8, device=self.device)
                st = self.lib.hb_decode_write_dev(self.ctx.handle, out.data_ptr(), out.numel())
        _raise(st)
        if out is None:
            out = torch.empty(0, dtype=torch.uint8, device=self.device)
        return out, n.value

    def decode_count(self, buf: torch.Tensor, avail_bits: int, own_begin: int, own_end: int, stream_bit0: int,
                     tree: HuffTree, entry_bit: int = -1):
        """Shard count pass: returns (entry_bit, exit_bit, n_letters) relative to `buf`."""
        self._check(buf)
        info = L.HbShardInfo(entry_bit, 0, 0)
        with self._ordered():
            _raise(self.lib.hb_decode_count_dev(self.ctx.handle, buf.data_ptr(), avail_bits, own_begin, own_end,
                                                stream_bit0, C.byref(tree.raw), C.byref(info)))
        return info.entry_bit, info.exit_bit, info.n_letters

    def decode_shard(self, buf: torch.Tensor, avail_bits: int, own_begin: int, own_end: int, stream_bit0: int,
                     tree: HuffTree, entry_bit: int, out: torch.Tensor):
        """Count + write of a shard with a KNOWN entry in one call (fused one-pass decoder when the tree allows it).
        Returns (entry_bit, exit_bit, n_letters)."""
        self._check(buf)
        self._check(out)
        info = L.HbShardInfo(entry_bit, 0, 0)
        with self._ordered():
            _raise(self.lib.hb_decode_shard_dev(self.ctx.handle, buf.data_ptr(), avail_bits, own_begin, own_end,
                                                stream_bit0, C.byref(tree.raw), C.byref(info), out.data_ptr(), out.numel()))
        return info.entry_bit, info.exit_bit, info.n_letters

    def decode_write(self, out: torch.Tensor):
        self._check(out)
        with self._ordered():
            _raise(self.lib.hb_decode_write_dev(self.ctx.handle, out.data_ptr(), out.numel()))

    # -- multi-GPU inside the library (hb_comm_* / hb_*_shard_dev): one rank per Engine
    def comm_init(self, world: int, rank: int, unique_id: bytes | None = None):
        """Bind this engine's ctx to rank `rank` of a `world`-rank communicator owned by the library (NCCL).  Rank 0 gets
        the id from `comm_unique_id()`; handing it to the other ranks is the host application's job."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _raise(self.lib.hb_comm_init(self.ctx.handle, world, rank, buf))
        self.comm_world, self.comm_rank = world, rank

    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        _raise(self.lib.hb_comm_get_unique_id(buf))
        return bytes(buf)

    def comm_finalize(self):
        _raise(self.lib.hb_comm_finalize(self.ctx.handle))
        self.comm_world = None

    def compress_shard(self, data: torch.Tensor, out: torch.Tensor, order: int = L.HB_ORDER_ASC):
        """Collective compress() of this rank's contiguous shard: returns (HbShardLayout, HuffTree)."""
        self._check(data)
        self._check(out)
        t, lay = L.HbTree(), L.HbShardLayout()
        with self._ordered():
            _raise(self.lib.hb_compress_shard_dev(self.ctx.handle, data.data_ptr(), data.numel(), order, C.byref(t),
                                                  out.data_ptr(), out.numel(), C.byref(lay)))
        return lay, HuffTree(t)

    def decompress_shard(self, comp: torch.Tensor, layout, tree: HuffTree, out: torch.Tensor) -> int:
        self._check(comp)
        self._check(out)
        n = C.c_size_t(0)
        with self._ordered():
            _raise(self.lib.hb_decompress_shard_dev(self.ctx.handle, comp.data_ptr(), C.byref(layout), C.byref(tree.raw),
                                                    out.data_ptr(), out.numel(), C.byref(n)))
        return n.value

    # -- host tree from a histogram
    def tree_from_histogram(self, hist: torch.Tensor, order: int = L.HB_ORDER_ASC) -> HuffTree:
        return self.tree_from_weights(hist.cpu().numpy(), order)

    def shard_plan(self, hists, order: int = L.HB_ORDER_ASC):
        """Multi-GPU plan (hb_shard_plan): (G, 256) gathered histograms -> (tree of their sum, [bits of every shard])."""
        h = np.ascontiguousarray(hists)
        h = h.view(np.uint64) if h.dtype == np.int64 else h.astype(np.uint64, copy=False)
        t = L.HbTree()
        bits = (C.c_uint64 * h.shape[0])()
        _raise(self.lib.hb_shard_plan(h.ctypes.data_as(C.POINTER(C.c_uint64)), h.shape[0], order, C.byref(t), bits))
        return HuffTree(t), list(bits)

    def tree_from_weights(self, weights, order: int = L.HB_ORDER_ASC) -> HuffTree:
        w = np.ascontiguousarray(np.asarray(weights).astype(np.uint64))
        t = L.HbTree()
        _raise(self.lib.hb_tree_from_weights(w.ctypes.data_as(C.POINTER(C.c_uint64)), order, C.byref(t)))
        return HuffTree(t)
