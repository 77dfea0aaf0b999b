"""Multi-GPU path: contiguous input shards, one process per GPU, torch.distributed for the two tiny exchanges.

compress:
    local histogram kernel -> ONE all_gather of the G local 256-bin histograms (G x 2 KiB): their sum is the
    all-reduced global histogram, from which every rank builds the identical tree on the host, and
    sum(hist_g * len) for every g gives all shard bit totals at once -> exclusive scan = the shard's GLOBAL bit
    offset -> encode kernel with start_bit = offset % 8 into a buffer whose byte 0 is global byte offset // 8.  Concatenating the shard buffers (OR-ing the one byte two neighbours may share) gives exactly
    the single-GPU stream; `gather_stream` does that and the tests check it against the oracle.
decompress:
    (a) of shards produced by `compress`: every rank knows its first code-word start exactly (start_bit), no exchange.
    (b) of a foreign stream cut at byte boundaries (`decompress_byte_sharded`): each rank runs the count pass with a
        speculative entry found by self-synchronisation in its halo, ranks all_gather (entry, exit, count), any rank
        whose entry is refuted by its left neighbour's exit re-runs with the known entry (chain converges left to
        right), then the write pass.

The engine argument is the CUDA Engine in production.  The gloo/CPU tests inject an engine with the same
methods backed by the oracle to exercise this orchestration without a GPU; the product never does.
"""
from __future__ import annotations

import numpy as np
import torch


class ShardedCodec:
    def __init__(self, engine, world: int = 1, rank: int = 0, dist=None):
        self.eng, self.world, self.rank, self.dist = engine, world, rank, dist
        self.last_info = None
        self._gathered = None

    # ------------------------------------------------------------ helpers
    def _event(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(self.eng.stream)
        return ev

    # ------------------------------------------------------------ compress
    def compress(self, data: torch.Tensor, comp_buf: torch.Tensor, marks: dict | None = None) -> dict:
        eng, dist = self.eng, self.dist
        if marks is not None:
            a = self._event()
        local = eng.histogram(data)
        if marks is not None:
            marks["hist"] = (a, self._event())
        if self.world > 1:
            # ONE exchange: all-gather the G local histograms (G x 2 KiB).  Every rank sums them into the global
            # histogram (= the all-reduce) and, once the tree is known, also knows every shard's bit total
            # (= the all-gather of totals + exclusive scan) without a second collective.
            if self._gathered is None:
                self._gathered = torch.zeros(self.world, 256, dtype=local.dtype, device=local.device)
            try:
                dist.all_gather_into_tensor(self._gathered, local)
            except Exception:                                           # backends without the flat variant (gloo)
                parts = [torch.zeros_like(local) for _ in range(self.world)]
                dist.all_gather(parts, local)
                self._gathered.copy_(torch.stack(parts))
            h = self._gathered.cpu().numpy()                            # host sync: the tree needs the histogram
        else:
            h = local.cpu().numpy()[None, :]
        # one host call (hb_shard_plan): tree of the summed histogram + every shard's bit total
        tree, all_bits = eng.shard_plan(h)
        my_bits = all_bits[self.rank]
        offset = sum(all_bits[: self.rank])
        total = sum(all_bits)
        start_bit = offset % 8
        comp_len = (start_bit + my_bits + 7) // 8 if my_bits else 0       # an empty shard owns no byte of the stream
        if comp_buf.numel() < ((comp_len + 3) // 4) * 4:
            raise ValueError("comp_buf too small for this shard")
        if marks is not None:
            a = self._event()
        eng.encode(data, tree, comp_buf, start_bit=start_bit)
        if marks is not None:
            marks["encode"] = (a, self._event())
        raw = tree.raw
        fixed = raw.max_len if (raw.min_len == raw.max_len and raw.max_len in (1, 2, 4, 8)) else 0
        info = {"fixed_len": fixed, "tree": tree, "bits": my_bits, "bit_offset": offset, "start_bit": start_bit, "comp_len": comp_len,
                "total_bits": total, "padding_bits": (8 - total % 8) % 8, "all_bits": all_bits}
        self.last_info = info
        return info

    # ------------------------------------------------------------ decompress of our own shards
    def decompress(self, comp_buf: torch.Tensor, info: dict, out_buf: torch.Tensor, marks: dict | None = None) -> int:
        eng = self.eng
        begin, end = info["start_bit"], info["start_bit"] + info["bits"]
        if marks is not None:
            a = self._event()
        if hasattr(eng, "decode_shard"):
            # one call: the fused one-pass decoder when the tree allows it (the entry of an own shard is known exactly)
            _, _, n = eng.decode_shard(comp_buf, end, begin, end, info["bit_offset"] - begin, info["tree"], begin, out_buf)
        else:
            _, _, n = eng.decode_count(comp_buf, end, begin, end, info["bit_offset"] - begin, info["tree"], entry_bit=begin)
            if out_buf.numel() < n:
                raise ValueError("out_buf too small")
            eng.decode_write(out_buf)
        if marks is not None:
            marks["decode"] = (a, self._event())
        info["n_letters"] = n
        return n

    def init_library_comm(self):
        """Move the exchange INTO the library: the ctx gets an NCCL communicator (id broadcast over torch.distributed once)
        and compress()/decompress() of a shard become single C calls (hb_compress_shard_dev / hb_decompress_shard_dev)."""
        eng = self.eng
        uid = None
        if self.world > 1:
            t = torch.zeros(128, dtype=torch.uint8, device=eng.device)
            if self.rank == 0:
                t.copy_(torch.frombuffer(bytearray(eng.comm_unique_id()), dtype=torch.uint8))
            self.dist.broadcast(t, 0)
            uid = bytes(t.cpu().numpy().tobytes())
        eng.comm_init(self.world, self.rank, uid)

    def _library_round_trip(self, data, comp_buf, out_buf):
        lay, tree = self.eng.compress_shard(data, comp_buf)
        n = self.eng.decompress_shard(comp_buf, lay, tree, out_buf)
        raw = tree.raw
        self.last_info = {"fixed_len": raw.max_len if (raw.min_len == raw.max_len and raw.max_len in (1, 2, 4, 8)) else 0,
                          "tree": tree, "bits": lay.bits, "bit_offset": lay.bit_offset, "start_bit": lay.start_bit,
                          "comp_len": lay.comp_len, "total_bits": lay.total_bits, "padding_bits": lay.padding_bits,
                          "n_letters": n, "all_bits": None}

    def round_trip(self, data, comp_buf, out_buf, want_events: bool = False):
        if not want_events and getattr(self.eng, "comm_world", None) == self.world:
            # the whole shard round trip inside the library, collectives included: no Python between the kernels
            self._library_round_trip(data, comp_buf, out_buf)
            return None
        if self.world == 1 and not want_events:
            # single GPU: the whole of compress() / decompress() runs inside the library (histogram -> host tree ->
            # encode, count -> write), no Python between the kernels
            _, clen, pad, tree = self.eng.compress(data, out=comp_buf)
            _, n = self.eng.decompress(comp_buf, clen, pad, tree, out=out_buf)
            raw = tree.raw
            self.last_info = {"fixed_len": raw.max_len if (raw.min_len == raw.max_len and raw.max_len in (1, 2, 4, 8)) else 0,
                              "tree": tree, "bits": clen * 8 - pad, "bit_offset": 0, "start_bit": 0, "comp_len": clen,
                              "total_bits": clen * 8 - pad, "padding_bits": pad, "all_bits": [clen * 8 - pad],
                              "n_letters": n}
            return None
        marks = {} if want_events else None
        info = self.compress(data, comp_buf, marks)
        self.decompress(comp_buf, info, out_buf, marks)
        return marks

    # ------------------------------------------------------------ concatenation (tests, and users who want one blob)
    def gather_stream(self, comp_buf: torch.Tensor, info: dict):
        """Rank 0 returns the whole stream (numpy u8, exactly ceil(total_bits/8) bytes) + padding_bits; others None."""
        total_bytes = (info["total_bits"] + 7) // 8
        mine = comp_buf[: info["comp_len"]]
        if self.world == 1:
            return mine.cpu().numpy().copy(), info["padding_bits"]
        if info.get("all_bits") is None:                     # shards compressed inside the library: exchange the sizes
            b = torch.tensor([info["bits"]], dtype=torch.int64, device=comp_buf.device)
            parts = [torch.zeros_like(b) for _ in range(self.world)]
            self.dist.all_gather(parts, b)
            info["all_bits"] = [int(x.item()) for x in parts]
        cap = max((info["all_bits"][g] + 7) // 8 + 2 for g in range(self.world))
        send = torch.zeros(cap, dtype=torch.uint8, device=comp_buf.device)
        send[: info["comp_len"]] = mine
        recv = [torch.zeros(cap, dtype=torch.uint8, device=comp_buf.device) for _ in range(self.world)] \
            if self.rank == 0 else None
        self.dist.gather(send, recv, dst=0)
        if self.rank != 0:
            return None
        out = np.zeros(total_bytes, dtype=np.uint8)
        off = 0
        for g in range(self.world):
            sb = off % 8
            ln = (sb + info["all_bits"][g] + 7) // 8 if info["all_bits"][g] else 0
            piece = recv[g][:ln].cpu().numpy()
            out[off // 8: off // 8 + ln] |= piece             # the shared boundary byte is OR-merged
            off += info["all_bits"][g]
        return out, info["padding_bits"]

    # ------------------------------------------------------------ decompress of a foreign stream cut at byte boundaries
    def decompress_byte_sharded(self, buf: torch.Tensor, buf_byte0: int, own_byte_begin: int, own_byte_end: int,
                                total_bits: int, tree, out_alloc):
        """`buf` holds stream bytes [buf_byte0, ...) including a halo of >= 128 bytes on both sides of the owned
        byte range [own_byte_begin, own_byte_end) (clipped at the stream ends).  Returns (out tensor, n_letters,
        global letter offset of this shard)."""
        eng, dist = self.eng, self.dist
        bit0 = buf_byte0 * 8
        avail = min(buf.numel() * 8, total_bits - bit0)
        own_b = own_byte_begin * 8 - bit0
        own_e = min(own_byte_end * 8, total_bits) - bit0
        entry = 0 if own_byte_begin == 0 else -1
        e, x, n = eng.decode_count(buf, avail, own_b, own_e, bit0, tree, entry_bit=entry)
        if self.world > 1:
            for _ in range(self.world):                     # a refuted entry can cascade at most world-1 times
                trip = torch.tensor([e + bit0, x + bit0, n], dtype=torch.int64, device=buf.device)
                parts = [torch.zeros_like(trip) for _ in range(self.world)]
                dist.all_gather(parts, trip)
                t = torch.stack(parts).cpu()
                bad = [g for g in range(1, self.world) if int(t[g, 0]) != int(t[g - 1, 1])]
                if not bad:
                    break
                if self.rank in bad:
                    want = int(t[self.rank - 1, 1]) - bit0
                    e, x, n = eng.decode_count(buf, avail, own_b, own_e, bit0, tree, entry_bit=want)
            counts = [int(v) for v in t[:, 2]]
        else:
            counts = [n]
        out = out_alloc(n)
        if n:
            eng.decode_write(out)
        return out, n, sum(counts[: self.rank])
