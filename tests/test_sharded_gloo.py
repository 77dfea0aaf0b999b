"""World-size-2 (and 3) `gloo` tests of the multi-GPU orchestration (huff_encoding_b200/sharded.py) on CPU.

The CUDA engine is replaced by an engine with the same methods backed by the ORACLE (test infrastructure), so what is
exercised here is the host logic: all-reduce of the histogram, the exclusive scan of shard bit totals, start-bit
encoding, the OR-merge of the byte two shards share, and the neighbour entry/exit verification of the byte-sharded
decode.  The concatenated stream must equal the oracle's single-stream output bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from huff_encoding_b200 import datagen as G
from huff_encoding_b200.api import HuffTree
from huff_encoding_b200.sharded import ShardedCodec
from oracle import oracle as O


class OracleEngine:
    """Same surface as huff_encoding_b200.engine.Engine, computed on the CPU by the oracle."""

    def __init__(self):
        self.device = torch.device("cpu")
        self.stream = None
        self._last = None

    def histogram(self, data):
        return torch.from_numpy(O.histogram(data.numpy()).astype(np.int64))

    def shard_plan(self, hists, order=0):
        import ctypes as C
        from huff_encoding_b200 import _lib as L
        h = np.ascontiguousarray(np.asarray(hists).astype(np.uint64))
        t = L.HbTree()
        bits = (C.c_uint64 * h.shape[0])()
        assert L.load().hb_shard_plan(h.ctypes.data_as(C.POINTER(C.c_uint64)), h.shape[0], order, C.byref(t), bits) == 0
        return HuffTree(t), list(bits)                                                # host code of the product library

    def tree_from_weights(self, w, order=0):
        return HuffTree.from_weights({b: int(w[b]) for b in range(256) if w[b]})     # host code of the product library

    def _otree(self, tree):
        return O.tree_from_pairs([l for l, _ in self._pairs(tree)], [w for _, w in self._pairs(tree)])

    @staticmethod
    def _pairs(tree):
        t = tree.raw
        return [(t.nodes[i].letter, t.nodes[i].weight) for i in range(t.n_leaves)]

    def encode(self, data, tree, out, start_bit=0, total_bits=None):
        comp, pad = O.compress_with_tree(data.numpy(), self._otree(tree))
        bits = np.unpackbits(comp)[: comp.size * 8 - pad]
        shifted = np.concatenate([np.zeros(start_bit, np.uint8), bits])
        packed = np.packbits(shifted)
        out[: packed.size] = torch.from_numpy(packed)

    def decode_count(self, buf, avail_bits, own_begin, own_end, stream_bit0, tree, entry_bit=-1):
        ot = self._otree(tree)
        bits = np.unpackbits(buf.numpy())[:avail_bits]
        lens_by_letter = ot.lens()

        def walk(pos, stop, collect):
            out = []
            while pos < stop:
                node = ot.root
                p = pos
                if ot.nodes[node].left == O.HO_NONE:
                    p += 1
                else:
                    while ot.nodes[node].left != O.HO_NONE:
                        if p >= avail_bits:
                            return None, out
                        node = ot.nodes[node].right if bits[p] else ot.nodes[node].left
                        p += 1
                if p > avail_bits:
                    return None, out
                if collect:
                    out.append(ot.nodes[node].letter)
                pos = p
            return pos, out

        if entry_bit < 0:                        # speculative: resynchronise over a 1024-bit look-back
            start = max(own_begin - 1024, 0)
            entry_bit, _ = walk(start, own_begin, False)
        exit_bit, letters = walk(entry_bit, own_end, True)
        self._last = np.array(letters, dtype=np.uint8)
        return entry_bit, (avail_bits if exit_bit is None else exit_bit), len(letters)

    def decode_write(self, out):
        out[: self._last.size] = torch.from_numpy(self._last)


def _engine():
    """The oracle-backed engine (host logic only), or -- under tests/emu (HB_EMU=1, see tests/test_emu_model.py) -- the engine
    over the CPU model of the library: the same orchestration with the real kernels' code."""
    if os.environ.get("HB_EMU") == "1":
        from tests.emu.model_engine import ModelEngine
        return ModelEngine()
    return OracleEngine()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = getattr(G, kind)(n_total)
        per = n_total // world
        lo, hi = rank * per, (n_total if rank == world - 1 else (rank + 1) * per)
        shard = torch.from_numpy(full[lo:hi].copy())
        codec = ShardedCodec(_engine(), world, rank, dist)
        comp_buf = torch.zeros(shard.numel() * 2 + 64, dtype=torch.uint8)
        info = codec.compress(shard, comp_buf)

        # 1. the concatenation of the shard streams is the single-stream output
        gathered = codec.gather_stream(comp_buf, info)
        ref_comp, ref_pad, ref_tree = O.compress(full)
        if rank == 0:
            stream, pad = gathered
            assert pad == ref_pad and np.array_equal(stream, ref_comp), "concatenated shards differ from one stream"
        assert info["tree"].read_codes() == ref_tree.codes()
        assert sum(info["all_bits"]) == ref_comp.size * 8 - ref_pad

        # 2. every rank decodes its own shard (known entry)
        out_buf = torch.zeros(shard.numel() + 64, dtype=torch.uint8)
        n = codec.decompress(comp_buf, info, out_buf)
        assert n == shard.numel() and torch.equal(out_buf[:n], shard)

        # 3. byte-sharded decode of the single stream: speculative entries verified against the neighbour's exit
        total_bits = ref_comp.size * 8 - ref_pad
        cut = [ref_comp.size * g // world for g in range(world + 1)]
        b0 = max(cut[rank] - 256, 0)
        b1 = min(cut[rank + 1] + 256, ref_comp.size)
        buf = torch.from_numpy(ref_comp[b0:b1].copy())
        out, cnt, letter_off = codec.decompress_byte_sharded(buf, b0, cut[rank], cut[rank + 1], total_bits,
                                                             info["tree"], lambda k: torch.zeros(k + 8, dtype=torch.uint8))
        got = out[:cnt].numpy()
        assert np.array_equal(got, full[letter_off: letter_off + cnt]), "byte-sharded decode mismatch"
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([cnt], dtype=torch.int64))
        assert sum(int(c) for c in counts) == n_total
        q.put((rank, "ok"))
    except Exception as e:                       # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


CASES = [(2, "english", 60_001), (2, "zipf", 50_003), (3, "uniform", 30_000), (2, "uniform", 4096)]
if os.environ.get("HB_EMU") == "1":          # real kernels (CPU model): sizes that span many chunks and sub-regions
    CASES += [(2, "zipf", 3_000_001), (4, "english", 1_234_567),
              (3, "english", 2), (4, "zipf", 5), (4, "english", 40)]        # empty and few-letter shards


@pytest.mark.parametrize("world,kind,n", CASES)
def test_sharded_codec_gloo(world, kind, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"
