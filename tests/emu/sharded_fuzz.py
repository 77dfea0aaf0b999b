"""TEST INFRASTRUCTURE.  Seeded fuzz of huff_encoding_b200.sharded.ShardedCodec (the torch.distributed orchestration: compress
of contiguous shards, gather_stream, shard decode, byte-sharded decode of a foreign stream with speculative entries and
the neighbour check) over the CPU model of the library, ranks as threads with a tiny in-process stand-in for the three
collectives it uses.  World sizes 1..8, shard sizes from 0 letters up, streams down to a few bytes.
usage: HB_EMU=1 HUFFB200_SO=.../libhuffb200_emu.so python tests/emu/sharded_fuzz.py [cases] [seed]"""
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from huff_encoding_b200 import datagen as G  # noqa: E402
from huff_encoding_b200.sharded import ShardedCodec  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.emu.model_engine import ModelEngine  # noqa: E402


class ThreadDist:
    """all_gather / gather / all_gather_into_tensor for ranks that are threads (one instance per rank, shared state)."""

    class Shared:
        def __init__(self, world):
            self.world, self.barrier, self.slots = world, threading.Barrier(world), [None] * world

    def __init__(self, shared, rank):
        self.s, self.rank = shared, rank

    def _exchange(self, t):
        self.s.slots[self.rank] = t.clone()
        self.s.barrier.wait()
        got = [x.clone() for x in self.s.slots]
        self.s.barrier.wait()
        return got

    def all_gather(self, parts, t):
        for p, g in zip(parts, self._exchange(t)):
            p.copy_(g)

    def all_gather_into_tensor(self, out, t):
        out.copy_(torch.stack(self._exchange(t)).reshape(out.shape))

    def gather(self, t, parts, dst=0):
        got = self._exchange(t)
        if self.rank == dst:
            for p, g in zip(parts, got):
                p.copy_(g)


def run_case(world, gen, sizes, engines):
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    full = getattr(G, gen)(int(offs[-1]), seed=int(offs[-1]) + world)
    ref_comp, ref_pad, ref_tree = O.compress(full)
    total_bits = ref_comp.size * 8 - ref_pad
    shared = ThreadDist.Shared(world)
    errors = []

    def rank_main(rank):
        try:
            dist = ThreadDist(shared, rank)
            codec = ShardedCodec(engines[rank], world, rank, dist)
            shard = torch.from_numpy(full[offs[rank]: offs[rank + 1]].copy())
            comp_buf = torch.full((shard.numel() * 2 + 64,), 0x5A, dtype=torch.uint8)
            info = codec.compress(shard, comp_buf)
            gathered = codec.gather_stream(comp_buf, info)
            if rank == 0:
                stream, pad = gathered
                assert pad == ref_pad and np.array_equal(stream, ref_comp), "concatenated shards differ from one stream"
            assert info["tree"].read_codes() == ref_tree.codes()
            out_buf = torch.zeros(shard.numel() + 64, dtype=torch.uint8)
            n = codec.decompress(comp_buf, info, out_buf)
            assert n == shard.numel() and torch.equal(out_buf[:n], shard), "shard decode"
            # the single stream cut at byte boundaries, 256-byte halos
            cut = [ref_comp.size * g // world for g in range(world + 1)]
            b0 = max(cut[rank] - 256, 0)
            b1 = min(cut[rank + 1] + 256, ref_comp.size)
            buf = torch.from_numpy(np.concatenate([ref_comp[b0:b1], np.zeros(16, np.uint8)]))[: b1 - b0]
            out, cnt, letter_off = codec.decompress_byte_sharded(buf, b0, cut[rank], cut[rank + 1], total_bits, info["tree"],
                                                                 lambda k: torch.zeros(k + 8, dtype=torch.uint8))
            assert np.array_equal(out[:cnt].numpy(), full[letter_off: letter_off + cnt]), "byte-sharded decode mismatch"
            counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(counts, torch.tensor([cnt], dtype=torch.int64))
            assert sum(int(c) for c in counts) == full.size, "byte-sharded decode lost or invented letters"
        except BaseException as e:                          # noqa: BLE001
            errors.append((rank, repr(e)))
            shared.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    first = [e for e in errors if "BrokenBarrier" not in e[1]] or errors
    assert not errors, (world, gen, sizes, first[:2])


if __name__ == "__main__":
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 31)
    engines = [ModelEngine() for _ in range(8)]
    pool = [0, 1, 2, 5, 16, 33, 1023, 1024, 4097, 33791, 33792, 33793, 65536]
    for it in range(n_cases):
        world = int(rng.integers(1, 9))
        sizes = [int(rng.choice(pool)) if rng.random() < 0.6 else int(rng.integers(1, 200_000)) for _ in range(world)]
        if sum(sizes) == 0:
            sizes[-1] = 1 + int(rng.integers(0, 40))
        gen = str(rng.choice(["english", "zipf", "uniform"]))
        run_case(world, gen, sizes, engines)
    print(f"sharded fuzz: ok ({n_cases} cases)")
