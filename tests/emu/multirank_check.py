"""TEST INFRASTRUCTURE.  The multi-GPU path INSIDE the library (hb_comm_init, hb_compress_shard_dev with its all-gather,
hb_decompress_shard_dev) with G ranks as threads of this process, on the CPU model (tests/emu; NCCL is nccl_emu.cpp):
the concatenated shard streams must be the oracle's stream of the concatenated input, bit for bit, and every shard must
decode to its letters.  usage: HB_EMU=1 HUFFB200_SO=.../libhuffb200_emu.so python tests/emu/multirank_check.py [G]"""
import os
import sys
import threading

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from huff_encoding_b200 import datagen as G  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.emu.model_engine import ModelEngine  # noqa: E402


def run(world: int, gen: str, sizes: list[int]):
    offs = np.concatenate([[0], np.cumsum(sizes)])
    whole = getattr(G, gen)(int(offs[-1]), seed=11)
    shards = [whole[offs[r]: offs[r + 1]] for r in range(world)]
    uid_box, results, errors = {}, [None] * world, []
    ready = threading.Barrier(world)

    def rank_main(r):
        try:
            eng = ModelEngine()
            if r == 0:
                uid_box["id"] = eng.comm_unique_id()
            ready.wait()
            eng.comm_init(world, r, uid_box["id"])
            d = torch.from_numpy(shards[r].copy())
            n = d.numel()
            comp = torch.full((n + n // 4 + 4096,), 0x5A, dtype=torch.uint8)
            lay, tree = eng.compress_shard(d, comp)
            out = torch.empty(n + 64, dtype=torch.uint8)
            m = eng.decompress_shard(comp, lay, tree, out)
            assert m == n and torch.equal(out[:n], d), f"rank {r}: shard round trip"
            results[r] = (lay.bit_offset, lay.bits, lay.total_bits, lay.start_bit, lay.padding_bits, lay.comp_len,
                          comp[: lay.comp_len].numpy().copy(), tree.read_codes())
            eng.comm_finalize()
        except BaseException as e:          # noqa: BLE001
            errors.append((r, repr(e)))
            try:
                ready.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    comp, pad, otree = O.compress(whole)
    got = np.zeros(comp.size, dtype=np.uint8)
    at = 0
    for r in range(world):
        off, bits, total, sb, p, clen, piece, codes = results[r]
        assert off == at and sb == off % 8 and p == pad and total == comp.size * 8 - pad, (r, off, at)
        assert codes == otree.codes(), f"rank {r}: tree differs from the oracle's"
        got[off // 8: off // 8 + clen] |= piece
        at += bits
    assert at == comp.size * 8 - pad and np.array_equal(got, comp), f"{gen}: concatenated shards differ from the oracle's stream"
    print(f"ok: {world} ranks, {gen}, {int(offs[-1])} letters -> {comp.size} bytes")


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    base = 700_001
    run(world, "english", [base + 13 * r for r in range(world)])
    run(world, "zipf", [base // 2 + 4097 * r for r in range(world)])
    run(2, "uniform", [300_000, 300_000])                 # fixed-length fast path, byte-aligned shards
    # empty and tiny shards: an empty shard owns no byte of the stream, not even the one its neighbours share
    run(8, "english", [0, 1, 17, 0, 100_001, 3, 65_536, 1])
    run(8, "zipf", [5, 0, 0, 0, 0, 0, 0, 300_001])
    run(3, "zipf", [1, 1, 1])
