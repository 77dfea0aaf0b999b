"""Two ranks, two GPUs, the communicator inside the library (hb_comm_*, hb_compress_shard_dev, hb_decompress_shard_dev):
the gathered shard streams must equal the single-GPU stream and the oracle's.  Skipped on boxes with one GPU."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rank(rank, world, tmp, n):
    import time

    import torch
    torch.cuda.set_device(rank)
    from huff_encoding_b200 import datagen as G
    from huff_encoding_b200.engine import Engine
    eng = Engine(rank)
    idf = os.path.join(tmp, "id.bin")
    if rank == 0:
        uid = eng.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        while not os.path.exists(idf):
            time.sleep(0.01)
        uid = open(idf, "rb").read()
    eng.comm_init(world, rank, uid)
    for trial, gen in enumerate(("english", "zipf")):
        data = getattr(G, gen)(n, offset=rank * n)
        d = torch.from_numpy(data).cuda()
        comp = torch.zeros(n + n // 4 + 64, dtype=torch.uint8, device="cuda")
        lay, tree = eng.compress_shard(d, comp)
        out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        m = eng.decompress_shard(comp, lay, tree, out)
        assert m == n and torch.equal(out[:n], d), "shard round trip"
        np.save(os.path.join(tmp, f"s{trial}_{rank}.npy"), comp[: lay.comp_len].cpu().numpy())
        with open(os.path.join(tmp, f"l{trial}_{rank}.txt"), "w") as f:
            f.write(f"{lay.bit_offset} {lay.bits} {lay.total_bits} {lay.start_bit} {lay.padding_bits} {lay.comp_len}")
    eng.comm_finalize()


def test_two_rank_streams_concatenate_to_the_single_gpu_stream():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from huff_encoding_b200 import build, datagen as G
    from oracle import oracle as O
    build.build()
    n, world = 6_000_003, 2
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_rank, args=(world, tmp, n), nprocs=world, join=True)
        for trial, gen in enumerate(("english", "zipf")):
            whole = np.concatenate([getattr(G, gen)(n, offset=r * n) for r in range(world)])
            comp, pad, _ = O.compress(whole)
            got = np.zeros(comp.size, dtype=np.uint8)
            for r in range(world):
                off, bits, total, sb, p, clen = (int(x) for x in open(os.path.join(tmp, f"l{trial}_{r}.txt")).read().split())
                piece = np.load(os.path.join(tmp, f"s{trial}_{r}.npy"))
                assert p == pad and total == comp.size * 8 - pad and piece.size == clen
                got[off // 8: off // 8 + clen] |= piece
            assert np.array_equal(got, comp), gen


def test_single_rank_comm_matches_compress(tmp_path):
    import torch
    from huff_encoding_b200 import build, datagen as G
    from tests._model import dev, make_engine
    from oracle import oracle as O
    build.build()
    eng = make_engine()
    eng.comm_init(1, 0, None)
    data = G.zipf(3_000_001)
    d = torch.from_numpy(data).to(dev())
    comp = torch.zeros(data.size + 64, dtype=torch.uint8, device=dev())
    lay, tree = eng.compress_shard(d, comp)
    ref, pad, _ = O.compress(data)
    assert lay.start_bit == 0 and lay.bit_offset == 0 and lay.padding_bits == pad and lay.comp_len == ref.size
    assert np.array_equal(comp[: ref.size].cpu().numpy(), ref)
    out = torch.empty(data.size + 64, dtype=torch.uint8, device=dev())
    assert eng.decompress_shard(comp, lay, tree, out) == data.size and torch.equal(out[: data.size], d)
    eng.comm_finalize()
