"""The warp-per-sub-region encoder (hb_encode_warps.cuh) against the oracle and against the region encoder.  Bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from huff_encoding_b200 import datagen as G
from oracle import oracle as O
from tests._model import dev, dev_sync, make_engine


@pytest.fixture(scope="module")
def hb():
    import huff_encoding_b200 as m
    from huff_encoding_b200 import build
    build.build()
    return m


def _engine(**env):
    from huff_encoding_b200.engine import Engine
    for k, v in env.items():
        os.environ[k] = v
    try:
        return make_engine()
    finally:
        for k in env:
            del os.environ[k]


# sizes around the tile (1024 letters), the smallest sub-region and the point where sub-regions grow past one tile
SIZES = [1, 31, 32, 33, 1023, 1024, 1025, 2047, 4096, 148 * 32 * 1024 - 1, 148 * 32 * 1024, 148 * 32 * 1024 + 1,
         148 * 32 * 1024 + 1025, 9_999_999]


@pytest.mark.parametrize("n", SIZES)
def test_warp_encoder_matches_oracle(hb, n):
    data = G.english(n, seed=n)
    cd = hb.compress(data)
    comp, pad, tree = O.compress(data)
    assert cd.padding_bits() == pad
    assert np.array_equal(cd.comp_bytes(), comp)


def test_warp_encoder_equals_region_encoder_with_start_bits():
    import torch
    new, old = _engine(), _engine(HB_NO_ENCODE_WARPS="1")
    for gen, n in (("zipf", 5_000_011), ("english", 3_000_000), ("uniform", 2_000_001)):
        data = getattr(G, gen)(n)
        d = torch.from_numpy(data).to(dev())
        tree = new.tree_from_weights(np.bincount(data, minlength=256))
        for start_bit in (0, 1, 7, 13, 31):
            outs = []
            for eng in (new, old):
                out = torch.zeros(n + n // 2 + 64, dtype=torch.uint8, device=dev())
                tb = torch.zeros(1, dtype=torch.int64, device=dev())
                eng.histogram(d)
                eng.encode(d, tree, out, start_bit=start_bit, total_bits=tb)
                dev_sync()
                bits = int(tb.item())
                outs.append((bits, out[: (start_bit + bits + 7) // 8].cpu().numpy()))
            assert outs[0][0] == outs[1][0], (gen, start_bit)
            assert np.array_equal(outs[0][1], outs[1][1]), (gen, start_bit)


def test_same_address_different_data_is_not_served_from_a_stale_histogram():
    # the per-region counts of one buffer must never be applied to another buffer at the same address
    import torch
    eng = _engine()
    n = 3_000_000
    a = torch.from_numpy(G.english(n, seed=1)).to(dev())
    tree = eng.tree_from_weights(np.bincount(np.concatenate([G.english(n, seed=1), G.english(n, seed=2)]), minlength=256))
    out = torch.zeros(n + 64, dtype=torch.uint8, device=dev())
    eng.histogram(a)
    eng.encode(a, tree, out)
    a.copy_(torch.from_numpy(G.english(n, seed=2)).to(dev()))     # same address, other letters, no new histogram call
    out2 = torch.zeros(n + 64, dtype=torch.uint8, device=dev())
    tb = torch.zeros(1, dtype=torch.int64, device=dev())
    eng.encode(a, tree, out2, total_bits=tb)
    dev_sync()
    bits = int(tb.item())
    otree = O.tree_from_weights(np.bincount(np.concatenate([G.english(n, seed=1), G.english(n, seed=2)]), minlength=256).astype(np.uint64))
    comp, pad = O.compress_with_tree(G.english(n, seed=2), otree)
    assert bits == comp.size * 8 - pad
    assert np.array_equal(out2[: comp.size].cpu().numpy(), comp)


def test_encode_dev_reports_letters_without_a_code():
    import torch
    eng = _engine()
    n = 100_000
    data = G.english(n)
    tree = eng.tree_from_weights(np.bincount(data, minlength=256))
    bad = data.copy()
    bad[77_777] = 0x01                                        # a letter the tree has no code for
    d = torch.from_numpy(bad).to(dev())
    out = torch.zeros(n + 64, dtype=torch.uint8, device=dev())
    eng.encode(d, tree, out)
    flag = C.c_uint32(0)
    assert eng.lib.hb_ctx_last_encode_error(eng.ctx.handle, C.byref(flag)) == 0 and flag.value == 1
    d2 = torch.from_numpy(data).to(dev())
    eng.encode(d2, tree, out)
    assert eng.lib.hb_ctx_last_encode_error(eng.ctx.handle, C.byref(flag)) == 0 and flag.value == 0
