"""The fused one-pass decoder (hb_decode_fused.cuh) against the oracle, and its fallbacks.  Bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from huff_encoding_b200 import datagen as G
from oracle import oracle as O


@pytest.fixture(scope="module")
def hb():
    import huff_encoding_b200 as m
    from huff_encoding_b200 import build
    build.build()
    return m


def _ctx_with(hb, **env):
    for k, v in env.items():
        os.environ[k] = v
    try:
        return hb.Context(0)
    finally:
        for k in env:
            del os.environ[k]


def _decode(hb, ctx, data):
    comp, pad, tree = O.compress(data)
    ours = hb.HuffTree.from_weights(hb.build_weights_map(data, ctx=ctx))
    assert ours.read_codes() == tree.codes()
    got = hb.decompress(hb.CompressData(comp, pad, ours), ctx=ctx)
    assert got.size == data.size and np.array_equal(got, data), (got.size, data.size, int(np.argmax(got[:min(got.size, data.size)] != data[:min(got.size, data.size)])))
    return ctx.last_decode_path()


@pytest.mark.parametrize("gen,n", [("zipf", 1 << 20), ("zipf", 3_000_017), ("zipf", 4_321_987), ("zipf", (1 << 24) + 5),
                                   ("zipf", 33 * 1024 * 4 + 1), ("zipf", 1000), ("zipf", 31), ("zipf", 70_000)])
def test_fused_path_is_taken_and_exact(hb, gen, n):
    ctx = hb.Context(0)
    path, slow = _decode(hb, ctx, getattr(G, gen)(n))
    if n >= 70_000:                    # (the tree of a few dozen letters may be near fixed-length: two-pass decoder)
        assert path == 1 and slow == 0, (path, slow)
    ctx.close()


@pytest.mark.parametrize("n", [1 << 20, 4_321_987, 33 * 1024 * 4 + 1, 31])
def test_fused_kernel_on_english_with_forced_slots(hb, n):
    # English-like text packs ~250 letters into a subsequence: only one team's slots fit an SM, so the library picks the
    # two-pass decoder for it; forcing the slot size runs the fused kernel on it all the same (one team per SM)
    ctx = _ctx_with(hb, HB_FUSED_SLOT_WORDS="85")
    path, slow = _decode(hb, ctx, G.english(n))
    assert path == 1 and slow == 0, (path, slow)
    ctx.close()
    ctx = hb.Context(0)
    assert _decode(hb, ctx, G.english(n))[0] == 0
    ctx.close()


def test_fused_matches_two_pass(hb):
    a, b = hb.Context(0), _ctx_with(hb, HB_NO_FUSED="1")
    for gen, n in (("zipf", 2_000_003), ("zipf", 1_234_567)):
        data = getattr(G, gen)(n, seed=n)
        assert _decode(hb, a, data)[0] == 1
        assert _decode(hb, b, data)[0] == 0
    a.close()
    b.close()


def test_fused_slot_overflow_takes_the_slow_chunk_path(hb):
    # slots far too small for the letters of a subsequence: every chunk is written letter by letter, still exact
    ctx = _ctx_with(hb, HB_FUSED_SLOT_WORDS="21")
    for gen, n in (("zipf", 1_500_001), ("english", 700_003)):
        path, slow = _decode(hb, ctx, getattr(G, gen)(n))
        assert path == 1 and slow > 0, (path, slow)
    ctx.close()


def test_fused_overflow_in_a_dense_region_only(hb):
    # zipf letters with a long run of the most frequent letter (2-bit code) in the middle: the chunks inside the run hold
    # ~2.6x the letters the slots were sized for
    data = G.zipf(6_000_011).copy()
    data[2_000_000:2_250_000] = 0                            # (short enough not to change the tree much)
    ctx = hb.Context(0)
    path, slow = _decode(hb, ctx, data)
    assert path == 1 and slow > 0, (path, slow)
    ctx.close()


def test_fused_refuted_speculation_falls_back(hb):
    ctx = _ctx_with(hb, HB_DEBUG_SPOIL_SPECULATION="1")
    path, _ = _decode(hb, ctx, G.zipf(3_000_017))
    assert path == 2
    ctx.close()


def test_fused_skewed_and_near_uniform_trees(hb):
    ctx = hb.Context(0)
    rng = np.random.default_rng(5)
    # two letters + rare third (min_len 1), and a 200-letter near-uniform alphabet (8-bit-ish codes, not a perfect tree)
    skew = rng.choice(np.array([65, 66, 67], np.uint8), size=2_000_003, p=[0.9, 0.09, 0.01])
    flat = rng.integers(0, 200, size=1_500_007).astype(np.uint8)
    # the skewed tree packs ~950 letters into a subsequence: no slot that large fits, the two-pass decoder serves it
    assert _decode(hb, ctx, skew)[0] in (0, 1)
    # codes of 7 and 8 bits only: such sets resynchronise slowly, the library routes them to the two-pass decoder
    assert _decode(hb, ctx, flat)[0] == 0
    ctx.close()


def test_fused_device_buffers_unaligned_output_and_capacity(hb):
    import torch
    from huff_encoding_b200.engine import Engine
    eng = Engine(0)
    data = G.zipf(2_345_679)
    d = torch.from_numpy(data).cuda()
    comp, n, pad, tree = eng.compress(d)
    base = torch.empty(data.size + 256, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 7, 32, 45):
        out = base[off: off + data.size + 64]
        dec, m = eng.decompress(comp, n, pad, tree, out=out)
        assert m == data.size and torch.equal(dec[:m], d), off
        assert eng.ctx.last_decode_path()[0] == 1
    # a buffer that is too small: the needed size is reported and a second call with a larger buffer works
    small = torch.empty(1000, dtype=torch.uint8, device="cuda")
    dec, m = eng.decompress(comp, n, pad, tree, out=small)
    assert m == data.size and torch.equal(dec[:m], d)
