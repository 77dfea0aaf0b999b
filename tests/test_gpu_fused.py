"""The fused single-kernel decoder (hb_decode_fused.cuh) against the oracle, and its fallbacks.  Bit-exact."""
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
    m = min(got.size, data.size)
    assert got.size == data.size and np.array_equal(got, data), (got.size, data.size, int(np.argmax(got[:m] != data[:m])))
    return ctx.last_decode_path()


# around one chunk (256 subsequences of 33 words = 33 792 stream bytes), several chunks, many teams' worth
SIZES = [31, 1000, 70_000, 1 << 20, 3_000_017, 4_321_987, (1 << 24) + 5]


@pytest.mark.parametrize("gen", ["zipf", "english"])
@pytest.mark.parametrize("n", SIZES)
def test_fused_path_is_taken_and_exact(hb, gen, n):
    ctx = hb.Context(0)
    path, _ = _decode(hb, ctx, getattr(G, gen)(n, seed=n))
    if n >= 70_000:                    # (the tree of a few dozen letters may be near fixed-length: two-pass decoder)
        assert path == 1, path
    ctx.close()


def test_fused_matches_two_pass(hb):
    a, b = hb.Context(0), _ctx_with(hb, HB_NO_FUSED="1")
    for gen, n in (("zipf", 2_000_003), ("english", 1_234_567)):
        data = getattr(G, gen)(n, seed=n)
        assert _decode(hb, a, data)[0] == 1
        assert _decode(hb, b, data)[0] == 0
    a.close()
    b.close()


@pytest.mark.parametrize("teams", ["1", "2", "3"])
def test_fused_with_fewer_teams_per_cta(hb, teams):
    ctx = _ctx_with(hb, HB_FUSED_TEAMS=teams)
    assert _decode(hb, ctx, G.zipf(5_000_011))[0] == 1
    ctx.close()


def test_fused_dense_and_sparse_regions(hb):
    # a long run of the most frequent letter (short code: ~2.6x the usual letters per subsequence) and a run of a rare
    # letter (12-bit code: few letters per subsequence, threads that own no row at all)
    data = G.zipf(6_000_011).copy()
    data[2_000_000:2_250_000] = 0
    data[4_000_000:4_050_000] = 255
    ctx = hb.Context(0)
    assert _decode(hb, ctx, data)[0] == 1
    ctx.close()


def test_fused_refuted_speculation_falls_back(hb):
    ctx = _ctx_with(hb, HB_DEBUG_SPOIL_SPECULATION="1")
    data = G.zipf(3_000_017)
    path, _ = _decode(hb, ctx, data)
    assert path == 2
    # the context remembers the code set: its next stream goes straight to the two-pass decoder (no second lost attempt) ...
    assert _decode(hb, ctx, data)[0] == 0
    # ... while another code set still gets its attempt
    assert _decode(hb, ctx, G.english(1_000_003))[0] == 2
    ctx.close()


def test_fused_skewed_and_near_uniform_trees(hb):
    ctx = hb.Context(0)
    rng = np.random.default_rng(5)
    # two letters + rare third (1-bit code: ~950 letters per subsequence), and a 200-letter near-uniform alphabet
    skew = rng.choice(np.array([65, 66, 67], np.uint8), size=2_000_003, p=[0.9, 0.09, 0.01])
    flat = rng.integers(0, 200, size=1_500_007).astype(np.uint8)
    assert _decode(hb, ctx, skew)[0] in (0, 1)
    # codes of 7 and 8 bits only: such sets resynchronise slowly, the library routes them to the two-pass decoder
    assert _decode(hb, ctx, flat)[0] == 0
    lop = rng.choice(np.arange(5, dtype=np.uint8), size=3_000_001, p=[0.6, 0.2, 0.1, 0.06, 0.04])
    assert _decode(hb, ctx, lop)[0] == 1
    ctx.close()


def test_fused_device_buffers_unaligned_output_and_capacity(hb):
    import torch
    from huff_encoding_b200.engine import Engine
    eng = make_engine()
    data = G.zipf(2_345_679)
    d = torch.from_numpy(data).to(dev())
    comp, n, pad, tree = eng.compress(d)
    base = torch.empty(data.size + 256, dtype=torch.uint8, device=dev())
    for off in (0, 1, 7, 32, 45):
        out = base[off: off + data.size + 64]
        out.zero_()
        dec, m = eng.decompress(comp, n, pad, tree, out=out)
        assert m == data.size and torch.equal(dec[:m], d), off
        assert eng.ctx.last_decode_path()[0] == 1
    # an output buffer of exactly n bytes (no slack behind the last row)
    exact = torch.empty(data.size, dtype=torch.uint8, device=dev())
    dec, m = eng.decompress(comp, n, pad, tree, out=exact)
    assert m == data.size and torch.equal(dec[:m], d)
    # a buffer that is too small: the needed size is reported and a second call with a larger buffer works
    small = torch.empty(1000, dtype=torch.uint8, device=dev())
    dec, m = eng.decompress(comp, n, pad, tree, out=small)
    assert m == data.size and torch.equal(dec[:m], d)


def test_serial_repair_of_every_chunk_stays_bounded(hb):
    # worst case of the two-pass decoder: EVERY chunk's speculative entry is wrong (debug switch), so dec_fix_kernel walks all
    # of them one after the other.  It must stay exact and its cost must stay linear and small per chunk (a chunk is 32 KiB
    # of stream): ~1000 chunks here, a generous 20 ms per chunk as the bound.
    import time
    ctx = _ctx_with(hb, HB_DEBUG_SPOIL_SPECULATION="1")
    data = G.zipf(48 << 20)
    comp, pad, tree = O.compress(data)
    ours = hb.HuffTree.from_weights(hb.build_weights_map(data, ctx=ctx))
    cd = hb.CompressData(comp, pad, ours)
    hb.decompress(cd, ctx=ctx)                                   # warm-up (buffers, tables)
    t0 = time.perf_counter()
    got = hb.decompress(cd, ctx=ctx)
    dt = time.perf_counter() - t0
    assert np.array_equal(got, data)
    n_chunks = comp.size // 32768
    assert ctx.last_decode_repairs() >= n_chunks // 2       # (a blind guess happens to be a code boundary now and then)
    assert dt < 0.02 * n_chunks, (dt, n_chunks)
    ctx.close()


def test_host_api_slab_pipelined_general_decode(hb):
    # >= 64 MiB of stream through hb_decompress_u8_into: slabs of the stream are copied, decoded (fused kernel, entry of a
    # slab = exit of the one before) and copied back concurrently
    ctx = hb.Context(0)
    for gen, n in (("zipf", (120 << 20) + 12345), ("english", (130 << 20) + 1)):
        data = getattr(G, gen)(n)
        cd = hb.compress(data, ctx=ctx)
        assert cd.comp_bytes().size >= (64 << 20)
        out = np.empty(n + 7, dtype=np.uint8)
        back = hb.decompress(cd, ctx=ctx, out=out)
        assert back.size == n and np.array_equal(back, data)
        assert ctx.last_decode_path()[0] == 1
        small = np.empty(n - 1000, dtype=np.uint8)           # too small: the exact size is reported, nothing is lost
        with pytest.raises(Exception):
            hb.decompress(cd, ctx=ctx, out=small)
    ctx.close()


def test_fused_wide_table_for_codes_of_13_and_14_bits(hb):
    # Zipf(1.5): the rare letters get 13- and 14-bit codes -> the kernel instance with the 14-bit emit table
    ctx = hb.Context(0)
    data = G.zipf(5_000_003, s=(15, 10))
    comp, pad, tree = O.compress(data)
    longest = max(len(c) for c in tree.codes().values())
    assert 12 < longest <= 14, longest
    assert _decode(hb, ctx, data)[0] == 1
    ctx.close()
