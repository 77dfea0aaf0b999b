"""BASELINE.json's full sizes on one B200: properties that do not need the (slow) oracle on the whole input.
configs[2]: 4 GiB Zipf(1.2) (bit offsets beyond 2^32);  configs[4]: Fibonacci-256, 1.71 GiB, 40-bit codes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from huff_encoding_b200 import datagen as G
from oracle import oracle as O
from tests._model import dev, dev_sync, make_engine


@pytest.fixture(scope="module")
def eng():
    from huff_encoding_b200 import build
    build.build()
    from huff_encoding_b200.engine import Engine
    return make_engine()


def _check_stream_invariants(eng, d, prefix_letters):
    import torch
    n = d.numel()
    hist = eng.histogram(d).cpu().numpy()
    assert hist.sum() == n
    ref_hist = np.zeros(256, dtype=np.int64)
    step = 1 << 28
    for s in range(0, n, step):
        ref_hist += torch.bincount(d[s:s + step].to(torch.int32), minlength=256).cpu().numpy()
    assert np.array_equal(hist, ref_hist)
    out, clen, pad, tree = eng.compress(d)
    ref_tree = O.tree_from_weights(hist.astype(np.uint64))
    assert tree.read_codes() == ref_tree.codes()                       # same tree as the oracle builds
    lens = ref_tree.lens()
    bits = int((hist * lens).sum())
    assert clen == (bits + 7) // 8 and pad == (8 - bits % 8) % 8
    head = d[:prefix_letters].cpu().numpy()
    comp_k, _ = O.compress_with_tree(head, ref_tree)                   # prefix property against the oracle
    kb = int(lens[head].sum())
    assert np.array_equal(out[: kb // 8].cpu().numpy(), comp_k[: kb // 8])
    tail_bits = np.unpackbits(out[clen - 1: clen].cpu().numpy())
    assert not tail_bits[8 - pad:].any() if pad else True              # pad bits are zero
    dec, m = eng.decompress(out, clen, pad, tree)
    assert m == n
    for s in range(0, n, step):
        e = min(s + step, n)                                           # `dec` is allocated with slack past n
        assert torch.equal(dec[s:e], d[s:e]), f"round trip differs in [{s}, {e})"
    return tree, bits


def test_config2_four_gib_zipf(eng):
    n = 1 << 32
    d = G.zipf(n, device=eng.device)
    tree, bits = _check_stream_invariants(eng, d, 4 << 20)
    assert bits > (1 << 32) and 2 <= tree.raw.min_len and tree.raw.max_len <= 16


def test_config4_fibonacci_256_forty_bit_codes(eng):
    w = G.fibonacci_weights()
    assert int(w.sum()) == 1_836_311_750
    d = G.from_weights_runs(w, device=eng.device)                      # contiguous runs: the 224 rare letters first
    tree, bits = _check_stream_invariants(eng, d, 2 << 20)
    assert tree.raw.max_len == 40 and tree.raw.min_len == 1
