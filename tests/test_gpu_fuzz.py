"""Randomised differential test: the CUDA path (through the C ABI) against the oracle on many small random inputs --
random alphabets and skews (heavy ties), lengths around tile / chunk / group borders, foreign trees, truncated streams.
Seeded, so a failure is reproducible from the printed case."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from tests._model import dev, dev_sync, make_engine

BORDERS = [1, 2, 31, 32, 33, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097, 8191, 8192, 8193, 32767, 32768, 32769,
           65535, 65536, 65537, 262143, 262144, 262145]


def _case(rng):
    n_sym = int(rng.choice([1, 2, 3, 4, 5, 8, 16, 17, 40, 100, 200, 256]))
    alphabet = rng.choice(256, size=n_sym, replace=False)
    shape = rng.choice(["flat", "geometric", "zipf", "steps"])
    if shape == "flat":
        p = np.ones(n_sym)
    elif shape == "geometric":
        p = rng.uniform(0.3, 0.9) ** np.arange(n_sym)
    elif shape == "zipf":
        p = 1.0 / np.arange(1, n_sym + 1) ** rng.uniform(0.5, 2.5)
    else:
        p = rng.integers(1, 4, size=n_sym).astype(float)
    p = p / p.sum()
    n = int(rng.choice(BORDERS)) if rng.random() < 0.6 else int(rng.integers(1, 400_000))
    data = rng.choice(alphabet, size=n, p=p).astype(np.uint8)
    return data


@pytest.mark.parametrize("seed", range(int(os.environ.get("HB_FUZZ_SEEDS", "12"))))   # HB_FUZZ_SEEDS=200 for a long soak
def test_fuzz_compress_decompress_parity(seed):
    import huff_encoding_b200 as hb
    rng = np.random.default_rng(7000 + seed)
    for it in range(25):
        data = _case(rng)
        tag = f"seed={seed} it={it} n={data.size} distinct={np.unique(data).size}"
        comp, pad, tree = O.compress(data)
        cd = hb.compress(data)
        assert cd.huff_tree().read_codes() == tree.codes(), tag
        assert cd.padding_bits() == pad and np.array_equal(cd.comp_bytes(), comp), tag
        assert np.array_equal(hb.decompress(cd), data), tag
        # decode a truncated stream with arbitrary padding: trailing partial code is dropped like the reference does
        if comp.size > 2:
            cut = int(rng.integers(1, min(comp.size, 64)))
            p2 = int(rng.integers(0, 8))
            want = O.decompress(comp[: comp.size - cut], p2, tree)
            got = hb.decompress(hb.CompressData(comp[: comp.size - cut], p2, cd.huff_tree()))
            assert np.array_equal(got, want), tag + f" cut={cut} pad={p2}"
        # compress_with_tree with a tree built from OTHER data over a superset alphabet
        if it % 5 == 0:
            other = np.concatenate([data, np.arange(256, dtype=np.uint8)])
            ftree = hb.HuffTree.from_weights(hb.build_weights_map(other))
            cd2 = hb.compress_with_tree(data, ftree)
            c2, p2 = O.compress_with_tree(data, O.tree_from_weights(O.histogram(other)))
            assert cd2.padding_bits() == p2 and np.array_equal(cd2.comp_bytes(), c2), tag + " foreign tree"
            assert np.array_equal(hb.decompress(cd2), data), tag + " foreign tree"
