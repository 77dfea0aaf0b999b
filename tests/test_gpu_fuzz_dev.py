"""Randomised test of the DEVICE-buffer entry points (what bench.py and the sharded path drive): encode at a random start
bit into a buffer full of stale bytes, shard decode with a known entry (fused decoder when the tree allows it) into
buffers of random alignment and exact or generous capacity, and decode of an arbitrary bit range cut in two with a
speculative entry on the right part.  Everything against the oracle's stream and the input, bit for bit.  Seeded."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import oracle as O
from tests._model import dev, dev_sync, make_engine
from tests.test_gpu_fuzz import _case


@pytest.fixture(scope="module")
def eng():
    from huff_encoding_b200 import build
    build.build()
    return make_engine()


def _data(rng):
    d = _case(rng)
    if rng.random() < 0.35:                                   # longer inputs: many decoder chunks
        big = int(rng.integers(300_000, 2_500_000))
        reps = big // max(d.size, 1) + 1
        d = np.tile(d, reps)[:big]
        rng.shuffle(d)
    return d


@pytest.mark.parametrize("seed", range(int(os.environ.get("HB_FUZZ_DEV_SEEDS", "6"))))
def test_fuzz_device_api(eng, seed):
    import torch
    rng = np.random.default_rng(9100 + seed)
    for it in range(12):
        data = _data(rng)
        n = data.size
        tag = f"seed={seed} it={it} n={n} distinct={np.unique(data).size}"
        comp, pad, otree = O.compress(data)
        bits = comp.size * 8 - pad
        ref_bits = np.unpackbits(comp)[:bits]
        tree = eng.tree_from_weights(np.bincount(data, minlength=256))
        assert tree.read_codes() == otree.codes(), tag

        # ---- encode at a random start bit into stale memory (16-byte aligned letters, as the header requires)
        sb = int(rng.integers(0, 32))
        d = torch.from_numpy(data).to(dev())
        cap = (sb + bits + 7) // 8
        buf = torch.full((cap + 64,), int(rng.integers(0, 256)), dtype=torch.uint8, device=dev())
        tb = torch.zeros(1, dtype=torch.int64, device=dev())
        eng.histogram(d)
        eng.encode(d, tree, buf, start_bit=sb, total_bits=tb)
        eng.sync()
        dev_sync()
        assert int(tb.item()) == bits, tag
        got_bits = np.unpackbits(buf[:cap].cpu().numpy())
        assert np.array_equal(got_bits[sb: sb + bits], ref_bits), tag + f" start_bit={sb}"
        if sb >= 8:
            assert not got_bits[: sb // 8 * 8].any(), tag + " leading bytes of the start word"

        # ---- shard decode with the known entry, output at a random alignment, exact or generous capacity
        off = int(rng.integers(0, 64))
        slack = int(rng.choice([0, 0, 1, 31, 64]))
        base = torch.full((n + 128 + slack,), 0xEE, dtype=torch.uint8, device=dev())
        out = base[off: off + n + slack]
        # stream_bit0 = position of buffer bit 0 in the "whole stream": the shard starts at buffer bit sb, so it must be
        # congruent to -sb modulo the gcd of the code lengths (the phase of the speculative entries)
        g = max(int(tree.raw.len_gcd), 1)
        stream_bit0 = g * int(rng.integers(64, 1 << 20)) - sb
        e, x, cnt = eng.decode_shard(buf, sb + bits, sb, sb + bits, stream_bit0, tree, sb, out)
        dev_sync()
        assert cnt == n and e == sb, tag + f" decode_shard sb={sb} off={off} slack={slack}"
        assert np.array_equal(out[:n].cpu().numpy(), data), tag + f" decode_shard sb={sb} off={off} slack={slack}"
        assert bool((base[:off] == 0xEE).all()) and bool((base[off + n + slack:] == 0xEE).all()), tag + " wrote outside the buffer"
        # (bytes of the buffer behind the letters are the decoder's scratch: include/huffb200.h, hb_decompress_u8_dev)

        # ---- an arbitrary cut: left part with the known entry, right part speculative; the chain must close
        if bits > 64:
            lens = np.array([tree.raw.code_len[b] for b in range(256)], dtype=np.int64)[data]
            starts = np.concatenate([[0], np.cumsum(lens)]) + sb
            cut = sb + int(rng.integers(1, bits))
            eL, xL, cL = eng.decode_count(buf, sb + bits, sb, cut, g * 64 - sb, tree, entry_bit=sb)
            first_right = int(np.searchsorted(starts[:-1], cut, side="left"))
            assert cL == first_right and xL == (starts[first_right] if first_right < n else sb + bits), tag + f" cut={cut}"
            left = torch.empty(cL + 32, dtype=torch.uint8, device=dev())
            eng.decode_write(left)
            eR, xR, cR = eng.decode_count(buf, sb + bits, cut, sb + bits, g * 64 - sb, tree, entry_bit=-1)
            if eR != xL:                                        # speculation refuted by the neighbour: redo with the truth
                eR, xR, cR = eng.decode_count(buf, sb + bits, cut, sb + bits, g * 64 - sb, tree, entry_bit=xL)
            right = torch.empty(cR + 32, dtype=torch.uint8, device=dev())
            eng.decode_write(right)
            dev_sync()
            assert cL + cR == n, tag + f" cut={cut}"
            assert np.array_equal(np.concatenate([left[:cL].cpu().numpy(), right[:cR].cpu().numpy()]), data), tag + f" cut={cut}"
