"""Three-way differential test of the container and tree parsers (comp.rs:128-184, 279-300; tree_inner.rs:522-668):

    product host code (csrc/hb_tree.cpp through the C ABI)  vs  oracle/huff_oracle.c  vs  oracle/py_restatement.py

The C oracle and the product's parser are both iterative C and were written side by side; the Python restatement follows the
reference's own shape (a recursive descent over an iterator of bits, string bit vectors) and shares no code or structure
with either, so agreement of all three is not common-mode evidence.  Compared: accept/reject, the reference's error
message or panic message, and (payload, padding, letter -> code) when accepted."""
import numpy as np
import pytest

from huff_encoding_b200.api import (CompressData, CompressedDataFromBytesError, FromBinError, HuffPanic, HuffTree,
                                    TreeTooLargeError)
from oracle import oracle as O
from oracle import py_restatement as P


def _bits_of(raw: bytes, n_bits: int) -> str:
    return "".join(format(b, "08b") for b in raw)[:n_bits]


def _random_tree_bits(rng, max_leaves=40):
    out, leaves, stack = [], 0, 1
    while stack:
        stack -= 1
        if leaves + stack < max_leaves - 1 and rng.random() < 0.55:
            out.append("1")
            stack += 2
        else:
            out.append("0" + format(int(rng.integers(0, 256)), "08b"))
            leaves += 1
    return "".join(out)


def _pack(bits: str):
    padded = bits + "0" * ((8 - len(bits) % 8) % 8)
    return bytes(int(padded[i:i + 8], 2) for i in range(0, len(padded), 8))


def test_try_from_bin_three_way():
    rng = np.random.default_rng(4242)
    seen = {"ok": 0, "small": 0, "big": 0}
    for it in range(1500):
        kind = it % 3
        if kind == 0:
            bits = _random_tree_bits(rng)
        elif kind == 1:                                          # damaged: cut, extended or bit-flipped
            bits = _random_tree_bits(rng)
            r = rng.random()
            if r < 0.4:
                bits = bits[: int(rng.integers(0, len(bits) + 1))]
            elif r < 0.8:
                bits += "".join(str(int(b)) for b in rng.integers(0, 2, size=int(rng.integers(1, 20))))
            else:
                k = int(rng.integers(0, len(bits)))
                bits = bits[:k] + ("1" if bits[k] == "0" else "0") + bits[k + 1:]
        else:
            bits = "".join(str(int(b)) for b in rng.integers(0, 2, size=int(rng.integers(0, 300))))
        raw = _pack(bits)
        # python restatement
        try:
            py = ("ok", P.read_codes(P.try_from_bin(bits)))
        except P.FromBinError as e:
            py = ("err", str(e) + "<u8>")                        # Display appends the letter type (tree_inner.rs:680-682)
        # C oracle
        try:
            co = ("ok", O.tree_from_bin(np.frombuffer(raw, np.uint8), len(bits)).codes())
        except O.OracleError:
            co = ("err", None)
        # product
        try:
            pr = ("ok", HuffTree.try_from_bin(raw, len(bits)).read_codes())
        except FromBinError as e:
            pr = ("err", str(e))
        except TreeTooLargeError:
            continue                                             # (documented capacity of hb_tree; not generated here)
        assert pr[0] == py[0] == co[0], (it, bits)
        if pr[0] == "ok":
            assert pr[1] == py[1] == co[1], (it, bits)
            seen["ok"] += 1
        else:
            assert pr[1] == py[1], (it, bits, pr[1], py[1])      # same message: too small / too big
            seen["small" if "small" in pr[1] else "big"] += 1
    assert min(seen.values()) > 50, seen


def test_container_three_way():
    rng = np.random.default_rng(777)
    seen = {}
    for it in range(1200):
        root = P.try_from_bin(_random_tree_bits(rng))
        payload = bytes(rng.integers(0, 256, size=int(rng.integers(1, 40)), dtype=np.uint8))
        pad = int(rng.integers(0, 8))
        good = P.to_bytes(payload, pad, root)
        if it % 4 == 0:
            blob = good
        elif it % 16 == 5:
            blob = good[: 5 + int.from_bytes(good[1:5], "big")]     # header and tree, no compressed data at all
        elif it % 4 == 3:
            blob = bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8))
        else:
            b = bytearray(good)
            for _ in range(int(rng.integers(1, 3))):
                k = int(rng.integers(0, 5))
                if k == 0:
                    del b[int(rng.integers(0, len(b) + 1)):]
                elif k == 1 and b:
                    b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                elif k == 2 and b:
                    b[0] = int(rng.integers(0, 256))             # both padding nibbles
                elif k == 3 and len(b) >= 5:
                    b[1:5] = int(rng.integers(0, 70)).to_bytes(4, "big")     # tree length
                else:
                    b += bytes(rng.integers(0, 256, size=int(rng.integers(1, 6)), dtype=np.uint8))
            blob = bytes(b)
        try:
            c, p, r = P.try_from_bytes(blob)
            py = ("ok", (c, p, P.read_codes(r)))
        except P.FromBytesError as e:
            py = ("err", str(e))
        except P.Panic as e:
            py = ("panic", str(e))
        try:
            c, p, t = O.try_from_bytes(np.frombuffer(blob, dtype=np.uint8))
            co = ("ok", (c.tobytes(), p, t.codes()))
        except O.OracleError:
            co = ("no", None)
        try:
            cd = CompressData.try_from_bytes(blob)
            pr = ("ok", (cd.comp_bytes().tobytes(), cd.padding_bits(), cd.huff_tree().read_codes()))
        except CompressedDataFromBytesError as e:
            pr = ("err", e.message())
        except HuffPanic as e:
            pr = ("panic", str(e))
        except TreeTooLargeError:
            continue
        assert pr == py, (it, blob.hex(), pr, py)
        assert (co[0] == "ok") == (pr[0] == "ok"), (it, blob.hex())
        if pr[0] == "ok":
            assert co[1] == pr[1], (it, blob.hex())
        seen[pr[0] + ":" + (pr[1] if pr[0] != "ok" else "")] = seen.get(pr[0] + ":" + (pr[1] if pr[0] != "ok" else ""), 0) + 1
    # every outcome of comp.rs:128-184 was met
    for key in ("ok:", "err:slice is empty", "err:slice too short to read tree length", "err:slice too short to read tree",
                "err:invalid tree in slice", "panic:stored tree length must be at least 2",
                "panic:provided comp_bytes are empty", "panic:padding bits cannot be larger than 7"):
        assert seen.get(key, 0) > 0, (key, seen)


def test_to_bytes_three_way_and_the_reference_doctest_blob():
    # comp.rs:219-262
    comp, pad, root = P.compress(b"abbccc")
    assert P.to_bytes(comp, pad, root).hex() == "370000000498e61310bc00"
    rng = np.random.default_rng(5)
    for _ in range(100):
        data = bytes(rng.choice(rng.integers(0, 256, size=int(rng.integers(1, 30))), size=int(rng.integers(1, 300))).astype(np.uint8))
        comp, pad, root = P.compress(data)
        blob = P.to_bytes(comp, pad, root)
        w = {}
        for b in data:
            w[b] = w.get(b, 0) + 1
        t = HuffTree.from_weights(dict(sorted(w.items())))
        assert CompressData(comp, pad, t).to_bytes() == blob
