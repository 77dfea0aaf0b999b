"""Pins the CPU oracle against every golden vector the reference's tests/doctests hold for the u8 path
 and cross-checks it against the independent pure-Python restatement.

Reads like the reference's own tests: huff_coding/tests/{tree_init,tree_bin,comp_decomp}.rs.
"""
import itertools
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import py_restatement as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))

Q_RSQRT = (b"float Q_rsqrt( float number )\n    {\n        long i;\n        float x2, y;\n"
           b"        const float threehalfs = 1.5F;\n    \n        x2 = number * 0.5F;\n        y  = number;\n"
           b"        i  = * ( long * ) &y;                       // evil floating point bit level hacking\n"
           b"        i  = 0x5f3759df - ( i >> 1 );               // what the fuck? \n"
           b"        y  = * ( float * ) &i;\n"
           b"        y  = y * ( threehalfs - ( x2 * y * y ) );   // 1st iteration\n"
           b"    //\ty  = y * ( threehalfs - ( x2 * y * y ) );   // 2nd iteration, this can be removed\n"
           b"    \n        return y;\n    }")


def _bytes_of(case):
    if "input_ascii" in case:
        return case["input_ascii"].encode()
    if "input_hex" in case:
        return bytes.fromhex(case["input_hex"])
    return bytes(case["input_bytes"])


# ---------------------------------------------------------------- B1 / B2: as_bin goldens
@pytest.mark.parametrize("key", ["B1_as_bin_abbccc", "B2_as_bin_ff_aa_cc"])
def test_as_bin_golden(key):
    g = GOLD[key]
    data = _bytes_of(g)
    # reference: HuffTree::from_weights(ByteWeights::from_bytes(..)).as_bin().to_string()
    for order in (O.ORDER_BYTEWEIGHTS, O.ORDER_ASC):
        t = O.tree_from_weights(O.histogram(data), order)
        b, n = O.tree_as_bin(t)
        assert O.bin_to_string(b, n) == g["as_bin_string"]
    # independent restatement
    root = P.tree_from_weights(P.ByteWeights(data))
    s = P.as_bin(root)
    assert "[" + ", ".join(s[i:i + 8] for i in range(0, len(s), 8)) + "]" == g["as_bin_string"]


# ---------------------------------------------------------------- B3: compress(b"abbccc").to_bytes()
def test_to_bytes_golden():
    g = GOLD["B3_to_bytes_abbccc"]
    data = _bytes_of(g)
    comp, pad, t = O.compress(data)
    blob = O.to_bytes(comp, pad, t)
    assert blob[0] == g["byte0"] == 0x37
    assert int.from_bytes(bytes(blob[1:5]), "big") == g["tree_len"]
    assert {chr(k): v for k, v in t.codes().items()} == g["codes"]
    assert list(blob[9:]) == g["data_bytes"] == [0b10111100, 0b00000000]
    assert bytes(blob).hex() == "3700000004" + "98e61310" + "bc00"
    # try_from_bytes(to_bytes()) round trip (comp.rs:105-116)
    comp2, pad2, t2 = O.try_from_bytes(blob)
    assert pad2 == pad and bytes(comp2) == bytes(comp) and t2.codes() == t.codes()
    assert bytes(O.decompress(comp2, pad2, t2)) == data


# ---------------------------------------------------------------- B4: three-letter code goldens
@pytest.mark.parametrize("case", GOLD["B4_codes_three_letters"]["cases"])
def test_codes_three_letters(case):
    data = _bytes_of(case)
    want = {ord(k): v for k, v in case["codes"].items()}
    assert O.tree_from_weights(O.histogram(data), O.ORDER_BYTEWEIGHTS).codes() == want
    assert O.tree_from_weights(O.histogram(data), O.ORDER_ASC).codes() == want
    assert P.read_codes(P.tree_from_weights(P.ByteWeights(data))) == want


# ---------------------------------------------------------------- B5: tests/tree_init.rs::tree_normal_init
def test_tree_normal_init_all_orders():
    g = GOLD["B5_tree_normal_init"]
    weights, codes = g["weights"], g["codes"]
    # the reference inserts from a HashMap (random order); no ties => every order must give the goldens
    for perm in itertools.permutations(range(6)):
        letters = [perm_i for perm_i in perm]
        t = O.tree_from_pairs(letters, [weights[i] for i in perm])
        assert [t.code_str(i) for i in range(6)] == codes
    root = P.tree_from_weights([(i, w) for i, w in enumerate(weights)])
    assert [P.read_codes(root)[i] for i in range(6)] == codes


# ---------------------------------------------------------------- B6: tests/tree_init.rs::tree_single_branch
def test_tree_single_branch():
    t = O.tree_from_pairs([0xF4], [78])          # -12i8 as a byte
    root = t.nodes[t.root]
    assert root.left == O.HO_NONE and root.letter == 0xF4      # letter branch is the root
    assert t.code_str(0xF4) == "0"
    comp, pad = O.compress_with_tree(bytes([0xF4] * 11), t)
    assert bytes(comp) == b"\x00\x00" and pad == 5
    assert bytes(O.decompress(comp, pad, t)) == bytes([0xF4] * 11)


# ---------------------------------------------------------------- B7: tests/tree_init.rs::tree_invalid_weights
def test_tree_invalid_weights():
    with pytest.raises(O.OracleError) as e:
        O.tree_from_weights(np.zeros(256, dtype=np.uint64))
    assert e.value.code == O.ERR_EMPTY_WEIGHTS
    with pytest.raises(ValueError, match="provided empty weights"):
        P.tree_from_weights([])


# ---------------------------------------------------------------- B8: tests/comp_decomp.rs::compress_decompress
def test_compress_decompress():
    comp, pad, t = O.compress(Q_RSQRT)
    assert bytes(O.decompress(comp, pad, t)) == Q_RSQRT
    c2, p2, r2 = P.compress(Q_RSQRT)
    assert bytes(comp) == c2 and pad == p2
    assert P.decompress(c2, p2, r2) == Q_RSQRT


# ---------------------------------------------------------------- B9: weights.rs doctests
def test_count_goldens():
    w = O.histogram(b"fffff")
    assert w[ord("f")] == 5 and np.count_nonzero(w) == 1
    for b, f in P.ByteWeights(bytes([0, 1, 1, 2, 2, 2])):
        assert b == f - 1                       # weights.rs:156-159 (holds even with the wrap-around re-yield of 0)
    w = O.histogram(b"aabbb") + O.histogram(b"aaabbc")
    assert (w[ord("a")], w[ord("b")], w[ord("c")]) == (5, 5, 1)


# ---------------------------------------------------------------- B10: tests/tree_bin.rs
def test_tree_from_bin_roundtrip():
    data = GOLD["B10_tree_bin_texts"]["input_ascii"].encode()
    t = O.tree_from_weights(O.histogram(data), O.ORDER_BYTEWEIGHTS)
    b, n = O.tree_as_bin(t)
    t2 = O.tree_from_bin(b, n)
    assert t2.codes() == t.codes()
    assert n == 10 * np.count_nonzero(O.histogram(data)) - 1


def test_tree_bin_invalid_vec():
    with pytest.raises(O.OracleError) as e:          # tests/tree_bin.rs:28-32
        O.tree_from_bin(b"", 0)
    assert e.value.code == O.ERR_BIN_TOO_SMALL
    t = O.tree_from_weights(O.histogram(b"abbccc"))
    b, n = O.tree_as_bin(t)
    with pytest.raises(O.OracleError) as e:          # too big: tree_inner.rs:586-590
        O.tree_from_bin(np.concatenate([b, np.zeros(1, np.uint8)]), n + 3)
    assert e.value.code == O.ERR_BIN_TOO_BIG
    with pytest.raises(O.OracleError) as e:          # truncated
        O.tree_from_bin(b, n - 2)
    assert e.value.code == O.ERR_BIN_TOO_SMALL


# ---------------------------------------------------------------- error conventions
def test_missing_letter_error():
    t = O.tree_from_weights(O.histogram(b"abb"))
    with pytest.raises(O.OracleError) as e:          # comp.rs:399-415 doctest
        O.compress_with_tree(b"abbccc", t)
    assert e.value.code == O.ERR_MISSING_LETTER and e.value.missing == ord("c")


def test_compressdata_new_panics():
    t = O.tree_from_weights(O.histogram(b"ab"))
    with pytest.raises(O.OracleError) as e:
        O.decompress(b"", 0, t)
    assert e.value.code == O.ERR_EMPTY_COMP
    with pytest.raises(O.OracleError) as e:
        O.decompress(b"\x00", 8, t)
    assert e.value.code == O.ERR_BAD_PADDING


# ---------------------------------------------------------------- ByteWeights wrap-around quirk
def test_byteweights_wrap_quirk():
    assert list(P.ByteWeights(bytes([0, 0, 0]))) == [(0, 3), (0, 3)]
    assert list(P.ByteWeights(bytes([0, 255]))) == [(0, 1), (255, 1)]
    assert list(P.ByteWeights(bytes([1, 2, 2]))) == [(1, 1), (2, 2)]
    # [0,0,0]: compress() gives code "0"; from_weights(ByteWeights) builds two leaves and byte 0 encodes as "1"
    assert O.tree_from_weights(O.histogram(bytes(3)), O.ORDER_ASC).codes() == {0: "0"}
    assert O.tree_from_weights(O.histogram(bytes(3)), O.ORDER_BYTEWEIGHTS).codes() == {0: "1"}
    assert P.read_codes(P.tree_from_weights(P.ByteWeights(bytes(3)))) == {0: "1"}


# ---------------------------------------------------------------- C oracle == Python restatement on tie-heavy inputs
@pytest.mark.parametrize("seed", range(40))
def test_c_oracle_matches_python_restatement(seed):
    rng = np.random.default_rng(seed)
    n_sym = int(rng.integers(1, 40))
    alphabet = rng.choice(256, size=n_sym, replace=False)
    # small counts => many equal weights => the heap tie-break paths are exercised
    data = rng.choice(alphabet, size=int(rng.integers(1, 400))).astype(np.uint8).tobytes()
    for o_mode, p_mode in ((O.ORDER_ASC, "asc"), (O.ORDER_BYTEWEIGHTS, "byteweights")):
        comp, pad, t = O.compress(data, o_mode) if o_mode == O.ORDER_ASC else (None, None, None)
        if o_mode == O.ORDER_BYTEWEIGHTS:
            t = O.tree_from_weights(O.histogram(data), o_mode)
            comp, pad = O.compress_with_tree(data, t)
        c2, p2, root = P.compress(data, p_mode)
        assert t.codes() == P.read_codes(root)
        assert bytes(comp) == c2 and pad == p2
        assert bytes(O.decompress(comp, pad, t)) == data == P.decompress(c2, p2, root)
        b, n = O.tree_as_bin(t)
        assert "".join(f"{x:08b}" for x in b)[:n] == P.as_bin(root)


def test_c_oracle_matches_python_all_256_equal_weights():
    data = bytes(range(256)) * 3
    comp, pad, t = O.compress(data)
    c2, p2, root = P.compress(data, "asc")
    assert bytes(comp) == c2 and pad == p2 == 0 and len(c2) == 768
    assert set(t.lens()) == {8}


# ---------------------------------------------------------------- stream invariants
@pytest.mark.parametrize("n", [1, 7, 8, 9, 1000])
def test_single_symbol_stream(n):
    comp, pad, t = O.compress(bytes([0x41]) * n)
    assert bytes(comp) == bytes((n + 7) // 8) and pad == (8 - n % 8) % 8
    assert bytes(O.decompress(comp, pad, t)) == bytes([0x41]) * n
