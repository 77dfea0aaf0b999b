"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/huffb200.h declares, and its
host-only entry points (tree, code table, serialisation) agree with the oracle.  No GPU compute is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from huff_encoding_b200 import _lib as L
from huff_encoding_b200 import build as hb_build
from huff_encoding_b200.api import (CompressData, CompressedDataFromBytesError, FromBinError, HuffCudaError, HuffPanic,
                                    HuffTree, ByteWeights)
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    hb_build.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "huffb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = C.CDLL(L.SO_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/huffb200.h but not exported"
    assert declared == {s[0] for s in L.SYMBOLS}, "ctypes binding and header disagree"
    assert L.load().hb_version() == 1


def test_struct_layout_matches_header():
    assert C.sizeof(L.HbNode) == 16
    assert C.sizeof(L.HbTree) == 24 + 16 * 513 + 256 + 512 + 2048
    assert L.HbTree.code.offset % 8 == 0


def _tree_pair(w, mode_hb, mode_o):
    t = L.HbTree()
    w = np.ascontiguousarray(w, dtype=np.uint64)
    st = L.load().hb_tree_from_weights(w.ctypes.data_as(C.POINTER(C.c_uint64)), mode_hb, C.byref(t))
    assert st == 0
    return HuffTree(t), O.tree_from_weights(w, mode_o)


@pytest.mark.parametrize("seed", range(60))
def test_host_tree_matches_oracle_on_tie_heavy_histograms(seed):
    rng = np.random.default_rng(1000 + seed)
    w = np.zeros(256, dtype=np.uint64)
    k = int(rng.integers(1, 257))
    idx = rng.choice(256, size=k, replace=False)
    hi = [2, 4, 16, 1000, 1 << 40][seed % 5]                  # small ranges force equal weights
    w[idx] = rng.integers(1, hi + 1, size=k).astype(np.uint64)
    for mode_hb, mode_o in ((L.HB_ORDER_ASC, O.ORDER_ASC), (L.HB_ORDER_BYTEWEIGHTS, O.ORDER_BYTEWEIGHTS)):
        ours, ref = _tree_pair(w, mode_hb, mode_o)
        assert ours.read_codes() == ref.codes()
        assert ours.raw.n_nodes == ref.n_nodes and ours.raw.root == ref.root
        for i in range(ref.n_nodes):
            a, b = ours.raw.nodes[i], ref.nodes[i]
            assert (a.left, a.right, a.weight) == (b.left, b.right, b.weight)
            if a.left == L.HB_NO_CHILD:
                assert a.letter == b.letter
        assert ours.as_bin() == (O.tree_as_bin(ref)[0].tobytes(), O.tree_as_bin(ref)[1])
        lens = ref.lens()
        assert ours.raw.max_len == lens.max() and ours.raw.min_len == lens[lens > 0].min()
        assert ours.raw.len_gcd == np.gcd.reduce(lens[lens > 0])


def test_host_tree_fibonacci_40_bit_codes():
    from huff_encoding_b200.datagen import fibonacci_weights
    w = fibonacci_weights()
    assert int(w.sum()) == 1836311750
    ours, ref = _tree_pair(w, L.HB_ORDER_ASC, O.ORDER_ASC)
    assert ours.raw.max_len == 40 and ours.raw.min_len == 1
    assert ours.read_codes() == ref.codes()


def test_reference_goldens_through_the_abi():
    # tree_inner.rs:623-628 and lib.rs:54-55
    assert HuffTree.from_weights({ord("a"): 1, ord("b"): 2, ord("c"): 3}).as_bin_string() == \
        "[10011000, 11100110, 00010011, 00010]"
    assert HuffTree.from_weights({0xFF: 3, 0xAA: 2, 0xCC: 1}).as_bin_string() == "[10111111, 11101100, 11000101, 01010]"
    # tests/tree_init.rs:10-47
    t = HuffTree.from_weights(list(enumerate([5, 9, 12, 13, 16, 45])))
    assert [t.code_str(i) for i in range(6)] == ["1100", "1101", "100", "101", "111", "0"]
    # tests/tree_init.rs:49-64
    t = HuffTree.from_weights({0xF4: 78})
    assert t.root_letter() == 0xF4 and t.code_str(0xF4) == "0"
    # tests/tree_init.rs:66-69
    with pytest.raises(HuffPanic, match="provided empty weights"):
        HuffTree.from_weights({})


def test_container_format_golden_and_errors():
    # comp.rs:219-262: compress(b"abbccc").to_bytes()
    t = HuffTree.from_weights({ord("a"): 1, ord("b"): 2, ord("c"): 3})
    cd = CompressData(bytes([0b10111100, 0]), 7, t)
    blob = cd.to_bytes()
    assert blob.hex() == "3700000004" + "98e61310" + "bc00"
    back = CompressData.try_from_bytes(blob)
    assert back.padding_bits() == 7 and back.comp_bytes().tobytes() == b"\xbc\x00"
    assert back.huff_tree().read_codes() == {ord("a"): "10", ord("b"): "11", ord("c"): "0"}
    with pytest.raises(CompressedDataFromBytesError, match="slice is empty"):
        CompressData.try_from_bytes(b"")
    with pytest.raises(CompressedDataFromBytesError, match="tree length"):
        CompressData.try_from_bytes(blob[:3])
    with pytest.raises(CompressedDataFromBytesError, match="too short to read tree"):
        CompressData.try_from_bytes(blob[:7])
    with pytest.raises(HuffPanic, match="at least 2"):
        CompressData.try_from_bytes(bytes([0x37, 0, 0, 0, 1, 0x98]))
    with pytest.raises(HuffPanic, match="comp_bytes are empty"):
        CompressData.try_from_bytes(blob[:9])
    with pytest.raises(CompressedDataFromBytesError, match="invalid tree"):
        CompressData.try_from_bytes(bytes([0x07, 0, 0, 0, 2, 0xFF, 0xFF, 0x00]))
    with pytest.raises(HuffPanic, match="larger than 7"):
        CompressData(b"\x00", 8, t)
    with pytest.raises(FromBinError):
        HuffTree.try_from_bin(b"", 0)          # tests/tree_bin.rs:28-32


def test_tree_bin_roundtrip_matches_oracle():
    rng = np.random.default_rng(5)
    for _ in range(20):
        w = np.zeros(256, dtype=np.uint64)
        idx = rng.choice(256, size=int(rng.integers(1, 257)), replace=False)
        w[idx] = rng.integers(1, 50, size=idx.size).astype(np.uint64)
        ours, ref = _tree_pair(w, L.HB_ORDER_BYTEWEIGHTS, O.ORDER_BYTEWEIGHTS)
        b, n = ours.as_bin()
        again = HuffTree.try_from_bin(b, n)
        assert again.read_codes() == ours.read_codes() == O.tree_from_bin(np.frombuffer(b, np.uint8), n).codes()


def test_byteweights_iterator_quirk_is_reproduced():
    bw = ByteWeights()
    bw.weights[0] = 3
    bw._len = 1
    assert list(bw) == [(0, 3), (0, 3)]                  # weights.rs:396-415 wrap-around
    assert HuffTree.from_weights(bw).read_codes() == {0: "1"}
    bw.compat = False
    assert list(bw) == [(0, 3)]
    assert HuffTree.from_weights(bw).read_codes() == {0: "0"}


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from huff_encoding_b200 import compress
    with pytest.raises(HuffCudaError):
        compress(b"abbccc")


def test_shard_plan_matches_the_per_shard_arithmetic():
    """hb_shard_plan = tree of the summed histograms + sum_b hist_g[b] * len[b] for every shard."""
    import ctypes as C
    from huff_encoding_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(11)
    for G_ in (1, 2, 3, 8):
        h = rng.integers(0, 5000, size=(G_, 256)).astype(np.uint64)
        h[:, rng.integers(0, 256, size=40)] = 0                       # some letters absent everywhere or in some shards
        t, t_ref = L.HbTree(), L.HbTree()
        bits = (C.c_uint64 * G_)()
        assert lib.hb_shard_plan(h.ctypes.data_as(C.POINTER(C.c_uint64)), G_, L.HB_ORDER_ASC, C.byref(t), bits) == L.HB_OK
        total = np.ascontiguousarray(h.sum(axis=0))
        assert lib.hb_tree_from_weights(total.ctypes.data_as(C.POINTER(C.c_uint64)), L.HB_ORDER_ASC, C.byref(t_ref)) == L.HB_OK
        assert bytes(t) == bytes(t_ref)
        lens = np.frombuffer(t.code_len, dtype=np.uint16).astype(np.uint64)
        assert list(bits) == [int(x) for x in (h * lens[None, :]).sum(axis=1)]
    empty = np.zeros((2, 256), dtype=np.uint64)
    t = L.HbTree()
    bits = (C.c_uint64 * 2)()
    assert lib.hb_shard_plan(empty.ctypes.data_as(C.POINTER(C.c_uint64)), 2, L.HB_ORDER_ASC, C.byref(t), bits) == L.HB_ERR_EMPTY_WEIGHTS


def test_try_from_bin_differential_on_random_bit_strings():
    """HuffTree::try_from_bin (tree_inner.rs:522-604) on arbitrary input: the product's host parser and the oracle accept
    and reject the same bit strings and, when they accept, read the same codes (duplicates, lone roots, deep chains)."""
    rng = np.random.default_rng(2024)
    accepted = rejected = 0
    for it in range(600):
        if it % 3 == 0:                                   # well-formed preorder strings of a random full binary tree
            bits = []
            pending = 1
            leaves = 0
            while pending:
                pending -= 1
                if leaves + pending < 200 and rng.random() < 0.48:
                    bits.append(1)
                    pending += 2
                else:
                    bits.append(0)
                    bits.extend(int(b) for b in np.unpackbits(np.array([rng.integers(0, 256)], dtype=np.uint8)))
                    leaves += 1
            if rng.random() < 0.3:                        # truncate or extend some of them
                cut = int(rng.integers(0, len(bits) + 1))
                bits = bits[:cut] if rng.random() < 0.5 else bits + [int(b) for b in rng.integers(0, 2, size=cut % 17)]
        else:
            bits = [int(b) for b in rng.integers(0, 2, size=int(rng.integers(0, 400)))]
        n_bits = len(bits)
        raw = np.packbits(np.array(bits, dtype=np.uint8)) if n_bits else np.zeros(0, dtype=np.uint8)
        try:
            ref = O.tree_from_bin(raw, n_bits)
        except O.OracleError:
            ref = None
        try:
            ours = HuffTree.try_from_bin(raw.tobytes(), n_bits)
        except FromBinError:
            ours = None
        assert (ours is None) == (ref is None), (it, n_bits)
        if ours is not None:
            accepted += 1
            assert ours.read_codes() == ref.codes(), (it, n_bits)
        else:
            rejected += 1
    assert accepted > 50 and rejected > 50


def test_container_differential_on_random_blobs():
    """CompressData::try_from_bytes (comp.rs:128-184) and to_bytes (:279-300): the product's host code and the oracle
    classify random and mutated blobs alike and agree on (payload, padding, codes) when they accept."""
    rng = np.random.default_rng(77)
    t = HuffTree.from_weights({b: int(w) for b, w in zip(rng.choice(256, 40, replace=False), rng.integers(1, 1000, 40))})
    good = CompressData(rng.integers(0, 256, size=50, dtype=np.uint8).tobytes(), 3, t).to_bytes()
    agree_ok = agree_err = 0
    for it in range(500):
        if it % 2 == 0:
            blob = bytearray(good)
            for _ in range(int(rng.integers(1, 4))):
                k = int(rng.integers(0, 4))
                if k == 0 and len(blob) > 1:
                    del blob[int(rng.integers(0, len(blob))):]
                elif k == 1:
                    blob[int(rng.integers(0, len(blob)))] ^= 1 << int(rng.integers(0, 8))
                elif k == 2:
                    blob[0:5] = bytes(rng.integers(0, 256, size=5, dtype=np.uint8))
                else:
                    blob += bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
            blob = bytes(blob)
        else:
            blob = bytes(rng.integers(0, 256, size=int(rng.integers(0, 64)), dtype=np.uint8))
        try:
            payload, pad, tree = O.try_from_bytes(np.frombuffer(blob, dtype=np.uint8))
            ref = (payload.tobytes(), pad, tree.codes())
        except O.OracleError:
            ref = None
        try:
            cd = CompressData.try_from_bytes(blob)
            ours = (cd.comp_bytes().tobytes(), cd.padding_bits(), cd.huff_tree().read_codes())
        except (CompressedDataFromBytesError, HuffPanic):
            ours = None
        assert (ours is None) == (ref is None), (it, blob.hex())
        if ours is not None:
            assert ours == ref, (it, blob.hex())
            agree_ok += 1
        else:
            agree_err += 1
    assert agree_ok > 20 and agree_err > 20


def _tree_bits(spec):
    """preorder bit string of a nested (left, right) / int-letter structure (tree_inner.rs:637-663)"""
    if isinstance(spec, int):
        return "0" + format(spec, "08b")
    return "1" + _tree_bits(spec[0]) + _tree_bits(spec[1])


def _pack(bits):
    n = len(bits)
    bits = bits + "0" * ((8 - n % 8) % 8)
    return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)), n


def test_foreign_tree_with_duplicates_up_to_the_node_cap_and_beyond():
    """N3: try_from_bin accepts any preorder tree (tree_inner.rs:522-604).  hb_tree holds 513 nodes = 257 leaves: a
    right-leaning chain with 257 leaves (letters repeat) parses and matches the oracle; 258 leaves is refused with the
    dedicated status, not truncated and not mistaken for a malformed tree."""
    from huff_encoding_b200.api import TreeTooLargeError

    def chain(n_leaves):
        spec = (n_leaves - 1) % 7
        for k in range(n_leaves - 2, -1, -1):
            spec = (k % 7, spec)
        return spec

    raw, n = _pack(_tree_bits(chain(257)))
    ours = HuffTree.try_from_bin(raw, n)
    ref = O.tree_from_bin(np.frombuffer(raw, np.uint8), n)
    assert ours.read_codes() == ref.codes()
    assert len(ours.read_codes()) == 7                      # duplicates: the last DFS visit wins (tree_inner.rs:396)
    raw, n = _pack(_tree_bits(chain(258)))
    with pytest.raises(TreeTooLargeError):
        HuffTree.try_from_bin(raw, n)
