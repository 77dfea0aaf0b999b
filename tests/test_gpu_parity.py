"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.
Bit-exact (all work is integer).  Reads like the reference's tests/comp_decomp.rs, plus the edge cases of
SURVEY.md section 8d.  Nothing here reads /root/reference."""
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


@pytest.fixture(scope="module")
def eng():
    from huff_encoding_b200.engine import Engine
    return make_engine()


def _first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    return f"sizes {a.size} vs {b.size}; first diff at {int(d[0]) if d.size else None}" + \
        (f": {a[d[0]]:#x} vs {b[d[0]]:#x}; diffs={d.size}" if d.size else "")


def _assert_compress_parity(hb, data):
    cd = hb.compress(data)
    comp, pad, tree = O.compress(data)
    assert cd.huff_tree().read_codes() == tree.codes()
    assert cd.padding_bits() == pad
    assert np.array_equal(cd.comp_bytes(), comp), _first_diff(cd.comp_bytes(), comp)
    out = hb.decompress(cd)
    assert np.array_equal(out, np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data), \
        _first_diff(out, data)
    return cd


# ---------------------------------------------------------------- histogram (a1/a2)
@pytest.mark.parametrize("n", [1, 15, 16, 17, 31, 4095, 4096, 4097, (1 << 20) + 3])
def test_histogram_sizes(hb, n):
    data = G.uniform(n, seed=n)
    w = hb.build_weights_map(data)
    ref = O.histogram(data)
    assert w == {b: int(ref[b]) for b in range(256) if ref[b]}


def test_histogram_distributions_and_alignment(eng):
    import torch
    for name, data in (("uniform", G.uniform(3_000_001)), ("english", G.english(2_000_003)),
                       ("single", np.full(5_000_000, 0x41, np.uint8)), ("zipf", G.zipf(1_000_000))):
        ref = O.histogram(data).astype(np.int64)
        base = torch.from_numpy(np.concatenate([np.zeros(32, np.uint8), data])).to(eng.device)
        for off in (32, 33, 35, 47):                     # 16-byte aligned and not
            view = base[off:]
            exp = ref.copy()
            exp[0] += 0
            skipped = data[: off - 32]
            for b in skipped:
                exp[b] -= 1
            got = eng.histogram(view).cpu().numpy()
            assert np.array_equal(got, exp), (name, off)


def test_byteweights_from_bytes(hb):
    bw = hb.ByteWeights.from_bytes(b"fffff")                      # weights.rs:148-150
    assert bw.get(ord("f")) == 5 and bw.len() == 1
    for byte, weight in hb.ByteWeights.from_bytes(bytes([0, 1, 1, 2, 2, 2])):    # weights.rs:156-159
        assert byte == weight - 1
    a = hb.ByteWeights.from_bytes(b"aabbb")
    a += hb.ByteWeights.from_bytes(b"aaabbc")                     # weights.rs:165-172
    assert (a.get(ord("a")), a.get(ord("b")), a.get(ord("c"))) == (5, 5, 1)
    data = G.english(100_000)
    t = hb.ByteWeights.threaded_from_bytes(data, 12)
    assert t == hb.ByteWeights.from_bytes(data)                   # no byte 0 in text: the quirk is silent


# ---------------------------------------------------------------- compress / decompress round trips (a8-a11)
def test_reference_round_trip_snippet(hb):
    from tests.test_oracle_golden import Q_RSQRT
    cd = _assert_compress_parity(hb, Q_RSQRT)                     # tests/comp_decomp.rs
    assert hb.decompress(cd).tobytes() == Q_RSQRT


def test_doc_example_abbccc(hb):
    cd = hb.compress(b"abbccc")                                   # comp.rs:219-262
    assert cd.to_bytes().hex() == "3700000004" + "98e61310" + "bc00"
    again = hb.CompressData.try_from_bytes(cd.to_bytes())         # comp.rs:105-116
    assert hb.decompress(again).tobytes() == b"abbccc"


def test_config0_english_1mib(hb):
    _assert_compress_parity(hb, G.english(1 << 20))


@pytest.mark.parametrize("gen,n", [("uniform", (4 << 20) + 5), ("zipf", (4 << 20) + 1), ("english", (3 << 20) + 7)])
def test_distributions_multi_tile(hb, gen, n):
    _assert_compress_parity(hb, getattr(G, gen)(n))


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 8193, 65536 + 17])
def test_ragged_lengths(hb, n):
    _assert_compress_parity(hb, G.zipf(n, seed=n))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1 << 20])
@pytest.mark.parametrize("byte", [0x41, 0x00, 0xFF])
def test_single_symbol(hb, n, byte):
    cd = _assert_compress_parity(hb, np.full(n, byte, np.uint8))
    assert cd.comp_bytes().tobytes() == bytes((n + 7) // 8) and cd.padding_bits() == (8 - n % 8) % 8


def test_two_symbols(hb):
    h = G.uniform(1 << 20, seed=77)
    _assert_compress_parity(hb, np.where(h & 1, 0x42, 0x41).astype(np.uint8))
    _assert_compress_parity(hb, np.where((h & 3) == 0, 0x42, 0x41).astype(np.uint8))     # 3:1 skew


def test_eight_equal_symbols_three_bit_codes_do_not_self_synchronise(hb):
    # 8 equiprobable letters -> all codes 3 bits: subsequence boundaries are not code boundaries and the decoder can
    # only rely on the gcd alignment, never on resynchronisation.
    data = (G.uniform(3_000_000, seed=5) & 7).astype(np.uint8) * 31
    cd = _assert_compress_parity(hb, data)
    assert set(len(c) for c in cd.huff_tree().read_codes().values()) == {3}


@pytest.mark.parametrize("mask,L", [(3, 2), (15, 4), (255, 8)])
@pytest.mark.parametrize("n", [1_000_003, 4096 * 3, 17])
def test_fixed_length_code_sets_fast_path(hb, mask, L, n):
    # equiprobable 2^L letters -> perfect tree, every code L bits: the table-translation fast path (hb_fixed.cuh)
    base = np.tile(np.arange(mask + 1, dtype=np.uint8), 4)                 # make sure every letter is present
    data = np.concatenate([base, (G.uniform(n, seed=L) & mask).astype(np.uint8)]) * (255 // mask)
    if n == 17:
        data = np.tile(np.arange(mask + 1, dtype=np.uint8), 3)[: max(17, mask + 1) + 2 * (mask + 1)] * (255 // mask)
    cd = _assert_compress_parity(hb, data)
    lens = set(len(c) for c in cd.huff_tree().read_codes().values())
    if len(lens) == 1:
        assert lens == {L}


def test_general_path_on_uniform_when_fast_path_is_disabled(hb):
    import os
    os.environ["HB_NO_FASTPATH"] = "1"
    try:
        ctx = hb.Context(0)
    finally:
        del os.environ["HB_NO_FASTPATH"]
    data = G.uniform((2 << 20) + 77)
    cd = hb.compress(data, ctx=ctx)
    comp, pad, tree = O.compress(data)
    assert set(tree.lens()) == {8}
    assert cd.padding_bits() == pad and np.array_equal(cd.comp_bytes(), comp)
    assert np.array_equal(hb.decompress(cd, ctx=ctx), data)
    ctx.close()


def test_lengths_with_common_factor(hb):
    # weights 4,4,4,1,1,1,1 -> lengths {2,2,2,4,4,4,4}: gcd 2 but not fixed length
    rng = np.random.default_rng(3)
    data = rng.choice(np.arange(7, dtype=np.uint8), size=2_000_000, p=np.array([4, 4, 4, 1, 1, 1, 1]) / 16)
    cd = _assert_compress_parity(hb, data)
    assert cd.huff_tree().raw.len_gcd == 2


@pytest.mark.parametrize("s", [(15, 10), (20, 10)])
def test_zipf_long_tail_codes_of_13_to_16_bits(hb, s):
    # codes longer than the 12-bit first-level table are frequent enough that every decoder thread meets them
    data = G.zipf((6 << 20) + 11, seed=77, s=s)
    cd = _assert_compress_parity(hb, data)
    assert 12 < cd.huff_tree().raw.max_len <= 20


def test_many_second_level_slots_take_the_global_table(hb):
    # 5 geometric letters + 251 equally likely ones: 246 codes of 13 bits under 123 different depth-12 nodes -- more
    # slots than the compact shared-memory copy of the second level holds, so the write pass reads it from global memory
    p = np.zeros(256)
    p[:5] = [0.5, 0.25, 0.125, 0.0625, 0.03125]
    p[5:] = (1 - p[:5].sum()) / 251
    data = np.random.default_rng(3).choice(np.arange(256, dtype=np.uint8), size=3_000_000, p=p)
    cd = _assert_compress_parity(hb, data)
    codes = cd.huff_tree().read_codes()
    assert len({c[:12] for c in codes.values() if len(c) > 12}) > 96


def test_geometric_tail_codes_in_all_three_decoder_levels(hb):
    # P(k) ~ 0.75^k over 64 letters: code lengths 2..~26, so the first-level table, the second-level table (13..20
    # bits) and the bit-serial walk (> 20 bits) are all hit, the first two often
    rng = np.random.default_rng(21)
    p = 0.75 ** np.arange(64)
    data = rng.choice(np.arange(64, dtype=np.uint8) * 3 + 5, size=5_000_003, p=p / p.sum())
    cd = _assert_compress_parity(hb, data)
    assert cd.huff_tree().raw.max_len > 20


def test_fibonacci_wide_codes(hb):
    # 36 Fibonacci-weighted letters: longest code 35 bits (> 32 -> wide encoder path), N = F(38)-1 ~ 39 M letters
    fib = [1, 1]
    while len(fib) < 36:
        fib.append(fib[-1] + fib[-2])
    w = np.zeros(256, dtype=np.uint64)
    w[10:46] = fib
    data = G.from_weights_runs(w)
    cd = _assert_compress_parity(hb, data)
    assert cd.huff_tree().raw.max_len == 35
    # scattered variant: same histogram, rare letters interleaved with common ones
    rng = np.random.default_rng(9)
    perm = rng.permutation(data.size)
    _assert_compress_parity(hb, data[perm])


def test_all_256_letters_tie_heavy(hb):
    data = np.tile(np.arange(256, dtype=np.uint8), 4099)
    cd = _assert_compress_parity(hb, data)
    assert cd.padding_bits() == 0 and cd.comp_bytes().size == data.size


def test_empty_input_panics_like_the_reference(hb):
    with pytest.raises(hb.HuffPanic, match="provided empty weights"):
        hb.compress(b"")                                          # tree_inner.rs:283-285 via comp.rs:354


# ---------------------------------------------------------------- compress_with_tree (a9)
def test_compress_with_tree_byteweights(hb):
    data = b"abbccc"                                              # comp.rs:385-392
    tree = hb.HuffTree.from_weights(hb.ByteWeights.from_bytes(data))
    cd = hb.compress_with_tree(data, tree)
    assert hb.decompress(cd).tobytes() == data
    comp, pad = O.compress_with_tree(data, O.tree_from_weights(O.histogram(data), O.ORDER_BYTEWEIGHTS))
    assert cd.comp_bytes().tobytes() == comp.tobytes() and cd.padding_bits() == pad


def test_compress_with_tree_byteweights_quirk_duplicate_zero_leaf(hb):
    data = np.concatenate([np.zeros(1000, np.uint8), G.english(5000)])         # has 0x00, no 0xFF
    tree = hb.HuffTree.from_weights(hb.ByteWeights.from_bytes(data))
    ref_tree = O.tree_from_weights(O.histogram(data), O.ORDER_BYTEWEIGHTS)
    assert tree.read_codes() == ref_tree.codes() and tree.raw.n_leaves == ref_tree.n_nodes // 2 + 1
    cd = hb.compress_with_tree(data, tree)
    comp, pad = O.compress_with_tree(data, ref_tree)
    assert np.array_equal(cd.comp_bytes(), comp) and cd.padding_bits() == pad
    assert np.array_equal(hb.decompress(cd), data)


def test_compress_with_tree_missing_letter(hb):
    tree = hb.HuffTree.from_weights(hb.ByteWeights.from_bytes(b"abb"))
    with pytest.raises(hb.CompressError) as e:                    # comp.rs:399-415
        hb.compress_with_tree(b"abbzccc", tree)
    assert e.value.missing_letter() == ord("z") and e.value.message() == "letter not found in codes"


def test_compress_with_foreign_tree(hb):
    # a tree built from other data that still covers the alphabet: stream differs from compress() but round-trips
    tree = hb.HuffTree.from_weights(hb.build_weights_map(G.english(50_000, seed=1)))
    data = G.english(300_000, seed=2)
    cd = hb.compress_with_tree(data, tree)
    ref_tree = O.tree_from_weights(O.histogram(G.english(50_000, seed=1)))
    comp, pad = O.compress_with_tree(data, ref_tree)
    assert np.array_equal(cd.comp_bytes(), comp) and cd.padding_bits() == pad
    assert np.array_equal(hb.decompress(cd), data)


# ---------------------------------------------------------------- decompress of oracle-made streams (a10)
def test_decompress_trailing_partial_code_is_dropped(hb):
    data = G.zipf(100_000)
    comp, pad, tree = O.compress(data)
    ours = hb.HuffTree.from_weights(hb.build_weights_map(data))
    for cut, p in ((1, 0), (2, 3), (5, 7), (1, 5)):
        c2 = comp[: comp.size - cut]
        want = O.decompress(c2, p, tree)
        got = hb.decompress(hb.CompressData(c2, p, ours))
        assert np.array_equal(got, want), _first_diff(got, want)


def test_decompress_garbage_stream_matches_oracle_walk(hb):
    # any bit string decodes under a complete prefix code; compare against the bit-serial reference walk
    tree_data = G.english(20_000)
    tree = O.tree_from_weights(O.histogram(tree_data))
    ours = hb.HuffTree.from_weights(hb.build_weights_map(tree_data))
    junk = G.uniform(700_001, seed=99)
    for pad in (0, 6):
        want = O.decompress(junk, pad, tree)
        got = hb.decompress(hb.CompressData(junk, pad, ours))
        assert np.array_equal(got, want), _first_diff(got, want)


def test_decompress_single_leaf_tree_emits_a_letter_per_bit(hb):
    tree = hb.HuffTree.from_weights({0x5A: 3})
    junk = G.uniform(10_001, seed=3)                              # comp.rs:496,506-509: bit values are ignored
    got = hb.decompress(hb.CompressData(junk, 5, tree))
    assert got.size == junk.size * 8 - 5 and (got == 0x5A).all()


def test_compressdata_new_invariants(hb):
    tree = hb.HuffTree.from_weights({1: 1, 2: 1})
    with pytest.raises(hb.HuffPanic, match="comp_bytes are empty"):
        hb.CompressData(b"", 0, tree)
    with pytest.raises(hb.HuffPanic, match="larger than 7"):
        hb.CompressData(b"\x00", 8, tree)


# ---------------------------------------------------------------- device-resident API (what bench.py times)
def test_device_api_matches_oracle_and_start_bit(eng):
    import torch
    data = G.zipf((2 << 20) + 13)
    d = torch.from_numpy(data).to(eng.device)
    out, n, pad, tree = eng.compress(d)
    comp, pad_o, _ = O.compress(data)
    assert n == comp.size and pad == pad_o
    assert np.array_equal(out[:n].cpu().numpy(), comp)
    dec, m = eng.decompress(out, n, pad, tree)
    assert m == data.size and torch.equal(dec[:m], d)
    # a shard that starts mid-word: same bits shifted right by start_bit, first bits left zero
    for sb in (1, 7, 13, 31):
        buf = torch.zeros(n + 8, dtype=torch.uint8, device=eng.device)
        tb = torch.zeros(1, dtype=torch.int64, device=eng.device)
        eng.encode(d, tree, buf, start_bit=sb, total_bits=tb)
        eng.sync()
        bits = np.unpackbits(buf.cpu().numpy())
        ref_bits = np.unpackbits(comp)[: comp.size * 8 - pad_o]
        assert int(tb.item()) == ref_bits.size
        assert not bits[:sb].any()
        assert np.array_equal(bits[sb: sb + ref_bits.size], ref_bits)
        assert not bits[sb + ref_bits.size: (sb + ref_bits.size + 7) // 8 * 8].any()


def test_device_input_pointer_alignment(eng):
    # the encoder takes 256-bit loads when the letters are 32-byte aligned and 128-bit loads otherwise; the device API
    # requires 16-byte alignment (include/huffb200.h) and says so instead of misreading
    import torch
    from huff_encoding_b200 import api
    data = G.zipf((3 << 20) + 77, seed=5)
    comp, pad_o, _ = O.compress(data)
    big = torch.zeros(data.size + 256, dtype=torch.uint8, device=eng.device)
    base = (-big.data_ptr()) % 32                      # first 32-byte aligned element
    for off in (base, base + 16, base + 48):
        view = big[off: off + data.size]
        view.copy_(torch.from_numpy(data))
        assert view.data_ptr() % 32 == (off - base) % 32
        out, n, pad, tree = eng.compress(view)
        assert n == comp.size and pad == pad_o
        assert np.array_equal(out[:n].cpu().numpy(), comp)
    view = big[base + 1: base + 1 + data.size]
    with pytest.raises((api.HuffCudaError, ValueError, RuntimeError)):
        eng.compress(view)


def test_shard_decode_building_blocks(eng):
    import torch
    data = G.english(3_000_000)
    d = torch.from_numpy(data).to(eng.device)
    out, n, pad, tree = eng.compress(d)
    total_bits = n * 8 - pad
    lens = np.array([tree.raw.code_len[b] for b in range(256)], dtype=np.int64)[data]
    starts = np.concatenate([[0], np.cumsum(lens)])               # code-word start of every letter
    cuts = [0, (total_bits // 3) // 8 * 8, (2 * total_bits // 3) // 8 * 8, total_bits]
    got = []
    for g in range(3):
        e, x, cnt = eng.decode_count(out, total_bits, cuts[g], cuts[g + 1], 0, tree, entry_bit=0 if g == 0 else -1)
        first = int(np.searchsorted(starts[:-1], cuts[g], side="left"))
        last = int(np.searchsorted(starts[:-1], cuts[g + 1], side="left"))
        assert e == starts[first] and cnt == last - first
        assert x == (starts[last] if last < data.size else total_bits)
        piece = torch.empty(cnt + 16, dtype=torch.uint8, device=eng.device)
        eng.decode_write(piece)
        got.append(piece[:cnt].cpu().numpy())
    assert np.array_equal(np.concatenate(got), data)


# ---------------------------------------------------------------- BASELINE.json sizes: size-independent properties
def test_config1_one_gib_uniform_properties(eng):
    import torch
    n = 1 << 30
    d = G.uniform(n, device=eng.device)
    hist = eng.histogram(d).cpu().numpy()
    assert hist.sum() == n
    assert np.array_equal(hist, torch.bincount(d.to(torch.int32), minlength=256).cpu().numpy())
    out, clen, pad, tree = eng.compress(d)
    lens = np.array([tree.raw.code_len[b] for b in range(256)], dtype=np.int64)
    bits = int((hist * lens).sum())
    assert clen == (bits + 7) // 8 and pad == (8 - bits % 8) % 8          # SURVEY A.6 invariants
    ref_tree = O.tree_from_weights(hist.astype(np.uint64))
    assert tree.read_codes() == ref_tree.codes()
    # prefix property: the stream of a prefix of the input is a prefix of the stream
    k = 8 << 20
    comp_k, _ = O.compress_with_tree(d[:k].cpu().numpy(), ref_tree)
    kb = int(lens[d[:k].cpu().numpy()].sum())
    assert np.array_equal(out[: kb // 8].cpu().numpy(), comp_k[: kb // 8])
    dec, m = eng.decompress(out, clen, pad, tree)
    assert m == n and torch.equal(dec[:m], d)


# ---------------------------------------------------------------- SURVEY 8f "next" rows
def test_n3_foreign_tree_deeper_than_64_bits(hb):
    # a foreign tree (as try_from_bin may deliver it) whose codes reach 79 bits: the decoder walks any depth; the
    # encoder refuses letters whose code exceeds 64 bits
    fib = [1, 1]
    while len(fib) < 80:
        fib.append(fib[-1] + fib[-2])
    pairs = [(b, fib[b]) for b in range(80)]
    tree = hb.HuffTree.from_weights(pairs)
    otree = O.tree_from_pairs([p[0] for p in pairs], [p[1] for p in pairs])
    assert tree.read_codes() == otree.codes() and tree.raw.max_len == 79
    rng = np.random.default_rng(4)
    deep = rng.integers(0, 80, size=50_000).astype(np.uint8)             # every depth, long codes everywhere
    comp, pad = O.compress_with_tree(deep, otree)
    got = hb.decompress(hb.CompressData(comp, pad, tree))
    assert np.array_equal(got, deep), _first_diff(got, deep)
    blob = hb.CompressData(comp, pad, tree).to_bytes()                    # through the container as a foreign file
    again = hb.CompressData.try_from_bytes(blob)
    assert np.array_equal(hb.decompress(again), deep)
    shallow = rng.integers(20, 80, size=200_000).astype(np.uint8)         # codes <= 61 bits: the encoder takes them
    cd = hb.compress_with_tree(shallow, tree)
    c2, p2 = O.compress_with_tree(shallow, otree)
    assert np.array_equal(cd.comp_bytes(), c2) and cd.padding_bits() == p2
    with pytest.raises(RuntimeError, match="longer than 64"):
        hb.compress_with_tree(deep, tree)


def test_n2_cli_file_round_trip(hb, tmp_path):
    from huff_encoding_b200 import cli
    src = tmp_path / "notes.txt"
    data = np.concatenate([G.english(300_000), np.zeros(1000, np.uint8)])     # has 0x00, no 0xFF: ByteWeights quirk path
    data.tofile(src)
    assert cli.main(["-n", str(src), str(tmp_path / "notes.txt")]) == 0       # huff/src/cli.rs: ".hff" is appended
    hff = tmp_path / "notes.txt.hff"
    blob = np.fromfile(hff, dtype=np.uint8)
    # same container a library user gets: header byte, BE u32 tree length, tree, stream (huff/src/comp.rs:47-70)
    bw = hb.ByteWeights()
    bw += hb.ByteWeights.threaded_from_bytes(data, 12)
    tree = hb.HuffTree.from_weights(bw)
    ref_tree = O.tree_from_pairs([b for b, _ in bw], [w for _, w in bw])
    assert tree.read_codes() == ref_tree.codes()
    comp, pad = O.compress_with_tree(data, ref_tree)
    assert bytes(blob) == bytes(O.to_bytes(comp, pad, ref_tree))
    out = tmp_path / "back.txt"
    assert cli.main(["-d", "-n", str(hff), str(out)]) == 0
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), data)
    assert cli.parse_block_size("2G") == 2_000_000_000 and cli.parse_block_size("4Ki") == 4096


def test_cross_chunk_repair_path_when_speculation_is_wrong(hb):
    # the serial repair kernel only runs when a chunk's 1024-bit look-back fails to resynchronise, which real code sets
    # practically never do; the debug switch makes every chunk-leading thread guess without look-back instead
    import os
    os.environ["HB_DEBUG_SPOIL_SPECULATION"] = "1"
    try:
        ctx = hb.Context(0)
    finally:
        del os.environ["HB_DEBUG_SPOIL_SPECULATION"]
    for gen, n in (("zipf", 3_000_017), ("english", 2_500_003)):
        data = getattr(G, gen)(n)
        comp, pad, tree = O.compress(data)
        ours = hb.HuffTree.from_weights(hb.build_weights_map(data, ctx=ctx))
        got = hb.decompress(hb.CompressData(comp, pad, ours), ctx=ctx)
        assert np.array_equal(got, data), _first_diff(got, data)
        assert ctx.last_decode_repairs() > 10, "the repair path was not exercised"
    ctx.close()


def test_host_api_pipelined_fixed_length_decode_multi_chunk(hb):
    # > 32 MiB of stream: hb_decompress_u8 overlaps H2D / translation / D2H chunk by chunk (fixed-length code sets)
    for mask, n in ((255, (80 << 20) + 7), (15, (90 << 20) + 5)):
        data = (G.uniform(n, seed=mask) & mask).astype(np.uint8)
        cd = hb.compress(data)
        assert len(set(len(c) for c in cd.huff_tree().read_codes().values())) == 1
        assert cd.comp_bytes().size > (32 << 20)
        back = hb.decompress(cd)
        assert np.array_equal(back, data), _first_diff(back, data)
        out = np.empty(n + 3, dtype=np.uint8)                      # caller-owned buffer variant
        back2 = hb.decompress(cd, out=out)
        assert back2.size == n and np.array_equal(back2, data)


def test_n3_deepest_trees_hb_tree_can_hold(hb):
    # right-leaning chains as try_from_bin may deliver them: 256 distinct leaves (codes of 1 .. 255 bits) and 257 leaves with
    # one letter twice (the longest code hb_tree can hold: 256 bits).  Code words longer than a whole 1 024-bit subsequence's
    # look-back, spanning many threads' subsequences
    def chain_bits(letters):
        return "".join("1" + "0" + format(l, "08b") for l in letters[:-1]) + "0" + format(letters[-1], "08b")

    rng = np.random.default_rng(12)
    for letters in (list(range(256)), list(range(256)) + [7]):
        bits = chain_bits(letters)
        raw = np.packbits(np.array([int(b) for b in bits], dtype=np.uint8))
        tree = hb.HuffTree.try_from_bin(raw.tobytes(), len(bits))
        otree = O.tree_from_bin(raw, len(bits))
        assert tree.read_codes() == otree.codes()
        assert tree.raw.max_len == len(letters) - 1
        data = rng.integers(0, 256, size=30_000).astype(np.uint8)          # mean code length ~128 bits
        data[::97] = 255                                                    # the deepest leaves, regularly
        comp, pad = O.compress_with_tree(data, otree)
        got = hb.decompress(hb.CompressData(comp, pad, tree))
        assert np.array_equal(got, O.decompress(comp, pad, otree)), _first_diff(got, data)
        # (with the duplicated letter the round trip still returns the letters: both leaves carry the same letter)
        assert np.array_equal(got, data)
