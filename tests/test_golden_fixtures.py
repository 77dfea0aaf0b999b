"""Committed oracle fixtures (tests/golden/oracle_vectors.json, made by tests/golden/make_oracle_vectors.py):
the oracle must keep reproducing them (CPU), and so must the CUDA path through the C ABI (GPU)."""
import hashlib
import json
import os

import numpy as np
import pytest

from huff_encoding_b200 import datagen as G
from oracle import oracle as O

VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))
STREAMS = [v for v in VEC if "comp_sha256" in v]


@pytest.mark.parametrize("v", STREAMS, ids=lambda v: f"{v['workload']}-{v['n']}")
def test_oracle_reproduces_fixture(v):
    comp, pad, tree = O.compress(getattr(G, v["workload"])(v["n"]))
    assert (comp.size, pad) == (v["comp_len"], v["padding_bits"])
    assert hashlib.sha256(comp.tobytes()).hexdigest() == v["comp_sha256"]
    assert hashlib.sha256(O.to_bytes(comp, pad, tree).tobytes()).hexdigest() == v["blob_sha256"]


def test_host_tree_reproduces_tie_heavy_fixture():
    from huff_encoding_b200.api import HuffTree
    v = [x for x in VEC if x["workload"] == "tie_heavy_weights"][0]
    w = np.random.default_rng(v["weights_seed"]).integers(1, 4, size=256).astype(np.uint64)
    t = HuffTree.from_weights({b: int(w[b]) for b in range(256)})
    codes = t.read_codes()
    assert [len(codes[b]) for b in range(256)] == v["lens"]
    assert hashlib.sha256(json.dumps({int(k): c for k, c in codes.items()}, sort_keys=True).encode()).hexdigest() == v["codes_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("v", STREAMS, ids=lambda v: f"{v['workload']}-{v['n']}")
def test_cuda_path_reproduces_fixture(v):
    import huff_encoding_b200 as hb
    cd = hb.compress(getattr(G, v["workload"])(v["n"]))
    assert (cd.comp_bytes().size, cd.padding_bits()) == (v["comp_len"], v["padding_bits"])
    assert hashlib.sha256(cd.comp_bytes().tobytes()).hexdigest() == v["comp_sha256"]
    assert hashlib.sha256(cd.to_bytes()).hexdigest() == v["blob_sha256"]
