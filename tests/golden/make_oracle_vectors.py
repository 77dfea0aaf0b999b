"""Regenerates tests/golden/oracle_vectors.json: digests of the ORACLE's output on seeded inputs.

These are NOT reference outputs (the Rust crate cannot run here); they freeze the oracle's behaviour -- including the
BinaryHeap tie-break restatement -- so that any later change to the oracle or to the CUDA path shows up as a diff.
    python tests/golden/make_oracle_vectors.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from huff_encoding_b200 import datagen as G
from oracle import oracle as O

CASES = [("english", 1 << 20), ("zipf", 300_007), ("uniform", 200_003), ("english", 4097), ("zipf", 31)]


def make():
    out = []
    for kind, n in CASES:
        data = getattr(G, kind)(n)
        comp, pad, tree = O.compress(data)
        tb, nb = O.tree_as_bin(tree)
        out.append({"workload": kind, "n": n, "seed": "datagen default", "comp_len": int(comp.size), "padding_bits": int(pad),
                    "comp_sha256": hashlib.sha256(comp.tobytes()).hexdigest(),
                    "tree_bin_bits": int(nb), "tree_bin_sha256": hashlib.sha256(tb.tobytes()).hexdigest(),
                    "blob_sha256": hashlib.sha256(O.to_bytes(comp, pad, tree).tobytes()).hexdigest()})
    # a tie-heavy histogram: 256 letters, weights from a tiny range
    rng = np.random.default_rng(20261018)
    w = rng.integers(1, 4, size=256).astype(np.uint64)
    t = O.tree_from_weights(w)
    out.append({"workload": "tie_heavy_weights", "weights_seed": 20261018,
                "lens": [int(x) for x in t.lens()], "codes_sha256": hashlib.sha256(json.dumps(t.codes(), sort_keys=True).encode()).hexdigest()})
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")
    json.dump(make(), open(path, "w"), indent=1)
    print("wrote", path)
