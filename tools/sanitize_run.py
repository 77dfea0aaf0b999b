"""Small end-to-end cases for compute-sanitizer (one tool per gpurun call):
compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import huff_encoding_b200 as hb
from huff_encoding_b200 import datagen as G

cases = [("zipf", G.zipf(1_500_007)), ("english", G.english(700_001)), ("uniform", G.uniform(600_005)),
         ("single", np.full(100_003, 7, np.uint8)), ("tiny", G.zipf(5))]
fib = [1, 1]
while len(fib) < 36:
    fib.append(fib[-1] + fib[-2])
w = np.zeros(256, dtype=np.uint64); w[10:46] = fib
cases.append(("fib35", G.from_weights_runs(w)[:3_000_000]))
for name, data in cases:
    cd = hb.compress(data)
    back = hb.decompress(cd)
    assert np.array_equal(back, data), name
    print("ok", name, data.size, cd.comp_bytes().size, flush=True)
os.environ["HB_NO_FASTPATH"] = "1"
ctx = hb.Context(0)
d = G.uniform(400_003)
assert np.array_equal(hb.decompress(hb.compress(d, ctx=ctx), ctx=ctx), d)
print("ok uniform general path", flush=True)
