"""Small driver for ncu: one warm round trip + one measured round trip of the device-resident path.
usage: python tools/prof_run.py [workload] [size_bytes] [rounds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from huff_encoding_b200 import datagen as G
from huff_encoding_b200.engine import Engine

workload = sys.argv[1] if len(sys.argv) > 1 else "uniform"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 256 << 20
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = Engine(0)
d = (G.zipf(size, device=eng.device, s=(15, 10)) if workload == "zipf15" else getattr(G, workload)(size, device=eng.device))
torch.cuda.synchronize()
dst = torch.empty(d.numel() + 64, dtype=torch.uint8, device=eng.device)
for r in range(rounds):
    out, n, pad, tree = eng.compress(d)
    dec, m = eng.decompress(out, n, pad, tree, out=dst)
    torch.cuda.synchronize()
    assert m == size and torch.equal(dec[:m], d)
print("prof_run ok", workload, size, n)
