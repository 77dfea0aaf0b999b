"""How often does the cross-CTA speculation of the count pass miss?  usage: python tools/repairs_check.py [bytes]
Prints, per workload, the number of 32 KiB chunks whose leading thread had to be repaired (hb_ctx_last_decode_repairs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from huff_encoding_b200 import datagen as G
from huff_encoding_b200.engine import Engine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
eng = Engine(0)
for name, gen in (("zipf", lambda n: G.zipf(n, device=eng.device)), ("english", lambda n: G.english(n, device=eng.device)),
                  ("zipf15", lambda n: G.zipf(n, device=eng.device, s=(15, 10)))):
    d = gen(size)
    torch.cuda.synchronize()
    out, n, pad, tree = eng.compress(d)
    dec, m = eng.decompress(out, n, pad, tree)
    torch.cuda.synchronize()
    assert m == size and torch.equal(dec[:m], d)
    print(f"{name}: {size} B -> {n} B, {(n + 32767) // 32768} chunks, repairs = {eng.ctx.last_decode_repairs()}")
