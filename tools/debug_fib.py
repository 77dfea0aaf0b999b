import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from huff_encoding_b200 import datagen as G
from huff_encoding_b200.engine import Engine
from oracle import oracle as O
eng = Engine(0)
w = G.fibonacci_weights()
d = G.from_weights_runs(w, device=eng.device)
n = d.numel()
out, clen, pad, tree = eng.compress(d)
dec = torch.full((n + 64,), 7, dtype=torch.uint8, device=eng.device)
dec, m = eng.decompress(out, clen, pad, tree, out=dec)
print("n", n, "m", m, "clen", clen, "pad", pad, "n%32", n % 32, "out_addr%32", dec.data_ptr() % 32)
bad = torch.nonzero(dec[:n] != d).flatten()
print("mismatches", bad.numel())
if bad.numel():
    b = bad.cpu().numpy()
    print("first", b[:10], "last", b[-10:])
    # contiguous runs of mismatches
    runs = np.split(b, np.nonzero(np.diff(b) != 1)[0] + 1)
    print("n runs", len(runs), [(int(r[0]), int(r[-1]), len(r)) for r in runs[:12]])
    i = int(b[0])
    print("around first:", dec[i - 4:i + 12].cpu().numpy(), d[i - 4:i + 12].cpu().numpy())
# suffix property of the stream against the oracle
hist = eng.histogram(d).cpu().numpy().astype(np.uint64)
rt = O.tree_from_weights(hist)
K = 1 << 16
tail = d[n - K:].cpu().numpy()
ct, pt = O.compress_with_tree(tail, rt)
tb = np.unpackbits(ct)[: ct.size * 8 - pt]
sb = np.unpackbits(out[clen - (tb.size // 8 + 2): clen].cpu().numpy())
sb = sb[: sb.size - pad]
print("stream suffix ok:", np.array_equal(sb[-tb.size:], tb))
