import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from huff_encoding_b200 import datagen as G
from huff_encoding_b200.engine import Engine
eng = Engine(0)
dev = eng.device
n_total, world = 1000000007, 8
d = G.zipf(n_total, device=dev)
out, clen, pad, tree = eng.compress(d)
total_bits = clen * 8 - pad
n_bytes = clen
pad4 = torch.zeros(((n_bytes + 15) // 16) * 16 + 4096, dtype=torch.uint8, device=dev)
pad4[:n_bytes] = out[:n_bytes]
cut = [n_bytes * g // world // 16 * 16 for g in range(world)] + [n_bytes]
prev_exit = None
letter_off = 0
for g in range(world):
    b0 = max(cut[g] - 4096, 0); b1 = min(cut[g + 1] + 4096, n_bytes)
    buf = pad4[b0: ((b1 + 15) // 16) * 16].contiguous()
    bit0 = b0 * 8
    avail = min(buf.numel() * 8, total_bits - bit0)
    own_b = cut[g] * 8 - bit0; own_e = min(cut[g + 1] * 8, total_bits) - bit0
    e, x, cnt = eng.decode_count(buf, avail, own_b, own_e, bit0, tree, entry_bit=0 if g == 0 else -1)
    o = torch.zeros(cnt + 64, dtype=torch.uint8, device=dev)
    eng.decode_write(o); eng.sync()
    exp = d[letter_off: letter_off + cnt]
    bad = torch.nonzero(o[:cnt] != exp).flatten()
    print(f"g={g} bit0={bit0} own=[{own_b},{own_e}) entry={e} (abs {e+bit0}) prev_exit={prev_exit} exit={x+bit0} cnt={cnt} letter_off={letter_off} mism={bad.numel()}",
          (int(bad[0]), int(bad[-1])) if bad.numel() else "")
    prev_exit = x + bit0
    letter_off += cnt
print("total letters", letter_off, n_total)
