"""Quick A/B timing of the device-resident phases (CUDA events on the ctx stream).
usage: [HUFFB200_SO=...] python tools/kern_bench.py [workload] [size_bytes] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import make_workload
from huff_encoding_b200.engine import Engine
from huff_encoding_b200.sharded import ShardedCodec

workload = sys.argv[1] if len(sys.argv) > 1 else "zipf"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 30
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
eng = Engine(0)
codec = ShardedCodec(eng, 1, 0, None)
ds = [make_workload(workload, size, 0, eng.device, seed_shift=k) for k in range(1 if workload == "fibonacci" else 2)]
n = max(x.numel() for x in ds)
comp = torch.empty(n + n // 4 + 4096, dtype=torch.uint8, device=eng.device)
out = torch.empty(n + 64, dtype=torch.uint8, device=eng.device)
torch.cuda.synchronize()          # the inputs are generated on torch's stream; the library's stream does not wait for it
with torch.cuda.stream(eng.stream):
    for i in range(2):
        codec.round_trip(ds[i % len(ds)], comp, out, want_events=True)
    eng.stream.synchronize()
    marks = [codec.round_trip(ds[i % len(ds)], comp, out, want_events=True) for i in range(iters)]
    eng.stream.synchronize()
last = ds[(iters - 1) % len(ds)]
assert torch.equal(out[:last.numel()], last)
ms = {k: sum(m[k][0].elapsed_time(m[k][1]) for m in marks) / iters for k in ("hist", "encode", "decode")}
print(os.environ.get("HUFFB200_SO", "default").split("/")[-1], workload, size,
      " ".join(f"{k}={v:.4f}" for k, v in ms.items()), "path", eng.ctx.last_decode_path())

import ctypes as C
cyc = (C.c_uint64 * 8)()
eng.lib.hb_ctx_fused_phase_cycles(eng.ctx.handle, cyc)
if cyc[6]:
    names = ("stage", "decode", "verify", "scan", "lookback", "compact")
    print("  fused phases, cycles per chunk:", " ".join(f"{nm}={cyc[i] / cyc[6]:.0f}" for i, nm in enumerate(names)), "chunks", cyc[6], "look-back alone", cyc[7] // cyc[6])
