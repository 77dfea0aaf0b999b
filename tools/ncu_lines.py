"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep (needs -lineinfo and
--import-source on): joins ncu's SASS page with nvdisasm's line table of the built library.
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep <kernel substring> [top N]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "huff_encoding_b200", "libhuffb200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin") and "hb_tree" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()

# line table of the kernel: list of (file, line) per instruction, in order
lines, cur, inside = [], ("?", 0), False
for ln in sass:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        inside = kname in m.group(1)
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln) and ".dword" not in ln and ".byte" not in ln:
        lines.append((cur, ln.split("*/", 1)[1].strip()))

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
if len(body) != len(lines):
    print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled (library rebuilt since the capture?)")
agg = {}
tot_inst = tot_samp = 0
for (loc, _txt), r in zip(lines, body):
    inst = int(float(r[ix["Instructions Executed"]] or 0))
    samp = int(float(r[ix["Warp Stall Sampling (All Samples)"]] or 0))
    wf = int(float(r[ix["L1 Wavefronts Shared"]] or 0)) if "L1 Wavefronts Shared" in ix else 0
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += inst
    a[1] += samp
    a[2] += wf
    tot_inst += inst
    tot_samp += samp
print(f"kernel {kname}: {tot_inst} warp instructions, {tot_samp} stall samples")
src_cache = {}
for loc, (inst, samp, wf) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    f, n = loc
    text = ""
    for d in ("huff_encoding_b200/csrc", "include"):
        pth = os.path.join(ROOT, d, f)
        if os.path.exists(pth):
            src_cache.setdefault(pth, open(pth).read().splitlines())
            if 0 < n <= len(src_cache[pth]):
                text = src_cache[pth][n - 1].strip()[:90]
    print(f"{f}:{n:<5d} inst {100 * inst / max(tot_inst, 1):5.1f}%  samples {100 * samp / max(tot_samp, 1):5.1f}%  smem wavefronts {wf:>10d}  | {text}")
