"""Summarise an .ncu-rep (raw page) into the handful of numbers the design is judged on.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed_op_shared_atom.sum', 'lts__t_bytes.sum', 'sm__cycles_active.avg']
stall = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
out = [f"# {rep}"]
for r in data:
    out.append("## " + r[idx['Kernel Name']].split('(')[0])
    for w in want:
        if w in idx:
            out.append(f"{w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    vals = sorted([(float(r[idx[h]] or 0), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''))
                   for h in stall], reverse=True)
    out.append("top stalls (warps per issue): " + '  '.join(f"{n}={v:.2f}" for v, n in vals[:7]))
    out.append("")
text = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
