"""Per-CUDA-source-line aggregation of an ncu report's source page (instructions executed, stall samples):
usage: python tools/ncu_lines.py <report.ncu-rep> [n_lines]   (report captured with --import-source on, code built with -lineinfo)"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None
agg = defaultdict(lambda: [0.0, 0.0, ''])
tot_i = tot_s = 0.0
for r in rows:
    if len(r) == 2 and r[0] == 'File Name':
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr): continue
    try:
        ins = float(r[ix['Instructions Executed']] or 0); smp = float(r[ix['# Samples']] or 0)
    except ValueError:
        continue
    key = (fname, r[0])
    agg[key][0] += ins; agg[key][1] += smp
    if not agg[key][2]: agg[key][2] = r[1]
    tot_i += ins; tot_s += smp
print('total warp instructions', tot_i, 'samples', tot_s)
for (f, ln), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{i / tot_i * 100:5.1f}% ins {s / max(tot_s, 1) * 100:5.1f}% smp  {f}:{ln}  {src.strip()[:100]}")
