#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per kernel: launches, total us, share)
for ONE training step (the launches between two consecutive adam_kernel launches) and for the whole capture.
usage: launch_summary.py gpurun_out/launches_X.csv profiles/r01_launches_bench_summary.md "<command line that was profiled>" """
import csv, re, sys
from collections import defaultdict
src, dst, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr = rows[0]; iK = hdr.index('Kernel Name'); iV = hdr.index('Metric Value'); iU = hdr.index('Metric Unit')
def us(r):
    v = float(r[iV].replace(',', '')); u = r[iU]
    return v / 1000 if u.startswith('n') else (v if u.startswith('u') else v * 1000)
def short(k): return re.sub(r'\(.*', '', k).replace('void ', '').replace('<unnamed>::', '').strip()
L = [(short(r[iK]), us(r)) for r in rows[1:]]
idx = [i for i, (k, u) in enumerate(L) if 'adam_kernel' in k]
def table(items):
    agg = defaultdict(lambda: [0, 0.0])
    for k, u in items: agg[k][0] += 1; agg[k][1] += u
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): out.append("| %s | %d | %.1f | %.1f%% |" % (k, v[0], v[1], 100 * v[1] / tot))
    return out, tot
with open(dst, 'w') as f:
    f.write("# ncu launch list of `%s` (first %d launches), round 2\n\n" % (cmd, len(L)))
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
    if len(idx) >= 2:
        step = L[idx[0] + 1: idx[1] + 1]
        t, tot = table(step)
        f.write("## One training step (%d launches, %.0f us serialised; the CUDA-graph replay of the same step overlaps the side branch and takes less)\n\n" % (len(step), tot))
        f.write("\n".join(t) + "\n\n")
    t, tot = table(L)
    f.write("## Whole capture (%d launches, %.0f us)\n\n" % (len(L), tot))
    f.write("\n".join(t) + "\n")
print("wrote", dst)
