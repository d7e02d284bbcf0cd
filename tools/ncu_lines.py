#!/usr/bin/env python3
"""Per CUDA source line totals of one .ncu-rep (warp instructions executed, stall samples), in file / line order: where the instructions of an
issue-bound kernel go.  usage: ncu_lines.py <rep> [min share in % to print, default 0.3]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, out = '?', None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Name': fname = r[1].split('/')[-1]; continue
    if '# Samples' in r: hdr = r; iS = r.index('# Samples'); iI = r.index('Instructions Executed'); continue
    if hdr is None or len(r) != len(hdr) or not r[0]: continue      # keep the per-CUDA-line aggregate rows only
    try: out.append((fname, int(r[0]), r[1], int(r[iS]), int(r[iI])))
    except ValueError: pass
ti = sum(o[4] for o in out); ts = sum(o[3] for o in out)
print('warp instructions %d, samples %d' % (ti, ts))
for f, ln, src, s, i in out:
    if 100.0 * i / max(1, ti) >= thr or 100.0 * s / max(1, ts) >= thr:
        print('%-16s %5d  inst %5.2f%%  samples %5.2f%%  %s' % (f, ln, 100.0 * i / ti, 100.0 * s / ts, src.strip()[:110]))
