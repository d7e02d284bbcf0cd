#!/usr/bin/env python3
"""Source-level hot spots of one .ncu-rep: stall-reason totals, opcode histogram, top sampled SASS lines with their CUDA source line."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
def I(v):
    try: return int(v)
    except Exception: return 0
tot_s = sum(I(r[iS]) for r in data); tot_i = sum(I(r[iI]) for r in data)
print('samples', tot_s, 'warp-instructions', tot_i)
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tt = {n: sum(I(r[hdr.index(n)]) for r in data) for n in names}
print(' '.join('%s=%.3f' % (n[6:], v / max(1, sum(tt.values()))) for n, v in sorted(tt.items(), key=lambda kv: -kv[1])[:8]))
ci = Counter(); cs = Counter()
for r in data:
    f = r[iSrc].split()
    if not f: continue
    op = f[1] if f[0].startswith('@') and len(f) > 1 else f[0]
    ci[op] += I(r[iI]); cs[op] += I(r[iS])
print('opcode            inst   samples')
for op, n in cs.most_common(14): print('%-16s %6.3f %6.3f' % (op, ci[op] / max(1, tot_i), n / max(1, tot_s)))
print('--- top lines')
for k, r in sorted(enumerate(data), key=lambda kr: -I(kr[1][iS]))[:top]:
    st = sorted(((I(r[hdr.index(n)]), n[6:]) for n in names), reverse=True)[:2]
    print('%5d %-58s %6s %8s  %s' % (k, r[iSrc][:58], r[iS], r[iI], ' '.join('%s:%d' % (n, v) for v, n in st)))
