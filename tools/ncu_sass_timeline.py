#!/usr/bin/env python3
"""Program-order SASS view of one .ncu-rep with stall samples, cut into segments at synchronisation instructions."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if '# Samples' in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def I(v):
    try: return int(v)
    except Exception: return 0
iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(I(r[iS]) for r in data)
print('total samples', tot)
seg_s = 0; seg_n = 0; seg_start = 0
SYNC = ('BAR.SYNC', 'SYNCS', 'UCGABAR', 'BRA', 'WARPSYNC', 'STAS', 'EXIT', 'MEMBAR', 'ERRBAR', 'LDGSTS', 'LDGDEPBAR', 'DEPBAR')
for k, r in enumerate(data):
    src = r[iSrc].strip()
    s = I(r[iS]); seg_s += s; seg_n += 1
    if any(t in src for t in SYNC) or s > tot * 0.01:
        st = sorted(((I(r[hdr.index(n)]), n[6:]) for n in names), reverse=True)[:2]
        print('%5d  seg[%4d instr %5d samp %4.1f%%]  %-60s samp=%-5d exec=%-8s %s' % (k, seg_n, seg_s, 100.0 * seg_s / tot, src[:60], s, r[iI], ' '.join('%s:%d' % (n, v) for v, n in st if v)))
        seg_s = 0; seg_n = 0
