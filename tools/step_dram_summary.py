#!/usr/bin/env python3
"""gpurun_out/step_dram_<tag>.csv (tools/ncu_step_dram.sh) -> profiles/r02_step_dram.json: DRAM bytes read + written by every kernel of ONE
eager training step (forward + backward + Adam; the third step of the capture), per kernel and in total."""
import csv
import json
import sys
from collections import OrderedDict

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_dram_r2.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_step_dram.json"
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
kn, mn, mv, idc = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
launch = OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) != len(H):
        continue
    e = launch.setdefault(int(r[idc]), {"kernel": r[kn].split("(")[0].split("::")[-1].strip()[:60]})
    unit = r[H.index("Metric Unit")]
    v = float(r[mv].replace(",", ""))
    if "bytes" in r[mn]:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    e[r[mn]] = v
L = list(launch.values())
starts = [i for i, e in enumerate(L) if "stn_trunk_fwd" in e["kernel"]]
ends = [i for i, e in enumerate(L) if "adam_kernel" in e["kernel"]]
assert len(starts) >= 3 and len(ends) >= 3, (len(starts), len(ends))
step = L[starts[2]:ends[2] + 1]
# the weight-image preparation kernels of the step are issued before stn_trunk_fwd on the side branch (eager mode: just before it)
pre = []
i = starts[2] - 1
while i > ends[1] and "prep_weight_images" in L[i]["kernel"] or (i > ends[1] and "Memset" in L[i]["kernel"]):
    pre.append(L[i]); i -= 1
step = pre[::-1] + step
per = OrderedDict()
for e in step:
    k = per.setdefault(e["kernel"], {"launches": 0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "time_us": 0.0})
    k["launches"] += 1
    k["dram_read_MB"] += e.get("dram__bytes_read.sum", 0) / 1e6
    k["dram_write_MB"] += e.get("dram__bytes_write.sum", 0) / 1e6
    k["time_us"] += e.get("gpu__time_duration.sum", 0) / 1e3
tot_r = sum(k["dram_read_MB"] for k in per.values()); tot_w = sum(k["dram_write_MB"] for k in per.values())
out = {"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one eager step (CRNN_GRAPH=0 CRNN_OVERLAP=0) of tools/prof_step.py: B=64, 128x32, GRU, dropout on",
       "launches": len(step), "dram_read_bytes": tot_r * 1e6, "dram_write_bytes": tot_w * 1e6, "dram_bytes_per_step": (tot_r + tot_w) * 1e6,
       "serialised_time_us": sum(k["time_us"] for k in per.values()),
       "kernels": [dict(kernel=n, **{a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items()}) for n, v in sorted(per.items(), key=lambda kv: -(kv[1]["dram_read_MB"] + kv[1]["dram_write_MB"]))]}
json.dump(out, open(dst, "w"), indent=1)
print("step: %d launches, read %.1f MB + write %.1f MB = %.2f GB" % (len(step), tot_r, tot_w, (tot_r + tot_w) / 1e3))
