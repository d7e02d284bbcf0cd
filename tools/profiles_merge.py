#!/usr/bin/env python3
"""Merge freshly summarised ncu records (tools/ncu_summarize.py <tag> /tmp/<tag>.json) into profiles/r02_ncu_full_summary.json with a capture_set
and a note, and print the markdown rows for profiles/README.md.  usage: profiles_merge.py <tag> "<note>" [<tag> "<note>" ...]"""
import json, subprocess, sys
dst = "profiles/r02_ncu_full_summary.json"
allr = json.load(open(dst))
args = sys.argv[1:]
rows = []
for tag, note in zip(args[0::2], args[1::2]):
    tmp = "/tmp/ncu_%s.json" % tag
    subprocess.run([sys.executable, "tools/ncu_summarize.py", tag, tmp], check=True, stdout=subprocess.DEVNULL)
    new = json.load(open(tmp))
    allr = [r for r in allr if r.get("capture_set") != tag]
    for r in new:
        r["capture_set"] = tag; r["note"] = note
        rows.append("| %s | %s | `%s` | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |" % (
            tag, r["capture"], r["kernel"].replace("void ", "").replace("<unnamed>::", ""), r.get("time_us", 0), r.get("dram_read_MB", 0), r.get("dram_write_MB", 0),
            r.get("dram_pct", 0), r.get("tensor_pipe_pct", 0), r.get("warps_active_pct", 0), note))
    allr = new + allr
json.dump(allr, open(dst, "w"), indent=1)
print("\n".join(rows))
