#!/usr/bin/env python3
"""Summarise gpurun_out/prof_<tag>_*.ncu-rep (ncu --set full captures) into profiles/<round>_ncu_full_summary.json."""
import csv, glob, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_name = sys.argv[2] if len(sys.argv) > 2 else "profiles/r01_ncu_full_summary.json"
want = {'gpu__time_duration.sum': 'time_us', 'dram__bytes_read.sum': 'dram_read_MB', 'dram__bytes_write.sum': 'dram_write_MB',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct', 'launch__registers_per_thread': 'regs',
        'launch__grid_size': 'grid', 'launch__block_size': 'block', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_scoreboard',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_scoreboard',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio': 'stall_membar',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
        'l1tex__t_sector_hit_rate.pct': 'l1_hit_pct', 'lts__t_sector_hit_rate.pct': 'l2_hit_pct'}
res = []
for f in sorted(glob.glob(f"gpurun_out/prof_{tag}_*.ncu-rep")):
    txt = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if not rows:
        continue
    hdr = rows[0]
    for r in rows[2:]:
        d = {'capture': os.path.basename(f)[len(f"prof_{tag}_"):-8], 'kernel': r[hdr.index('Kernel Name')].split('(')[0][-60:]}
        for k, v in want.items():
            if k in hdr:
                try:
                    d[v] = round(float(r[hdr.index(k)]), 3)
                except ValueError:
                    d[v] = r[hdr.index(k)]
        res.append(d)
json.dump(res, open(out_name, 'w'), indent=1)
for d in res:
    print(d['capture'], d.get('time_us'), 'us dram%', d.get('dram_pct'), 'tensor%', d.get('tensor_pipe_pct'), 'warps%', d.get('warps_active_pct'))
