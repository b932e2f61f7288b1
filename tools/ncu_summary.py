#!/usr/bin/env python
"""Key metrics + stall reasons per captured launch of an ncu report.  python tools/ncu_summary.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if name in seen and '--all' not in sys.argv:
        continue
    seen.add(name)
    print('----')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:75s} {r[i][:100]} {units[i]}")
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.25:
                print(f"     stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.2f}")
