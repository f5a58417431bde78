#!/usr/bin/env python
"""Print the headline metrics of one ncu report: ncu_summary.py <file.ncu-rep> [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active', 'launch__grid_size',
        'launch__block_size', 'sm__cycles_active.avg', 'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
print(f"# {title}")
print("# kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
for i, h in enumerate(hdr):
    if h in keep:
        print(f"{h},{units[i]},{vals[i]}")
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') and float(vals[i] or 0) > 3:
        print(f"{h},{units[i]},{vals[i]}")
