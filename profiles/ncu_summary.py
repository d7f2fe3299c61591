#!/usr/bin/env python
"""Print the judged metrics from `ncu -i X.ncu-rep --page raw --csv` output."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
for r in rows[2:]:
    print('\n'.join(f"  {hdr[i]:70s} {r[i][:60]} {units[i]}" for i in idx))
    print('-' * 100)
