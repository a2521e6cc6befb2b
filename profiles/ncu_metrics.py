#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep raw-page csv (stdin): ncu -i x.ncu-rep --page raw --csv | ncu_metrics.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
H = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__registers_per_thread',
        'sm__inst_executed.sum', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
for r in rows[2:]:
    for w in want:
        for i, h in enumerate(H):
            if h == w:
                print(f'{w:80s} {rows[1][i]:12s} {r[i][:60]}')
    print()
