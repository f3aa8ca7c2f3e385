#!/usr/bin/env python
"""Stall breakdown and pipe utilisation of the last kernel in an `ncu --page raw --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, r = rows[0], rows[-1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__cycles_elapsed.avg']
for i, h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h and float(r[i] or 0) > 0.03):
        print(f"  {h:88s} {r[i]}")
