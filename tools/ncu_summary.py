#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')], r[hdr.index('Block Size')])
    for w in want:
        if w in hdr:
            print(f"  {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'ratio' in h and 'not_issued' not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.15:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per')[0]} = {v:.2f}")
