#!/usr/bin/env python
"""Developer tool: the handful of ncu metrics this kernel is tuned by, from a .ncu-rep.  usage: python tools/ncu_brief.py X.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'sm__warps_active.avg.per_cycle_active',
        'sm__icc_request_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum']
stalls = {}
for i, x in enumerate(h):
    if x in want:
        print(f"{x:75s} {v[i]}")
    if x.startswith('smsp__pcsamp_warps_issue_stalled_') and not x.endswith('_not_issued'):
        try:
            stalls[x[len('smsp__pcsamp_warps_issue_stalled_'):]] = float(v[i])
        except ValueError:
            pass
tot = sum(stalls.values()) or 1
print("stalls:", ", ".join(f"{k} {100 * s / tot:.1f}%" for k, s in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]))
