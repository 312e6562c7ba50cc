#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (run here, no GPU): one block per captured launch with the metrics the roofline
discussion uses.  usage: tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_x.txt"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]

for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print(f"== {rep}: no data"); continue
    hdr, units = rows[0], rows[1]
    print(f"== {rep}")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"-- {d.get('Kernel Name', '?')[:110]}  grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}")
        for k in KEYS:
            if k in d and k not in ("launch__grid_size", "launch__block_size"):
                print(f"   {k:88s} {d[k]:>16s} {units[hdr.index(k)]}")
