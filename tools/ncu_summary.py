#!/usr/bin/env python
"""Key metrics of one kernel from an `ncu --set full` report (ncu -i <rep> --page raw --csv), as a short text block.
  python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (smem)"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (regs)"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA (FP64 tensor) pipe % of peak, active"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 (DFMA) pipe % of peak, active"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier / issue"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait / issue"),
]
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
idx = {h: i for i, h in enumerate(hdr)}
for key, label in WANT:
    if key in idx:
        print(f"{label:48s} {vals[idx[key]]} {units[idx[key]]}")
