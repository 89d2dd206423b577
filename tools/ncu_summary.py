#!/usr/bin/env python
"""Selected metrics of an `ncu --set full` report, one column per kernel launch:
    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x_raw.csv ; python tools/ncu_summary.py x_raw.csv "header text" > profiles/....csv
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size",
    "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
units = rows[rows.index(hdr) + 1]
data = [r for r in rows[rows.index(hdr) + 2:] if len(r) == len(hdr)]
ki = hdr.index("Kernel Name")
print("# " + (sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none") + "; value unit per cell")
print("metric," + ",".join('"' + r[ki][:70].replace('"', "'") + '"' for r in data))
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        print(m + "," + ",".join((r[i].replace(",", "") + " " + units[i]).strip() for r in data))
