#!/usr/bin/env python3
"""Print the handful of ncu metrics the profiling recipe asks for, one kernel per block (reads a .ncu-rep here, no GPU)."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed_op_global_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum"]
STALL = "smsp__average_warps_issue_stalled_"
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:100])
    for k in KEYS:
        if k in d:
            print("  %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = sorted(((float(v.replace(",", "")), k) for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
    for v, k in st[:8]:
        print("  stall %-60s %.2f" % (k[len(STALL):-len("_per_issue_active.ratio")], v))
