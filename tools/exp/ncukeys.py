#!/usr/bin/env python
"""the handful of ncu metrics the design log quotes, out of a .ncu-rep: ncukeys.py x.ncu-rep"""
import csv, subprocess, sys, io
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_active.min", "sm__cycles_active.avg", "sm__cycles_active.max",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.avg.per_cycle_active"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:60])
    for i, h in enumerate(hdr):
        if h in keys or ("stalled" in h and "per_issue_active" in h and float(r[i] or 0) > 0.05):
            print("  %-85s %-8s %s" % (h, units[i], r[i]))
