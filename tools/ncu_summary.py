#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the text kept under profiles/:
    tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt
Key launch / throughput / memory / stall metrics of every captured launch."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel: %s" % d.get("Kernel Name", ("", "?"))[1])
    for k in KEYS:
        if k in d:
            print("  %-62s %-16s %s" % (k, d[k][0], d[k][1]))
    print("  warp stall cycles per issued instruction (> 0.1):")
    for h in hdr:
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                v = float(d[h][1])
            except ValueError:
                continue
            if v > 0.1:
                print("    %-28s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
