#!/usr/bin/env python
"""Selected metrics of an .ncu-rep (read here, no GPU): python tools/ncu_metrics.py rep [substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
                        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "smsp__average_warp_latency_issue_stalled",
                        "smsp__average_warps_issue_stalled", "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_bank_conflicts", "smsp__inst_executed_op_shared", "launch__occupancy_limit"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, v in zip(hdr, vals):
        if any(w in h for w in want) and "pct_of_peak_sustained_elapsed" not in h.replace(want[0], "") or h in want:
            if ".min" in h or ".max" in h: continue
            print("  %-90s %s" % (h, v))
