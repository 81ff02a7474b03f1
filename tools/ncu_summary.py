#!/usr/bin/env python3
"""Prints the handful of ncu metrics we track from a .ncu-rep (one block per captured launch).
Usage: tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("==== %s :: %s" % (rep, r[hdr.index("Kernel Name")][:90]))
            for k in KEYS:
                if k in hdr:
                    print("  %-72s %18s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
            stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]]
            for v, h in sorted(stalls, reverse=True)[:8]:
                print("  stall %-66s %18.3f warps/issue" % (h[len(STALL):-len("_per_issue_active.ratio")], v))


if __name__ == "__main__":
    main()
