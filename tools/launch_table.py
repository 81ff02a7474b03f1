#!/usr/bin/env python3
"""Markdown table of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
launch count, total time and share per kernel.    tools/launch_table.py launches.csv"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    k_name, k_val, k_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[k_metric] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[k_name], [0, 0.0])
        a[0] += 1
        a[1] += float(r[k_val].replace(",", "")) / 1e6
    total = sum(a[1] for a in agg.values()) or 1.0
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (name[:90], n, ms, 100.0 * ms / total))


if __name__ == "__main__":
    main()
