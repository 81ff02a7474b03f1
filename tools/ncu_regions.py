#!/usr/bin/env python3
"""Warp instructions / stall samples of one kernel, grouped by source file and by ranges of lines.
    tools/ncu_regions.py source.csv kernel_regex [file:lo-hi=name ...]
source.csv = `ncu -i rep --page source --print-source cuda,sass --csv`"""
import csv, re, sys
src, kern = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    loc, name = a.split("=")
    f, rng = loc.split(":"); lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
hdr, fn, path, want = None, "", "", None
agg = {}
for r in csv.reader(open(src)):
    if r and r[0] == "File Path": path = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        fn = r[1]
        if want is None and re.search(kern, fn): want = fn; print("kernel:", fn[:100])
        continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or fn != want or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    def num(col):
        v = r[hdr.index(col)].replace(",", ""); return int(v) if v.isdigit() else 0
    ln = int(r[0]); key = path
    for f, lo, hi, name in regions:
        if f == path and lo <= ln <= hi: key = name; break
    a = agg.setdefault(key, [0, 0, 0]); a[0] += num("# Samples"); a[1] += num("Instructions Executed"); a[2] += num("Thread Instructions Executed")
ts = sum(a[0] for a in agg.values()) or 1; tw = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %d, lanes %.1f" % (tw, sum(a[2] for a in agg.values()) / tw))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5.1f%% ins %5.1f%% smp lanes %4.1f | %s" % (100.0 * a[1] / tw, 100.0 * a[0] / ts, a[2] / max(1, a[1]), k))
