#!/usr/bin/env python3
"""Top CUDA source lines of one kernel from an ncu report (needs -lineinfo + --import-source on):
warp-stall samples, executed warp instructions and average active lanes per line.
    tools/ncu_hot_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    import re
    hdr, lines, fn, path, want = None, {}, "", "", None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r and r[0] == "Function Name":
            fn = r[1]
            if want is None and re.search(kern, fn):
                want = fn
                print("kernel:", fn[:110])
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or fn != want or len(r) < len(hdr) or not r[0].strip().isdigit():
            continue
        def num(col):
            v = r[hdr.index(col)].replace(",", "")
            return int(v) if v.isdigit() else 0
        key = (path + ":" + r[0], r[1].strip()[:130])
        s, wi, ti = num("# Samples"), num("Instructions Executed"), num("Thread Instructions Executed")
        a = lines.setdefault(key, [0, 0, 0]); a[0] += s; a[1] += wi; a[2] += ti
    tot_s = sum(a[0] for a in lines.values()) or 1
    tot_w = sum(a[1] for a in lines.values()) or 1
    print("total samples %d, warp instructions %d, average lanes %.1f" % (tot_s, tot_w, sum(a[2] for a in lines.values()) / tot_w))
    for (ln, src), a in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp %5.1f%% ins lanes %4.1f | %s: %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_w, a[2] / max(1, a[1]), ln, src))


if __name__ == "__main__":
    main()
