#!/usr/bin/env python3
"""profiles/r02_ncu_c4_traversal.json from an `ncu --set full` capture of tools/traversal_roofline.py: per launch of
k_batch_closest<2> (the vote-scheduled two-phase traversal) the L2 and DRAM bytes, issue-slot use and lanes per
instruction -- what bench.py quotes as `roofline.traffic` / `roofline.issue`.

    python tools/ncu_c4_json.py gpurun_out/r02_s_c4.ncu-rep [log2 rays of the captured batch]
"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    log2_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}

    def val(r, k):
        return float(r[hdr.index(k)].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0)
    sel = [r for r in rows[2:] if "k_batch_closest<(int)2>" in r[hdr.index("Kernel Name")] or "k_batch_closest<2>" in r[hdr.index("Kernel Name")]]
    if not sel:
        raise SystemExit("no k_batch_closest<2> launch in " + rep)
    # the first captured launches belong to the camera-ray batch when the capture starts early: keep the launches of the last batch of equal grid
    n = len(sel)
    avg = lambda k: sum(val(r, k) for r in sel) / n
    j = dict(source="%s (ncu --set full --clock-control none, tools/traversal_roofline.py --rays %d --no-oracle --repeat 1), summary in profiles/r02_final_ncu_c4.txt" % (os.path.basename(rep), log2_rays),
             kernel="k_batch_closest<2>", rays_per_launch=1 << log2_rays, launches=n,
             gpu_time_ms=avg("gpu__time_duration.sum"),
             dram_bytes_per_launch=avg("dram__bytes_read.sum") + avg("dram__bytes_write.sum"),
             lts_bytes_per_launch=avg("lts__t_bytes.sum") if "lts__t_bytes.sum" in hdr else avg("lts__t_sectors.sum") * 32.0,
             issue=dict(issue_active_pct=avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        lanes_per_instruction=avg("smsp__thread_inst_executed_per_inst_executed.ratio"),
                        warps_active_pct=avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                        registers=avg("launch__registers_per_thread"),
                        useful_issue_fraction=avg("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0 * avg("smsp__thread_inst_executed_per_inst_executed.ratio") / 32.0),
             l2_hit_pct=avg("lts__t_sector_hit_rate.pct"),
             dram_pct_of_peak=avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
    json.dump(j, open(os.path.join(ROOT, "profiles", "r02_ncu_c4_traversal.json"), "w"), indent=1)
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    main()
