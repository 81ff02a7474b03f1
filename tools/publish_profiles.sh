#!/bin/bash
# Turns what tools/collect_profiles.sh left in gpurun_out/ into the committed files under profiles/
# (JSON lines copied, launch list as a table, ncu reports as text summaries + hottest source lines).
#   tools/publish_profiles.sh r01_s3
set -e
TAG=${1:?tag}
G=gpurun_out
P=profiles
grep '^"' $G/${TAG}_launches.csv > $P/${TAG}_launches.csv
for f in bench.json bench_reference.json configs.jsonl traversal_c4.jsonl; do cp $G/${TAG}_$f $P/${TAG}_$f; done
python - "$TAG" <<'PY'
import json, subprocess, sys
tag = sys.argv[1]
j = json.load(open('gpurun_out/%s_bench.json' % tag))
sm = j['stage_ms']; tot = sm['total']
shares = ", ".join("%s %.1f%%" % (k, 100 * sm[k] / tot) for k in ('trace_closest', 'trace_shadow', 'shade', 'raygen_film'))
table = subprocess.run(['python', 'tools/launch_table.py', 'gpurun_out/%s_launches.csv' % tag], capture_output=True, text=True).stdout
open('profiles/%s_launches.md' % tag, 'w').write('''# %s -- ncu launch list, Cornell box

Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --spp 32 --steps 1 --warmup 1 --no-cpu-baseline`
(every launch of the warm-up, the timed and the profiled render, one 32-spp pass each -- the pass size the 1024-spp bench uses; per-launch times under ncu are cold-cache and serialised, so read the SHARES). Table made by `tools/launch_table.py`.

%s
Stage shares from CUDA events inside the un-profiled 1024-spp bench, serial schedule (`profiles/%s_bench.json`, `stage_ms`): %s (shade = classify + all shading kernels) -- in agreement with the shares above. The timed render itself overlaps independent stages on three streams and takes %.0f ms per step instead of the %.0f ms of the serial schedule.
''' % (tag, table, tag, shares, j['ms_per_step'], tot))
# DRAM traffic of k_trace_closest per ms of kernel time, from the full ncu capture (bench.py scales it to its own launch time)
import csv
out = subprocess.run(['ncu', '-i', 'gpurun_out/%s_cornell.ncu-rep' % tag, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
def val(r, k):
    v = float(r[hdr.index(k)].replace(',', '')); u = units[hdr.index(k)]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u, 1.0)
ls = []
for r in rows[2:]:
    if 'k_trace_closest' in r[hdr.index('Kernel Name')]:
        ls.append(dict(gpu_time_ms=val(r, 'gpu__time_duration.sum'), dram_bytes_read=val(r, 'dram__bytes_read.sum'), dram_bytes_write=val(r, 'dram__bytes_write.sum')))
per_ms = sum(l['dram_bytes_read'] + l['dram_bytes_write'] for l in ls) / sum(l['gpu_time_ms'] for l in ls)
json.dump({"source": "profiles/%s_ncu_cornell.txt (ncu --set full, bench.py --spp 32: one 32-spp pass, the pass shape of the 1024-spp bench)" % tag,
           "k_trace_closest": {"launches": ls, "dram_bytes_per_ms": per_ms,
                               "note": "almost all of it is path-state records (ray / MIS records read, hit and shading-kind bytes written); the 33-primitive BVH itself stays in L1"}},
          open('profiles/%s_traffic.json' % tag, 'w'), indent=1)
PY
(echo "# ncu --set full --clock-control none --import-source on, Cornell box, one 32-spp pass (bench.py --spp 32), launches 14-20 of the trace / classify / shade kernels (second bounce of the timed pass)"
 python tools/ncu_summary.py $G/${TAG}_cornell.ncu-rep
 echo; echo "# hot CUDA source lines, k_trace_closest<flat search>"; python tools/ncu_hot_lines.py $G/${TAG}_cornell.ncu-rep "k_trace_closest" 16
 echo; echo "# hot CUDA source lines, k_shade<matte>"; python tools/ncu_hot_lines.py $G/${TAG}_cornell.ncu-rep "k_shade<\(int\)1" 16) > $P/${TAG}_ncu_cornell.txt 2>&1
(echo "# ncu --set full --clock-control none --import-source on, C4 (1,000,008 triangles), tools/traversal_roofline.py --rays 22 --no-oracle --repeat 1"
 echo "# k_batch_closest<2> = vote-scheduled two-phase traversal on the random-ray batch (4,194,304 rays); <1> = reference-order counting traversal of the same batch; k_batch_any<2> = shadow-ray batch"
 python tools/ncu_summary.py $G/${TAG}_c4.ncu-rep
 echo; echo "# hot CUDA source lines, k_batch_closest<2>"; python tools/ncu_hot_lines.py $G/${TAG}_c4.ncu-rep "k_batch_closest<\(int\)2>" 16) > $P/${TAG}_ncu_c4.txt 2>&1
ls -la $P/${TAG}_*
