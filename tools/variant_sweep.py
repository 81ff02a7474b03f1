#!/usr/bin/env python3
"""Compile-time tuning variants of libcpbrt.so, built here (nvcc cross-compiles) and timed on the GPU box.

    tools/variant_sweep.py build            # here: one libcpbrt_<name>.so per variant under curry-pbrt_b200/build/variants/
    tools/variant_sweep.py run [--spp 256]  # on the GPU box: bench.py (and config_bench.py C4 with --c4) per variant -> gpurun_out/variants.jsonl

A variant is a set of -D macros for cpbrt_api.cu (the only translation unit with device code of the
render path); every other object is shared with the normal build.  The library is chosen with
CPBRT_LIB (curry-pbrt_b200/__init__.py)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "curry-pbrt_b200")
VAR = os.path.join(PKG, "build", "variants")

VARIANTS = {
    "base": [],
    "checked": ["-DCPBRT_CHECKED"],
    "flat_unroll_2": ["-DCPBRT_FLAT_UNROLL=2"],
    "flat_unroll_4": ["-DCPBRT_FLAT_UNROLL=4"],
    "flat_unroll_8": ["-DCPBRT_FLAT_UNROLL=8"],
    "shade_64_threads": ["-DCPBRT_SHADE_THREADS=64"],
    "shade_32_threads": ["-DCPBRT_SHADE_THREADS=32"],
    "shade_256_threads": ["-DCPBRT_SHADE_THREADS=256"],      # device-side bounds checks (dmath.cuh CP_CHECK); run tools/sanitize_workload.py with CPBRT_LIB set to it
    # examples of what has been swept (profiles/r01_s3_variant_sweeps.jsonl):
    "closest_loop_8_blocks": ["-DCPBRT_BLOCKS_CLOSEST_LOOP=8"],
    "loop_refill_12": ["-DCPBRT_REFILL_CLOSEST_LOOP=12", "-DCPBRT_REFILL_SHADOW_LOOP=12"],
    "vote_refill_8": ["-DCPBRT_REFILL_CLOSEST_VOTE=8", "-DCPBRT_REFILL_SHADOW_VOTE=8"],
    "halton_one_by_one": ["-DCPBRT_HALTON_PAIR=0"],
    "matte_4_blocks": ["-DCPBRT_BLOCKS_MATTE=4"],
    "vote_c5_s6": ["-DCPBRT_BLOCKS_CLOSEST_VOTE=5", "-DCPBRT_BLOCKS_SHADOW_VOTE=6"],
    "vote_c5_s8": ["-DCPBRT_BLOCKS_CLOSEST_VOTE=5"],
    "vote_c6_s6": ["-DCPBRT_BLOCKS_SHADOW_VOTE=6"],
    "vote_c4_s6": ["-DCPBRT_BLOCKS_CLOSEST_VOTE=4", "-DCPBRT_BLOCKS_SHADOW_VOTE=6"],
    "vote_c7_s8": ["-DCPBRT_BLOCKS_CLOSEST_VOTE=7"],
    "sp7": ["-DCPBRT_BLOCKS_SPECULAR=7"],
    "sp6": ["-DCPBRT_BLOCKS_SPECULAR=6"],
    "c5_s6_m6": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=6", "-DCPBRT_BLOCKS_MATTE=6"],
    "c5_s6_m7": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=6", "-DCPBRT_BLOCKS_MATTE=7"],
    "c5_s5_m7": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=5", "-DCPBRT_BLOCKS_MATTE=7"],
    "c5_s4_m6": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=4", "-DCPBRT_BLOCKS_MATTE=6"],
    "c5_s5_m6_sp5": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=5", "-DCPBRT_BLOCKS_MATTE=6", "-DCPBRT_BLOCKS_SPECULAR=5"],
    "c5_s7_m6": ["-DCPBRT_BLOCKS_FLAT_CLOSEST=5", "-DCPBRT_BLOCKS_FLAT_SHADOW=7", "-DCPBRT_BLOCKS_MATTE=6"],
    "flat_5_blocks": ["-DCPBRT_BLOCKS_FLAT=5"],
    "flat_4_blocks": ["-DCPBRT_BLOCKS_FLAT=4"],
    "flat_6_matte_6": ["-DCPBRT_BLOCKS_FLAT=6", "-DCPBRT_BLOCKS_MATTE=6"],
    "flat_5_matte_6": ["-DCPBRT_BLOCKS_FLAT=5", "-DCPBRT_BLOCKS_MATTE=6"],
    "flat_7_blocks": ["-DCPBRT_BLOCKS_FLAT=7"],
    "flat_6_blocks": ["-DCPBRT_BLOCKS_FLAT=6"],
    "matte_6_blocks": ["-DCPBRT_BLOCKS_MATTE=6"],
    "specular_6_blocks": ["-DCPBRT_BLOCKS_SPECULAR=6"],
    "fused_matte_4": ["-DCPBRT_BLOCKS_MATTE_FUSED=4"],
    "fused_matte_3": ["-DCPBRT_BLOCKS_MATTE_FUSED=3"],
    "fused_matte_6": ["-DCPBRT_BLOCKS_MATTE_FUSED=6"],
}


def build(names):
    sys.path.insert(0, PKG)
    import build as B
    B.build()
    os.makedirs(VAR, exist_ok=True)
    others = [os.path.join(B.OBJ, f) for f in os.listdir(B.OBJ) if f.endswith(".o") and f not in ("cpbrt_api.cu.o", "cpbrt_api_checked.o")]
    procs = []
    for name in names:
        obj = os.path.join(VAR, "cpbrt_api_%s.o" % name)
        cmd = [B.NVCC] + B.NVCC_FLAGS + VARIANTS[name] + ["-c", os.path.join(B.CSRC, "cpbrt_api.cu"), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, obj, p in procs:
        out = p.communicate()[0]
        open(os.path.join(VAR, "ptxas_%s.log" % name), "w").write(out)
        if p.returncode:
            sys.stderr.write(out)
            raise SystemExit("variant %s failed to compile" % name)
        so = os.path.join(VAR, "libcpbrt_%s.so" % name)
        subprocess.check_call([B.NVCC, "-shared", "-o", so, obj] + others + ["-lz", "-ldl", "-lpthread", "-cudart", "static"])
        print("built", so)


def run(names, spp, c4):
    out_path = os.path.join(ROOT, "gpurun_out", "variants.jsonl")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "a") as out:
        for name in names:
            so = os.path.join(VAR, "libcpbrt_%s.so" % name)
            if not os.path.exists(so):
                continue
            env = dict(os.environ, CPBRT_LIB=so)
            rec = dict(variant=name, flags=VARIANTS[name])
            if not name.startswith("vote"):      # vote* variants only change the big-tree schedule, loop* the small-tree one
                r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--spp", str(spp), "--steps", "2", "--warmup", "2", "--no-cpu-baseline"],
                                   env=env, capture_output=True, text=True, cwd=ROOT)
                try:
                    j = json.loads(r.stdout.strip().splitlines()[-1])
                    rec.update(msamples_s=j["value"], stage_ms=j.get("stage_ms"))
                except Exception:
                    rec.update(error=(r.stderr or r.stdout)[-400:])
            if c4 and not name.startswith("loop"):
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "config_bench.py"), "--configs", "C4"],
                                   env=env, capture_output=True, text=True, cwd=ROOT)
                try:
                    j = json.loads(r.stdout.strip().splitlines()[-1])
                    rec.update(c4_msamples_s=j["msamples_s"], c4_stage_ms=j.get("stage_ms"), c4_checks=j.get("checks"))
                except Exception:
                    rec.update(c4_error=(r.stderr or r.stdout)[-400:])
            out.write(json.dumps(rec) + "\n")
            out.flush()
            print(json.dumps(rec))


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "build"
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    spp = int(sys.argv[sys.argv.index("--spp") + 1]) if "--spp" in sys.argv else 256
    if mode == "build":
        build(names)
    else:
        run(names, spp, "--c4" in sys.argv)
