"""Builds libcpbrt.so (host parser + sm_100a CUDA kernels) in-tree with nvcc.

    python curry-pbrt_b200/build.py            # incremental
    python curry-pbrt_b200/build.py --force

-fmad=false / -ffp-contract=off are part of the contract, not tuning: BVH slots and Halton
indices must be bit-identical to the reference's unfused f32 arithmetic (see csrc/device/dmath.cuh).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcpbrt.so")
OUT_CHECKED = os.path.join(HERE, "libcpbrt_checked.so")      # same library with device-side bounds checks (tests/test_gpu_parity.py::test_checked_build...)
OBJ = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wextra", "-pthread"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-Xptxas", "-v"]

HOST_SOURCES = ["host/parser.cpp", "host/scene_build.cpp", "host/halton_tables.cpp", "host/host_api.cpp",
                "host/image_io.cpp", "bvh_build.cpp"]
CUDA_SOURCES = ["cpbrt_api.cu", "bvh_build_gpu.cu"]


def _deps():
    deps = [os.path.join(HERE, "..", "include", "cpbrt.h")]
    for root, _, files in os.walk(CSRC):
        deps += [os.path.join(root, f) for f in files]
    return deps


# which headers each CUDA translation unit depends on (besides itself and include/cpbrt.h)
CUDA_DEPS = {
    "cpbrt_api.cu": ["device", "bvh_build.h"],
    "bvh_build_gpu.cu": ["bvh_build.h"],
}


def _unit_deps(src):
    deps = [os.path.join(CSRC, src), os.path.join(HERE, "..", "include", "cpbrt.h"), os.path.abspath(__file__)]
    for d in CUDA_DEPS.get(src, []):
        path = os.path.join(CSRC, d)
        if os.path.isdir(path):
            deps += [os.path.join(path, f) for f in os.listdir(path)]
        else:
            deps.append(path)
    return deps


def _compile_cuda(src, obj, verbose, extra=()):
    r = subprocess.run([NVCC] + NVCC_FLAGS + list(extra) + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
    with open(os.path.join(OBJ, os.path.basename(obj)[:-2] + ".ptxas.log"), "w") as f:
        f.write(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        raise RuntimeError("nvcc failed on " + src)
    if verbose:
        sys.stderr.write(r.stderr)


def _fingerprint(paths, extra=""):
    """Content hash of the sources: staleness must not depend on file times (a snapshot copied to
    another machine keeps contents, not necessarily mtimes)."""
    import hashlib
    h = hashlib.sha256(extra.encode())
    for p in sorted(os.path.abspath(x) for x in paths):
        if os.path.isfile(p) and not p.endswith((".pyc",)):
            h.update(os.path.relpath(p, HERE).encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return ""


def build(force=False, verbose=False):
    """Idempotent and safe to call from several processes at once (torchrun ranks): one holds the
    lock and builds, the others wait and find the library up to date."""
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    flags = " ".join(NVCC_FLAGS + HOST_CXX)
    want = _fingerprint(_deps(), flags)
    stamp = os.path.join(OBJ, "libcpbrt.stamp")
    if not force and os.path.exists(OUT) and os.path.exists(OUT_CHECKED) and _read(stamp) == want:
        return OUT
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(OUT) and os.path.exists(OUT_CHECKED) and _read(stamp) == want:
            return OUT
        objs = []
        # CUDA translation units in parallel threads (nvcc is a subprocess), each only when its inputs changed
        from concurrent.futures import ThreadPoolExecutor
        jobs = []
        with ThreadPoolExecutor(max_workers=len(CUDA_SOURCES)) as pool:
            for src in CUDA_SOURCES:
                o = os.path.join(OBJ, src.replace("/", "_") + ".o")
                unit = _fingerprint(_unit_deps(src), flags)
                if force or not os.path.exists(o) or _read(o + ".stamp") != unit:
                    jobs.append((pool.submit(_compile_cuda, src, o, verbose), o, unit))
                objs.append(o)
            # the bounds-checked twin of the render translation unit (dmath.cuh CP_CHECK) -> libcpbrt_checked.so
            co = os.path.join(OBJ, "cpbrt_api_checked.o")
            cunit = _fingerprint(_unit_deps("cpbrt_api.cu"), flags + " checked")
            if force or not os.path.exists(co) or _read(co + ".stamp") != cunit:
                jobs.append((pool.submit(_compile_cuda, "cpbrt_api.cu", co, verbose, ["-DCPBRT_CHECKED"]), co, cunit))
            for src in HOST_SOURCES:
                o = os.path.join(OBJ, src.replace("/", "_") + ".o")
                subprocess.check_call(HOST_CXX + ["-c", os.path.join(CSRC, src), "-o", o])
                objs.append(o)
            for j, o, unit in jobs:
                j.result()
                with open(o + ".stamp", "w") as f:
                    f.write(unit)
        tmp = OUT + ".tmp.%d" % os.getpid()
        subprocess.check_call([NVCC, "-shared", "-o", tmp] + objs + ["-lz", "-ldl", "-lpthread", "-cudart", "static"])
        os.replace(tmp, OUT)
        cobjs = [co if o.endswith("cpbrt_api.cu.o") else o for o in objs]
        tmp = OUT_CHECKED + ".tmp.%d" % os.getpid()
        subprocess.check_call([NVCC, "-shared", "-o", tmp] + cobjs + ["-lz", "-ldl", "-lpthread", "-cudart", "static"])
        os.replace(tmp, OUT_CHECKED)
        with open(stamp, "w") as f:
            f.write(want)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
