"""Builds oracle/liboracle.so (the CPU restatement; test infrastructure only).

The reference itself (Rust, 18 un-vendored crates) cannot be compiled in this image, so there
is no oracle/_ref: the restatement is the only CPU implementation available.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "liboracle.so")
SRC = [os.path.join(HERE, "oracle.cpp")]
DEPS = SRC + [os.path.join(HERE, "oracle_math.h"), os.path.join(HERE, "..", "include", "cpbrt.h")]


def _fingerprint():
    import hashlib
    h = hashlib.sha256()
    for p in DEPS:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False):
    """Content-stamped and locked: safe to call from several processes at once, independent of file times."""
    import fcntl
    stamp = OUT + ".stamp"
    want = _fingerprint()

    def fresh():
        try:
            return os.path.exists(OUT) and open(stamp).read().strip() == want
        except OSError:
            return False
    if not force and fresh():
        return OUT
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and fresh():
            return OUT
        tmp = OUT + ".tmp.%d" % os.getpid()
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                               "-Wall", "-Wextra"] + SRC + ["-o", tmp])
        os.replace(tmp, OUT)
        with open(stamp, "w") as f:
            f.write(want)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
