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


def build(force=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                           "-Wall", "-Wextra"] + SRC + ["-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
