#!/usr/bin/env python3
"""Imports the reference's benchmark input scenes as DATA fixtures (dev container only).

BASELINE.json names scenes/cornel_box.pbrt and scenes/sphere.pbrt as benchmark inputs, and the
GPU box has no /root/reference, so the scene descriptions must live in this repo.  Only the
scene content is kept: comments (including the author's commented-out experiments) are dropped
and whitespace is normalised; the token stream -- and therefore the parsed scene -- is unchanged
(tests/test_host_parser.py checks the primitive/light/material tables).  box-texture.png is the
texture config C3 names; it is copied byte for byte.
"""
import pathlib, re, shutil, sys

ref = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/scenes")
out = pathlib.Path(__file__).resolve().parents[1] / "scenes"
out.mkdir(exist_ok=True)
for name in ("cornel_box.pbrt", "sphere.pbrt"):
    lines = []
    for ln in (ref / name).read_text().splitlines():
        ln = ln.split("#", 1)[0].rstrip()
        if ln.strip():
            lines.append(re.sub(r"[ \t]+", " ", ln.strip()))
    header = "# %s -- input fixture for the curry-pbrt hot path (scene content of the reference's scenes/%s)\n" % (name, name)
    (out / name).write_text(header + "\n".join(lines) + "\n")
shutil.copyfile(ref / "box-texture.png", out / "box-texture.png")
print("wrote", sorted(p.name for p in out.iterdir()))
