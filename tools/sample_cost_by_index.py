#!/usr/bin/env python3
"""Does a camera sample get dearer with its index?  One GPU, the Cornell box of bench.py: 128-sample ranges at the
bottom, middle and top of [0, 1024), device-timed, film-point replay timed separately.  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np, torch
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    from conftest import SCENES, scene_text
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=900, yres=900, spp=1024), base_dir=SCENES)
    r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
    film = torch.zeros(900 * 900 * 3, dtype=torch.float32, device="cuda")
    out = {}
    for rep in range(3):
        for b in (0, 128, 256, 448, 640, 768, 896):
            _, st = r.render(sample_begin=b, sample_end=b + 128, device_film_ptr=film.data_ptr())
            if rep:
                out.setdefault(b, []).append(st.ms_total)
    row = {"range_of_128_samples_starting_at": {str(b): round(float(np.mean(v)), 3) for b, v in out.items()},
           "note": "ms_total includes the replay of the samples before the range (k_film_replay)"}
    print(json.dumps(row))


if __name__ == "__main__":
    main()
