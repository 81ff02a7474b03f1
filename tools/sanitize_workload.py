#!/usr/bin/env python3
"""A small workload that touches every kernel of the render path, for compute-sanitizer:

    compute-sanitizer --tool memcheck|initcheck|synccheck|racecheck python tools/sanitize_workload.py

or, where compute-sanitizer is not available (it is closed on the GPU pool this was developed on), for the
library's own checked build (device-side bounds checks, include/cpbrt.h cpbrt_check_failures):

    python tools/variant_sweep.py build checked
    CPBRT_LIB=curry-pbrt_b200/build/variants/libcpbrt_checked.so python tools/sanitize_workload.py

Cornell box 48 x 48 (matte, mirror, glass; flat search, then the loop and vote tree walks forced with
CPBRT_SCHEDULE), one lane and two lanes (several passes: the three-stream bounce schedule and the
cross-lane event chain), a sample range that starts in the middle (film-point replay), the general
materials, sphere.pbrt (directlighting, spheres, point / distant lights), directlighting recursing
over delta lobes, a 60 x 50 height field (device SAH levels forced), the batch entry points.
One process, one GPU; it prints a checksum per step so that a run under the sanitizer can be
compared with a plain one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    from conftest import SCENES, lcg_rays, scene_text
    from test_oracle import _synthetic_mesh_scene
    from test_gpu_parity import GENERAL_MATERIALS, _cornell_with_white_material

    def run(tag, scene, **kw):
        r = cpbrt.Renderer(0); r.upload(scene); r.build_bvh()
        film, st = r.render(**kw)
        assert np.isfinite(film).all()
        checked, bad, site = r.check_failures()
        print("%-44s mean %.6f launches %d rays %d%s" % (tag, float(film.mean()), st.kernel_launches, st.rays_closest + st.rays_any,
                                                       "  bounds violations so far: %d (first site %d)" % (bad, site) if checked else ""), flush=True)
        return r

    cornell = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=48, yres=48, spp=6), base_dir=SCENES)
    for sched in ("flat", "fused", "loop", "vote"):
        os.environ["CPBRT_SCHEDULE"] = sched
        for lanes in ("1", "2"):
            os.environ["CPBRT_LANES"] = lanes
            r = run("cornell %s lanes=%s" % (sched, lanes), cornell, samples_per_pass=2)
        rays = np.concatenate([r.camera_rays(0), lcg_rays(4096, (-50, -50, -50), (610, 600, 610), seed=7)])
        h = r.trace_closest(rays); sh = rays.copy(); sh[:, 3] = np.float32(0.9); o = r.trace_any(sh)
        hc, _ = r.trace_closest(rays, count=True)
        assert np.array_equal(h["prim"], hc["prim"])
        print("%-44s hits %d occluded %d" % ("cornell %s batches" % sched, int((h["prim"] != 0xFFFFFFFF).sum()), int(o.sum())), flush=True)
        r.close()
    os.environ.pop("CPBRT_SCHEDULE"); os.environ.pop("CPBRT_LANES")
    r = run("cornell sample range [3, 6) (replay)", cornell, sample_begin=3, sample_end=6, samples_per_pass=2)
    idx, val = r.halton(np.arange(64, dtype=np.uint32) % 48, np.arange(64, dtype=np.uint32) % 48, np.arange(64, dtype=np.uint32) % 6, np.arange(64, dtype=np.uint32) % 40)
    film8, _ = r.render(srgb8=True)
    print("%-44s halton sum %.6f srgb8 mean %.3f" % ("halton + output stage", float(val.sum()), float(film8.mean())), flush=True)
    r.close()
    for name in ("plastic", "uber", "mix"):
        run("general material " + name, cpbrt.Scene.parse(_cornell_with_white_material(GENERAL_MATERIALS[name], xres=32, yres=32, spp=2), base_dir=SCENES)).close()
    run("sphere.pbrt directlighting", cpbrt.Scene.parse(scene_text("sphere.pbrt", xres=64, yres=36, spp=4), base_dir=SCENES)).close()
    run("directlighting over delta lobes", cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=32, yres=32, spp=2, integrator=("directlighting", 4)), base_dir=SCENES)).close()
    os.environ["CPBRT_BVH_BUILD"] = "gpu"; os.environ["CPBRT_BVH_GPU_MINLEN"] = "64"
    run("height field 60 x 50, device SAH levels", cpbrt.Scene.parse(_synthetic_mesh_scene(60, 50, res=48))).close()
    r = cpbrt.Renderer(0); r.upload(cornell); r.build_bvh(); r.render()
    checked, bad, site = r.check_failures()
    if checked and not bad:      # the checker itself: one deliberate violation must be counted, with its site
        os.environ["CPBRT_CHECK_SELFTEST"] = "1"
        _, bad2, site2 = r.check_failures()
        os.environ.pop("CPBRT_CHECK_SELFTEST")
        print("self-test of the checker: %d violation(s) counted, site %d (expected 1, site 101)" % (bad2, site2), flush=True)
        if (bad2, site2) != (1, 101):
            sys.exit(2)
    r.close()
    print("sanitize workload done; checked build: %s; bounds violations: %d%s" % (checked, bad, " (first site %d)" % site if bad else ""), flush=True)
    if checked and bad:
        sys.exit(1)


if __name__ == "__main__":
    main()
