#!/usr/bin/env python3
"""Benchmark of the curry-pbrt rendering hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--spp S] [--impl reference]

One "step" = one complete render() of scenes/cornel_box.pbrt (900x900, path integrator,
maxdepth 5) at --spp samples per pixel (default 1024, the configuration BASELINE.json's target is
quoted on).  The metric is camera samples per second (Msamples/s); Mrays/s, the traversal
roofline and the CPU baseline ride along in the same JSON line.

  value   scene and BVH resident in HBM, film left on the device; device time, max over ranks
  e2e     the same render through the C-ABI with host buffers: cpbrt_upload_scene (H2D of the
          scene + Halton tables) + cpbrt_build_bvh + cpbrt_render into a host film (D2H)
  N > 1   one process per GPU (torchrun); every rank renders its share of the sample indices of
          every pixel, one ncclReduce sums the film accumulators on rank 0 ("strong" scaling:
          the job is fixed, each GPU gets 1/N of it)

--impl reference times the CPU restatement of the reference's render loop (oracle/, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SCENE = os.path.join(ROOT, "scenes", "cornel_box.pbrt")
PEAKS = os.path.join(ROOT, "MEASURED_PEAKS.json")
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback


def cpu_sample_rect(width, height):
    """Bounded CPU sample: the whole film at a reduced sample count (--cpu-spp), ~7 s on 16 cores."""
    return (0, 0, width, height)


def run_cpu_baseline(cpbrt, spp_sample, threads, count_visits):
    """Oracle (C++ restatement of src/render.rs:7-50) on the whole film at spp_sample samples per pixel."""
    import oracle_api
    scene = cpbrt.Scene.load(SCENE)
    scene.set_spp(spp_sample)
    world = oracle_api.OracleWorld(scene)
    w, h = scene.resolution
    rect = cpu_sample_rect(w, h)
    _, sec, st = world.render(rect=rect, threads=threads, count_visits=count_visits)
    n = (rect[2] - rect[0]) * (rect[3] - rect[1]) * spp_sample
    return dict(samples=n, seconds=sec, msamples_s=n / sec / 1e6, mrays_s=(st["rays_closest"] + st["rays_any"]) / sec / 1e6,
                stats=st, rect=rect, spp=spp_sample)


def traversal_roofline(cpbrt, cornell_renderer):
    """Roofline of the BVH traversal on the traversal-bound configuration C4 (1,000,008 triangles, written as
    .pbrt + binary PLY and parsed like any scene): 2^22 random rays resident in HBM through the SHIPPED
    closest-hit kernel (vote-scheduled two-phase walk), timed with CUDA events on the library's stream.

    achieved = the kernel's OWN work, counted by the same kernel in its counting variant on the same rays:
               node steps x 64 B (one Parent record) + primitive tests x 48 B (one record) + 48 B ray in / hit out
    peak     = this GPU's L2 read bandwidth, measured in the same run by a streaming kernel over 64 MB
               (the 94 MB of nodes + triangles live in the 126 MB L2; DRAM traffic is a few % of peak)
    The reference's own visit counts (far-first, loose boxes) are 2-3x larger and are reported as
    work_avoided_vs_reference, not as a fraction of anything."""
    import importlib
    import numpy as np
    import torch
    from conftest import lcg_rays
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    cfg = dict(syn.CONFIGS["C4"]); cfg["spp"] = 1
    timing = {}
    scene = syn.load_heightfield(cpbrt, timing=timing, **cfg)
    r = cpbrt.Renderer(torch.cuda.current_device())
    t0 = time.perf_counter(); r.upload(scene); r.build_bvh(); build_s = time.perf_counter() - t0
    l2_gbs = r.read_bandwidth(64 << 20, 50)
    n = 1 << 22
    rays = torch.from_numpy(lcg_rays(n, (0, 0, 0), (556, 548, 559), seed=7)).cuda()
    hits = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    own = r.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), count="own")
    ref = r.trace_closest_device(rays.data_ptr(), n, hits.data_ptr(), count=True)
    ms = []
    for i in range(6):
        flush.fill_(i); torch.cuda.synchronize()            # L2 flushed before every launch
        st = r.trace_closest_device(rays.data_ptr(), n, hits.data_ptr())
        if i:
            ms.append(st.ms_trace)
    ms_avg = sum(ms) / len(ms)
    own_bytes = 64.0 * own.trace_parent_visits + 48.0 * own.trace_prim_tests + 48.0 * n
    ref_bytes = 32.0 * ref.trace_parent_visits + 48.0 * ref.trace_prim_tests + 48.0 * n
    achieved = own_bytes / (ms_avg * 1e-3) / 1e9
    ncu = None
    npath = os.path.join(ROOT, "profiles", "r02_ncu_c4_traversal.json")
    if os.path.exists(npath):
        ncu = json.load(open(npath))
    out = dict(bound="l2", kernel="k_batch_closest<vote> (the closest-hit traversal cpbrt_render uses for scenes of >= 256 primitives)",
               workload="C4: 1,000,008 triangles (.pbrt + binary PLY), 2^22 random rays (LCG seed 7) resident in HBM, L2 flushed before every launch",
               achieved=achieved, peak=l2_gbs, unit="GB/s", frac=achieved / l2_gbs,
               peak_source="measured in this run: read-only streaming kernel over 64 MB (cpbrt_measure_read_bandwidth)",
               avg_launch_ms=ms_avg, launches_timed=len(ms), grays_s=n / (ms_avg * 1e-3) / 1e9,
               own_node_steps_per_ray=own.trace_parent_visits / n, own_prim_tests_per_ray=own.trace_prim_tests / n,
               algorithmic_bytes_per_ray=own_bytes / n,
               work_avoided_vs_reference=dict(reference_parent_visits_per_ray=ref.trace_parent_visits / n, reference_prim_tests_per_ray=ref.trace_prim_tests / n,
                                              reference_bytes_per_ray=ref_bytes / n, ratio=ref_bytes / own_bytes),
               traffic=(ncu or {}).get("dram_bytes_per_launch"), traffic_l2=(ncu or {}).get("lts_bytes_per_launch"),
               traffic_rays_per_launch=(ncu or {}).get("rays_per_launch"),
               traffic_unit="bytes per launch of the same kernel on the same ray batch, from one ncu --set full capture (profiles/r02_ncu_c4_traversal.json): "
                            "traffic = dram__bytes_read.sum + dram__bytes_write.sum, traffic_l2 = lts__t_bytes.sum; compare traffic_l2 with "
                            "algorithmic_bytes_per_ray x traffic_rays_per_launch (the tree lives in L2, DRAM only sees it once per launch)",
               issue=(ncu or {}).get("issue"),
               hbm_peak_gbs=(json.load(open(PEAKS))["hbm_gbs"] if os.path.exists(PEAKS) else FALLBACK_HBM_GBS),
               scene_load=dict(ply_write_s=timing.get("write_s"), parse_s=timing.get("parse_s"), upload_and_bvh_build_s=build_s),
               note="the kernel is instruction-issue bound (divergent tree walk); see DESIGN.md 5 for the issue-slot figures from the same capture. "
                    "The Cornell box of this bench line has 33 primitives and is traced without a tree walk (flat search from shared memory): no memory roofline applies there.")
    r.close()
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.samples = []
        self.stop_flag = threading.Event()
        self.gpu = gpu_index

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower().startswith("active")})
        pw = sorted(float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit())
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_min_mhz=sm[0] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    power_w_median=pw[len(pw) // 2] if pw else None, power_w_max=pw[-1] if pw else None, n=len(self.samples))


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its
    version banner to fd 1 when the communicator is created), so fd 1 is pointed at stderr for the
    whole run and the JSON line goes to a private copy of the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--spp", type=int, default=1024)
    ap.add_argument("--impl", default="cpbrt", choices=["cpbrt", "reference"])
    ap.add_argument("--cpu-spp", type=int, default=32, help="spp of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--samples-per-pass", type=int, default=0, help="0 = library default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import __graft_entry__
    cpbrt = __graft_entry__.build()
    config = dict(workload="scenes/cornel_box.pbrt 900x900 path maxdepth=5 halton %d spp" % args.spp, spp=args.spp,
                  film=[900, 900], primitives=33, partition="sample-index" if world > 1 else "none",
                  l2_policy="working set exceeds L2: every pass rewrites ~5.8 GB of path state (32 spp x 810k pixels x 224 B) "
                            "plus its queues, against 126 MB of L2; no flush needed")

    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        vals = []
        for i in range(args.warmup + args.steps):
            r = run_cpu_baseline(cpbrt, args.cpu_spp, threads, False)
            if i >= args.warmup:
                vals.append(r)
        sec = sum(v["seconds"] for v in vals)
        n = sum(v["samples"] for v in vals)
        v = n / sec / 1e6
        sample = "whole film %s at %d spp (%d samples) of the 1024 per step" % (list(vals[0]["rect"]), args.cpu_spp, vals[0]["samples"])
        json_out.write(json.dumps(dict(metric="Msamples/s", value=v, unit="Msamples/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=1e3 * sec / len(vals), higher_is_better=True, scaling="strong", vs_baseline=None,
                              dtype="f32", data="synthetic", impl="reference", config=config,
                              mrays_s=sum(x["mrays_s"] * x["seconds"] for x in vals) / sec,
                              cpu_baseline=dict(value=v, unit="Msamples/s", cores=threads, kind="port", sample=sample),
                              e2e=dict(value=v, unit="Msamples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))) + "\n")
        json_out.flush()
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = cpbrt.Scene.load(SCENE)
    scene.set_spp(args.spp)
    W, H = scene.resolution
    r = cpbrt.Renderer(local_rank)
    r.upload(scene)
    r.build_bvh()
    if world > 1:
        idt = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        r.comm_init(world, rank, bytes(idt.cpu().tolist()))
    film_dev = torch.zeros(H * W * 3, dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(profile=False):
        kw = dict(profile_stages=profile, samples_per_pass=args.samples_per_pass)
        _, st = r.render(device_film_ptr=film_dev.data_ptr(), **kw) if (rank == 0) else r.render(**kw)
        return st

    for _ in range(args.warmup):
        step()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, rays_c, rays_a, rays_ct = 0.0, 0, 0, 0, 0
    step_ms_list = []
    for _ in range(args.steps):
        st = step()
        step_ms_list.append(round(st.ms_total, 2))
        dev_ms += st.ms_total; launches += st.kernel_launches; rays_c += st.rays_closest; rays_a += st.rays_any; rays_ct += st.rays_closest_traced
    barrier()
    wall = time.perf_counter() - t0
    clocks.stop_flag.set()

    # max over ranks of the device time; totals over ranks for rays / launches
    agg = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(launches), float(rays_c), float(rays_a), float(rays_ct)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms = float(agg[0]), float(agg[1])
    launches, rays_c, rays_a, rays_ct = (int(x) for x in tot.tolist())
    samples = W * H * args.spp * args.steps
    # the step time is the larger of the summed device time and the barrier-to-barrier wall time
    step_ms = max(dev_ms, wall_ms) / args.steps
    value = samples / args.steps / (step_ms * 1e-3) / 1e6

    extra = {}
    # ---- e2e: the call a user makes, host buffers in and out, at N GPUs ---------------------------
    # every step: cpbrt_upload_scene (H2D of the scene + Halton tables) + cpbrt_build_bvh +
    # cpbrt_render into a host film on rank 0 (D2H); barrier-to-barrier wall time, max over ranks
    # the host film is pinned memory (the read-back is then one DMA, not a staged copy through a bounce buffer)
    host_film = None
    if rank == 0:
        pinned = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
        host_film = pinned.numpy()
    e2e_steps = max(1, args.steps)
    for _ in range(2):          # warm-up of the upload / build path (first-use allocations)
        r.upload(scene); r.build_bvh()
    h2d_before = r.h2d_bytes()
    phase = [0.0, 0.0, 0.0]
    e2e_each, e2e_dev = [], []
    clocks2 = ClockSampler(local_rank)
    if rank == 0:
        clocks2.start()
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        ta = time.perf_counter(); r.upload(scene)
        tb = time.perf_counter(); r.build_bvh()
        tc = time.perf_counter(); _, est = r.render(film=host_film) if rank == 0 else r.render()
        td = time.perf_counter()
        phase[0] += tb - ta; phase[1] += tc - tb; phase[2] += td - tc
        e2e_each.append(round(1e3 * (td - ta), 2)); e2e_dev.append(round(est.ms_total, 2))
    barrier()
    e2e_wall = time.perf_counter() - t1
    clocks2.stop_flag.set()
    e2e_t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    h2d = torch.tensor([float(r.h2d_bytes() - h2d_before)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
    e2e_s = float(e2e_t[0]) / e2e_steps
    extra["e2e"] = dict(value=W * H * args.spp / e2e_s / 1e6, unit="Msamples/s", h2d_bytes_per_step=int(float(h2d[0]) / e2e_steps),
                        d2h_bytes_per_step=W * H * 3 * 4, ms_per_step=e2e_s * 1e3, steps=e2e_steps, rank0_ms_of_each_step=e2e_each, rank0_device_ms_of_each_step=e2e_dev, clocks=clocks2.summary() if rank == 0 else None,
                        rank0_ms_per_step=dict(upload_scene=1e3 * phase[0] / e2e_steps, build_bvh=1e3 * phase[1] / e2e_steps, render_and_readback=1e3 * phase[2] / e2e_steps),
                        note="scene upload + BVH build + render + film read-back into pinned host memory per step, every rank uploads its own replica; "
                             "h2d bytes are counted by the library where the copies are issued (cpbrt_h2d_bytes), summed over ranks")

    # ---- stage profile + e2e + roofline + CPU baseline on rank 0 / N == 1 ---------------------
    if world == 1:
        st = step(profile=True)
        trace_ms = st.ms_trace_closest
        extra["stage_ms"] = dict(trace_closest=st.ms_trace_closest, trace_shadow=st.ms_trace_shadow, shade=st.ms_shade,
                                 raygen_film=st.ms_raygen_film, total=st.ms_total)
        # CPU baseline (also yields the algorithmic traversal bytes per ray)
        cores = os.cpu_count() or 1
        if not args.no_cpu_baseline:
            cb = run_cpu_baseline(cpbrt, args.cpu_spp, cores, True)
            s = cb["stats"]
            bytes_closest = 48.0 + (32.0 * s["closest_parent_visits"] + 48.0 * s["closest_prim_tests"]) / max(1, s["rays_closest"])
            extra["cpu_baseline"] = dict(value=cb["msamples_s"], unit="Msamples/s", cores=cores, kind="port", mrays_s=cb["mrays_s"],
                                         sample="whole film %s at %d spp (%d samples) of the 1024, C++ restatement of the reference CPU path"
                                                % (list(cb["rect"]), cb["spp"], cb["samples"]))
            extra["reference_visit_bytes_per_ray_cornell"] = bytes_closest      # SURVEY 8d figure (the REFERENCE's far-first loose-box walk), kept for context
        extra["roofline"] = traversal_roofline(cpbrt, r)
        # the stage that dominates THIS workload (Cornell box): shading, bound by path-state traffic and latency.
        # Algorithmic bytes per shaded hit = the 32-byte state records a bounce must read (ray, throughput, sampler
        # state, pending direct-lighting terms) and write (new ray, throughput, state, shadow ray, two direct-lighting
        # records, MIS ray) + queue entries: 4 x 32 + 7 x 32 + 16 = 368 B.
        hbm = float(json.load(open(PEAKS))["hbm_gbs"]) if os.path.exists(PEAKS) else FALLBACK_HBM_GBS
        shade_bytes = 368.0 * st.shade_items
        extra["roofline_shading_stage"] = dict(bound="hbm", kernel="k_classify + k_shade<*> (all launches of the profiled step)",
                                               achieved=shade_bytes / (st.ms_shade * 1e-3) / 1e9, peak=hbm, unit="GB/s",
                                               frac=shade_bytes / (st.ms_shade * 1e-3) / 1e9 / hbm, bytes_per_shaded_hit=368, shaded_hits=int(st.shade_items),
                                               stage_ms=st.ms_shade, peak_source="MEASURED_PEAKS.json hbm_gbs" if os.path.exists(PEAKS) else "fallback",
                                               note="latency-bound: the matte kernel issues on 44 % of cycles at 24 resident warps per SM (80 registers), DRAM 30 % of peak (ncu, profiles/r02_final_ncu_cornell.txt)")
    if rank == 0:
        out = dict(metric="Msamples/s", value=value, unit="Msamples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=step_ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32", data="synthetic",
                   config=config, mrays_s=(rays_c + rays_a) / args.steps / (step_ms * 1e-3) / 1e6,
                   mrays_traced_s=(rays_ct + rays_a) / args.steps / (step_ms * 1e-3) / 1e6,
                   rays_note="mrays_s counts the Scene::intersect / intersect_predicate calls of the reference whose result is used; "
                             "mrays_traced_s only those that needed a BVH traversal here (DESIGN.md 4.2)",
                   gpu_launches=launches,
                   device_ms_per_step=dev_ms / args.steps, wall_ms_per_step=wall_ms / args.steps, rank0_device_ms_of_each_step=step_ms_list,
                   clocks=clocks.summary())
        out.update(extra)
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
