#!/usr/bin/env python3
"""A BASELINE configuration on N GPUs of one box (launch with torchrun): parse on every rank,
communicator first, ONE BVH build on rank 0 broadcast to the others, sample-partitioned render,
NCCL film reduce.  Rank 0 prints one JSON line.  Config 5 is the 10-million-triangle height field at
3840x2160 and 4096 spp.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
        tools/config_multi_gpu.py --config C5 [--spp 4096]
"""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5")
    ap.add_argument("--spp", type=int, default=0)
    args = ap.parse_args()
    import numpy as np, torch, torch.distributed as dist
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    syn = importlib.import_module("curry-pbrt_b200.synthetic")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(syn.CONFIGS[args.config])
    if args.spp:
        cfg["spp"] = args.spp

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier(); t0 = time.perf_counter()
    if rank == 0:
        syn.write_heightfield_files(syn.default_dir(), **cfg)      # one writer; the others find the files
    barrier()
    scene = syn.load_heightfield(cpbrt, **cfg)
    barrier(); t_parse = time.perf_counter() - t0
    r = cpbrt.Renderer(local)
    r.upload(scene)
    if world > 1:
        idt = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        r.comm_init(world, rank, bytes(idt.cpu().tolist()))
    barrier(); t_upload = time.perf_counter() - t0 - t_parse
    r.build_bvh()
    barrier(); t_build = time.perf_counter() - t0 - t_parse - t_upload
    W, H = scene.resolution
    # warm-up: NCCL channels, allocations (a 2-sample render)
    r.render(sample_begin=0, sample_end=world * 2) if world == 1 else r.render(samples_per_pass=1, sample_begin=rank * 2, sample_end=rank * 2 + 2)
    barrier(); t1 = time.perf_counter()
    film, st = r.render()
    barrier(); t_render = time.perf_counter() - t1
    tot = torch.tensor([float(st.camera_samples), float(st.rays_closest + st.rays_any), float(st.nan_dropped)], dtype=torch.float64, device="cuda")
    ms = torch.tensor([st.ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        samples = W * H * cfg["spp"]
        print(json.dumps(dict(config=args.config, n_gpus=world, film=[W, H], spp=cfg["spp"], primitives=int(scene.desc.n_primitives),
                              camera_samples=int(tot[0]), samples_expected=samples, msamples_s=samples / t_render / 1e6,
                              mrays_s=float(tot[1]) / t_render / 1e6, render_wall_s=t_render, render_device_s_max=float(ms[0]) * 1e-3,
                              parse_s=t_parse, upload_and_comm_s=t_upload, bvh_build_and_broadcast_s=t_build,
                              device_levels_rank0=r.bvh_gpu_levels(), nan_dropped=int(tot[2]), film_mean=float(film.mean()),
                              film_finite=bool(np.isfinite(film).all()))), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
