#!/usr/bin/env python3
"""Where a sample-partitioned render spends its time on every rank (launch with torchrun): the Cornell box of
bench.py at 1024 spp, per rank the share, the time up to the film stage (replay + passes) and the film stage
(contribution + ncclReduce + resolve: includes waiting for the slowest rank), and the same share rendered by the same
GPU without a communicator.  One JSON line per rank."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np, torch, torch.distributed as dist
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    from conftest import SCENES, scene_text
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=900, yres=900, spp=spp), base_dir=SCENES)
    solo = cpbrt.Renderer(local); solo.upload(scene); solo.build_bvh()
    r = cpbrt.Renderer(local); r.upload(scene); r.build_bvh()
    idt = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    r.comm_init(world, rank, bytes(idt.cpu().tolist()))
    b, e = cpbrt.sample_share(spp, rank, world) if hasattr(cpbrt, "sample_share") else __import__("importlib").import_module("curry-pbrt_b200.partition").sample_share(spp, rank, world)
    film_dev = torch.zeros(900 * 900 * 3, dtype=torch.float32, device="cuda")
    rows = []
    for it in range(4):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, st = r.render(device_film_ptr=film_dev.data_ptr())
        torch.cuda.synchronize(); t_mine = time.perf_counter() - t0
        dist.barrier()
        t_all = time.perf_counter() - t0
        _, ss = solo.render(sample_begin=b, sample_end=e, device_film_ptr=film_dev.data_ptr())
        if it:
            rows.append(dict(rank=rank, share=[b, e], wall_ms=t_mine * 1e3, wall_all_ranks_ms=t_all * 1e3, device_ms=st.ms_total,
                             before_film_stage_ms=st.ms_total - st.ms_reduce, film_stage_ms=st.ms_reduce,
                             same_share_without_communicator_ms=ss.ms_total, launches=int(st.kernel_launches)))
    avg = {k: (sum(x[k] for x in rows) / len(rows) if isinstance(rows[0][k], float) else rows[0][k]) for k in rows[0]}
    for k in range(world):
        if k == rank:
            print(json.dumps(avg), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
