#!/usr/bin/env python3
"""Run under torchrun with N >= 2 ranks: sample-partitioned render + NCCL film reduce must equal
the single-GPU render of the same scene (up to f32 summation order)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__
    cpbrt = __graft_entry__.build()
    from conftest import SCENES, scene_text
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = cpbrt.Scene.parse(scene_text("cornel_box.pbrt", xres=160, yres=120, spp=16), base_dir=SCENES)
    r = cpbrt.Renderer(local); r.upload(scene); r.build_bvh()
    idt = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    r.comm_init(world, rank, bytes(idt.cpu().tolist()))
    film, st = r.render()
    total = torch.tensor([float(st.camera_samples)], device="cuda", dtype=torch.float64)
    dist.all_reduce(total)
    ok = True
    if rank == 0:
        single = cpbrt.Renderer(local); single.upload(scene); single.build_bvh()
        ref, _ = single.render()
        err = np.abs(film - ref) / (np.abs(ref) + 1e-3)
        ok = bool(err.max() < 1e-4) and int(total.item()) == 160 * 120 * 16
        print("multi-gpu check: ranks=%d max rel err=%.3g samples=%d reduce_ms=%.3f -> %s" % (world, err.max(), int(total.item()), st.ms_reduce, "OK" if ok else "FAIL"))
    # second phase: communicator attached BEFORE the build -> rank 0 builds, the others receive the
    # finished BVH by ncclBroadcast (forced here; by default only scenes of >= 65536 primitives share)
    os.environ["CPBRT_BVH_SHARE"] = "1"
    r2 = cpbrt.Renderer(local); r2.upload(scene)
    idt2 = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt2.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt2, 0)
    r2.comm_init(world, rank, bytes(idt2.cpu().tolist()))
    order2 = r2.build_bvh()
    o = torch.from_numpy(order2.astype(np.int64)).cuda()
    o0 = o.clone(); dist.broadcast(o0, 0)
    same_order = torch.tensor([1.0 if bool((o == o0).all()) else 0.0], device="cuda")
    dist.all_reduce(same_order, op=dist.ReduceOp.MIN)
    film2, _ = r2.render()
    ok2 = True
    if rank == 0:
        err2 = np.abs(film2 - ref) / (np.abs(ref) + 1e-3)
        ok2 = bool(err2.max() < 1e-4) and same_order.item() == 1.0 and np.array_equal(film2.view(np.uint32), film.view(np.uint32))
        print("shared-build check: same order on all ranks=%s, max rel err=%.3g, film identical to the replicated-build run=%s -> %s" % (
            same_order.item() == 1.0, err2.max(), np.array_equal(film2.view(np.uint32), film.view(np.uint32)), "OK" if ok2 else "FAIL"))
    ok = ok and ok2
    # third phase: the same hand-over when rank 0's build never left the device (device-resident SAH build: the
    # broadcast arrays are pieces of one device allocation there) -- a height field, device build forced
    from test_oracle import _synthetic_mesh_scene
    mesh = cpbrt.Scene.parse(_synthetic_mesh_scene(120, 100, res=64))
    os.environ["CPBRT_BVH_BUILD"] = "gpu"
    r3 = cpbrt.Renderer(local); r3.upload(mesh)
    idt3 = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt3.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt3, 0)
    r3.comm_init(world, rank, bytes(idt3.cpu().tolist()))
    order3 = r3.build_bvh()
    resident = r3.bvh_device_layout()["resident"] if rank == 0 else True
    film3, _ = r3.render()
    os.environ.pop("CPBRT_BVH_SHARE"); os.environ["CPBRT_BVH_BUILD"] = "host"
    ok3 = True
    if rank == 0:
        s3 = cpbrt.Renderer(local); s3.upload(mesh); order_host = s3.build_bvh(); ref3, _ = s3.render()
        err3 = np.abs(film3 - ref3) / (np.abs(ref3) + 1e-3)
        ok3 = bool(resident) and np.array_equal(order3, order_host) and bool(err3.max() < 1e-4)
        print("resident shared-build check: resident=%s, order equals the host build=%s, max rel err=%.3g -> %s" % (
            resident, np.array_equal(order3, order_host), err3.max(), "OK" if ok3 else "FAIL"))
    os.environ.pop("CPBRT_BVH_BUILD")
    ok = ok and ok3
    # fourth phase: ranks that disagree about the scene (rank 0: the Cornell box, the others: sphere.pbrt) must ALL come
    # back from the shared build with an error -- nobody may be left waiting inside a broadcast -- and the same
    # contexts must still work afterwards with the right scene
    os.environ["CPBRT_BVH_SHARE"] = "1"
    other = cpbrt.Scene.parse(scene_text("sphere.pbrt", xres=160, yres=120, spp=16), base_dir=SCENES)
    r4 = cpbrt.Renderer(local); r4.upload(scene if rank == 0 else other)
    idt4 = torch.zeros(cpbrt.abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt4.copy_(torch.tensor(list(cpbrt.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt4, 0)
    r4.comm_init(world, rank, bytes(idt4.cpu().tolist()))
    failed = 0.0
    try:
        r4.build_bvh()
    except cpbrt.CpbrtError:
        failed = 1.0
    all_failed = torch.tensor([failed], device="cuda"); dist.all_reduce(all_failed, op=dist.ReduceOp.MIN)
    r4.upload(scene); r4.build_bvh(); film4, _ = r4.render()
    os.environ.pop("CPBRT_BVH_SHARE")
    if rank == 0:
        ok4 = all_failed.item() == 1.0 and np.array_equal(film4.view(np.uint32), film.view(np.uint32))
        print("mismatched-scene check: every rank got an error=%s, contexts usable afterwards and film identical=%s -> %s" % (
            all_failed.item() == 1.0, np.array_equal(film4.view(np.uint32), film.view(np.uint32)), "OK" if ok4 else "FAIL"))
        ok = ok and ok4
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
