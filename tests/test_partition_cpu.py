"""Host-side logic of the multi-GPU layout, on CPU with gloo (world_size 2): the sample-index
shares tile [0, spp) exactly, and the communicator id reaches every rank unchanged."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, %r)
import importlib
part = importlib.import_module("curry-pbrt_b200.partition")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ident = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    ident = torch.arange(128, dtype=torch.uint8)
dist.broadcast(ident, 0)
assert ident.tolist() == list(range(128))
for spp in (1, 2, 7, 16, 1024, 4096):
    b, e = part.sample_share(spp, rank, world)
    t = torch.tensor([b, e, e - b], dtype=torch.int64)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    spans = sorted((int(g[0]), int(g[1])) for g in gathered)
    assert spans[0][0] == 0 and spans[-1][1] == spp
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0
    # shares shrink slightly with the rank (replay of the skipped samples), never by much
    assert max(int(g[2]) for g in gathered) - min(int(g[2]) for g in gathered) <= 1 + spp // 100
# film accumulators add up: sum of per-rank (sum, count) == whole
acc = torch.tensor([float(rank + 1), 2.0 * (rank + 1)])
dist.reduce(acc, 0)
if rank == 0:
    assert acc.tolist() == [sum(range(1, world + 1)), 2.0 * sum(range(1, world + 1))]
dist.destroy_process_group()
print("ok", rank)
'''


def test_sample_shares_and_id_broadcast_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_sample_share_matches_library_rule(cpbrt):
    """partition.py mirrors cpbrt_sample_share (the rule cpbrt_render applies): same boundaries for every
    (spp, ranks); the shares tile [0, spp), are non-empty when spp >= ranks, and later ranks never get more."""
    import ctypes as C
    import importlib
    sys.path.insert(0, ROOT)
    part = importlib.import_module("curry-pbrt_b200.partition")
    for spp in (1, 5, 8, 100, 128, 1024, 4096, 65536):
        for world in (1, 2, 3, 4, 8, 16):
            covered, sizes = 0, []
            for rank in range(world):
                b, e = part.sample_share(spp, rank, world)
                lb, le = C.c_uint32(), C.c_uint32()
                cpbrt.lib.cpbrt_sample_share(spp, rank, world, C.byref(lb), C.byref(le))
                assert (b, e) == (lb.value, le.value)
                assert b == covered and e >= b
                covered = e
                sizes.append(e - b)
            assert covered == spp
            assert all(x + 1 >= y for x, y in zip(sizes, sizes[1:]))      # shrinking, up to rounding
            if spp >= 16 * world:
                assert min(sizes) >= 1 and max(sizes) - min(sizes) <= 1 + spp // (10 * world)
    assert part.sample_share(1024, 0, 8) == (0, 129) and part.sample_share(1024, 7, 8) == (897, 1024)
