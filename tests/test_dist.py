"""N > 1: host-side logic on CPU with gloo (world size 2), and -- on a box with >= 2 GPUs -- the NCCL path,
whose clustering must be bit-identical to the single-GPU one (scripts/dist_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acvd_b200 import partition
    import torch
    # the unique id made on rank 0 reaches every rank unchanged
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    # tile ranges are disjoint, contiguous and cover the mesh
    V = 1000003
    t0, t1 = partition.tile_range(V, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (t0, t1))
    # whole-job metric = units of all ranks / max-over-ranks time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, uid[0] == bytes(range(128)), spans, float(t.item())))
    dist.destroy_process_group()


def test_gloo_world2_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, uid_ok, spans, tmax in res:
        assert uid_ok and tmax == 2.0
        n_tiles = (1000003 + 31) // 32
        assert spans[0][0] == 0 and spans[-1][1] == n_tiles
        assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))


def test_tile_range_properties():
    from acvd_b200 import partition
    for V in (1, 31, 32, 33, 163842, 40000002):
        for world in (1, 2, 3, 8):
            spans = [partition.tile_range(V, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == (V + 31) // 32
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.gpu
def test_two_gpus_bit_identical_to_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py"), "C1", "C2s"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert "[dist_check] PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
