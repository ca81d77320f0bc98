"""N > 1: host-side logic on CPU with gloo (world size 2), and -- on a box with >= 2 GPUs -- the NCCL path,
whose clustering must be bit-identical to the single-GPU one (scripts/dist_check.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acvd_b200 import capi
    import torch
    # the unique id made on rank 0 reaches every rank unchanged
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    # the library's own partition (acvd_dist_partition, host arithmetic in libacvd_b200.so): every rank asks for its ranges
    V, F, K = 1000003, 2000002, 10007
    mine = capi.dist_partition(V, F, K, rank, world)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    # whole-job metric = units of all ranks / max-over-ranks time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, uid[0] == bytes(range(128)), parts, float(t.item())))
    dist.destroy_process_group()


def _check_partition(parts, V, F, K):
    world = len(parts)
    n_tiles = (V + 31) // 32
    for key, total in (("tiles", n_tiles), ("points", V), ("faces", F)):
        spans = [p[key] for p in parts]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))           # disjoint, contiguous, covering
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    cl = [p["clusters"] for p in parts]
    chunk = (K + world - 1) // world
    assert cl[0][0] == 0 and max(b for _, b in cl) == K
    assert all(a[1] == b[0] or b[0] == K for a, b in zip(cl, cl[1:]))
    assert all(b - a <= chunk for a, b in cl) and all(a == min(K, r * chunk) for r, (a, _) in enumerate(cl))   # equal chunks (in-place all-gather)


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
    for rank, uid_ok, parts, tmax in res:
        assert uid_ok and tmax == 2.0
        _check_partition(parts, 1000003, 2000002, 10007)


def test_partition_properties():
    from acvd_b200 import capi
    for V, K in ((1, 1), (31, 3), (32, 8), (33, 9), (163842, 3000), (40000002, 400000), (162000000, 1600000)):
        for world in (1, 2, 3, 8):
            _check_partition([capi.dist_partition(V, 2 * V, K, r, world) for r in range(world)], V, 2 * V, K)
    with pytest.raises(capi.AcvdError):
        capi.dist_partition(10, 10, 1, 2, 2)


@pytest.mark.gpu
def test_two_gpus_bit_identical_to_one():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py"), "C1", "C2s"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert "[dist_check] PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
