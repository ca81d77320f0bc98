"""Multi-GPU correctness check (run under torchrun, one rank per GPU):
the N-GPU clustering must be bit-identical to the single-GPU clustering on the same input.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dist_check.py [workload ...]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from acvd_b200 import capi, meshgen  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    names = sys.argv[1:] or ["C1", "C2s"]
    ok = True
    for name in names:
        w = meshgen.workload(name)
        uncon = 1 if w["metric"] == "qem" else 0
        g = capi.Context(local)
        uid = [capi.Context.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.dist_init(rank, world, uid[0])
        g.set_mesh(w["points"], w["triangles"])
        g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
        g.set_num_clusters(w["K"])
        g.initial_sampling()
        g.save_clustering()
        passes = int(os.environ.get("DIST_CHECK_PASSES", "0"))     # 0 = the library default (the same on any number of GPUs)
        bulk = int(os.environ.get("DIST_CHECK_BULK", "1000"))      # bulk rounds forced on (automatic = off below 500 k vertices): the exchange of the bulk rounds is what is checked
        g.minimize(unconstrained_init=uncon, commit_passes=passes, bulk_rounds=bulk)     # warm-up (NCCL channels, allocations)
        g.restore_clustering()
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        rep = g.minimize(unconstrained_init=uncon, commit_passes=passes, bulk_rounds=bulk)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        cl = g.clustering()
        # all ranks hold the same replica
        t = torch.from_numpy(cl).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        same_across_ranks = bool((t == ref).all().item())
        if rank == 0:
            s = capi.Context(local)              # single-GPU run of the same problem
            s.set_mesh(w["points"], w["triangles"])
            s.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
            s.set_num_clusters(w["K"])
            s.initial_sampling()
            s.save_clustering()
            s.minimize(unconstrained_init=uncon, commit_passes=passes, bulk_rounds=bulk)
            s.restore_clustering()
            t0 = time.perf_counter()
            rep1 = s.minimize(unconstrained_init=uncon, commit_passes=passes, bulk_rounds=bulk)
            dt1 = time.perf_counter() - t0
            cl1 = s.clustering()
            identical = bool(np.array_equal(cl, cl1))
            print(f"[dist_check] {name}: world={world} rounds {rep['rounds']} vs {rep1['rounds']}, energy {rep['energy']:.12g} vs "
                  f"{rep1['energy']:.12g}, identical={identical}, replicas_equal={same_across_ranks}, "
                  f"t_dist={dt:.3f}s t_single={dt1:.3f}s tests {rep['tests']} vs {rep1['tests']}", flush=True)
            ok = ok and identical and same_across_ranks and rep["energy"] == rep1["energy"]
            s.close()
        else:
            ok = ok and same_across_ranks
        g.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("[dist_check] PASS" if flag.item() else "[dist_check] FAIL", flush=True)
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
