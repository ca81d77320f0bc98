"""Kernel micro-benchmark on the GPU box: times the dense bulk scan variants on a realistic mid-run state.

usage: python scripts/bench_kernels.py [workload] [bulk rounds before timing] [variants, comma separated] [launches per variant]
Runs `n` bulk rounds of the normal path (so cluster ids, centroids and modified bits are those of a real round),
then times every (stages, blocks/SM) variant of k_scan_bulk_dense and the list-based k_scan on that state.
"""
import sys

sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen  # noqa: E402

VARIANTS = {-1: "k_scan<W,true> (list)", 0: "fused S=3 B=4", 1: "fused S=2 B=4",
            20: "gen3 S=3 B=4", 22: "gen3 S=3 B=4 pf2", 25: "gen3 static S=3 B=4 pf2", 30: "gen3 S=3 B=4 no-decision", 31: "gen3 S=3 B=4 no-gathers",
            40: "split S=3 B=6 vpl1", 42: "split S=2 B=4 vpl2 (shipped)", 44: "split S=2 B=5 vpl2", 46: "split S=2 B=4 vpl4"}

if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    only = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    w = meshgen.workload(wl)
    g = capi.Context(0)
    g.set_mesh(w["points"], w["triangles"])
    g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    g.set_num_clusters(int(w["K"]))
    g.initial_sampling()
    rep = g.minimize(unconstrained_init=1, max_loops=n)
    V = w["points"].shape[0]
    print(f"{wl}: V={V} state after {rep['rounds']} rounds ({rep['bulk_rounds']} bulk)", flush=True)
    for stage in (0, 1):
        for v, name in VARIANTS.items():
            if only is not None and v not in only:
                continue
            ms = g.bench_kernel(0, v, stage, reps)
            print(f"stage {stage} variant {v:2d} {name:24s} {1e3*ms:8.1f} us/launch  {V*56/ms/1e6:7.1f} GB/s (56 B/vertex)", flush=True)
    g.close()
