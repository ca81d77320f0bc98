"""First-contact script for a GPU box: exercises every stage with prints (not a test, not a bench)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen  # noqa: E402
from oracle import oracle  # noqa: E402


def run(name, p, t, metric, K, uncon, gradation=0.0, ind=None, with_oracle=True):
    print(f"=== {name}: V={p.shape[0]} F={t.shape[0]} K={K} metric={metric}", flush=True)
    g = capi.Context(0)
    t0 = time.time(); g.set_mesh(p, t); print(f"set_mesh {time.time()-t0:.3f}s E={g.num_edges()}", flush=True)
    t0 = time.time(); g.build_items(metric, gradation, ind); print(f"build_items {time.time()-t0:.3f}s", flush=True)
    g.set_num_clusters(K)
    t0 = time.time(); g.initial_sampling(); print(f"initial_sampling {time.time()-t0:.3f}s", flush=True)
    cl0 = g.clustering()
    t0 = time.time(); rep = g.minimize(unconstrained_init=uncon); dt = time.time() - t0
    print(f"gpu minimize {dt:.3f}s", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in rep.items()}, flush=True)
    n = rep["round_launches"]
    print(f"scan: {rep['ms_scan']/n*1e3:.1f} us/launch, {rep['scan_bytes']/(rep['ms_scan']*1e-3)/1e9:.1f} GB/s | "
          f"evaluate: {rep['ms_evaluate']/n*1e3:.1f} us/launch, {rep['evaluate_bytes']/(rep['ms_evaluate']*1e-3)/1e9:.1f} GB/s | "
          f"commit {rep['ms_commit']/n*1e3:.1f} us/launch; tests/s={rep['tests']/dt:.3e}")
    if with_oracle:
        o = oracle.Oracle(p, t)
        o.build_metric(metric, gradation, ind)
        o.set_num_clusters(K)
        o.set_clustering(cl0)
        o.set_params(unconstrained_init=uncon)
        t0 = time.time(); o.minimize(); dt_o = time.time() - t0
        r = o.report()
        o.recompute_statistics()
        e_o = o.global_energy()
        print(f"oracle minimize {dt_o:.3f}s {r} energy={e_o:.12g}")
        print(f"energy gpu={rep['energy']:.12g} rel diff={(rep['energy']-e_o)/abs(e_o):.3e}  speedup={dt_o/dt:.1f}x")
        it = o.items(); w = it[:, 3]
        const = float(np.sum((it[:, :3] ** 2).sum(axis=1) / w))
        print(f"true energy: oracle={const+e_o:.6e} gpu={const+rep['energy']:.6e} ratio={(const+rep['energy'])/(const+e_o):.4f}")
    g.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "C1"]
    if "small" in which:
        p, t = meshgen.geodesic_icosphere(32)
        run("small-iso", p, t, "iso", 200, 0)
        run("small-qem", p, t, "qem", 200, 1)
    if "C1" in which:
        p, t = meshgen.geodesic_icosphere(128)
        run("C1", p, t, "iso", 3000, 0)
        run("C1-qem", p, t, "qem", 3000, 1)
    if "C2" in which:
        w = meshgen.workload("C2")
        run("C2", w["points"], w["triangles"], "qem", w["K"], 1, w["gradation"], w["indicator"], with_oracle="--oracle" in which)
    if "C4" in which:
        w = meshgen.workload("C4")
        run("C4", w["points"], w["triangles"], "qem", w["K"], 1, with_oracle=False)
