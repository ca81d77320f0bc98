"""C5-scale run on one GPU (BASELINE configs[4] without its host-side -l / -m passes): ACVD isotropic clustering of the
160 M-vertex thin torus (8:1 elongated cells) into 1.6 M clusters; checks the size-independent properties and prints
stage timings.  usage: python scripts/c5_check.py [nu nv K]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen  # noqa: E402

if __name__ == "__main__":
    nu, nv, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16000, 10000, 1600000)
    t0 = time.time()
    p, t = meshgen.thin_torus(nu, nv)
    print(f"mesh V={p.shape[0]} F={t.shape[0]} generated in {time.time() - t0:.1f}s", flush=True)
    g = capi.Context(0)
    t0 = time.time(); g.set_mesh(p, t); print(f"set_mesh {time.time() - t0:.2f}s E={g.num_edges()}", flush=True)
    t0 = time.time(); g.build_items("iso"); print(f"build_items {time.time() - t0:.2f}s", flush=True)
    g.set_num_clusters(K)
    t0 = time.time(); g.initial_sampling(); print(f"initial_sampling (host) {time.time() - t0:.1f}s", flush=True)
    t0 = time.time(); rep = g.minimize(); dt = time.time() - t0
    print(f"minimize {dt:.3f}s", {k: rep[k] for k in ("rounds", "bulk_rounds", "convergences", "tests", "modifications", "disconnected", "energy", "ms_device")}, flush=True)
    cl = g.clustering()
    sz = np.bincount(cl, minlength=K)
    assert cl.min() >= 0 and cl.max() < K and sz.min() >= 1 and sz.sum() == p.shape[0]
    assert rep["disconnected"] == 0 and g.clean_clustering() == 0
    again = g.reassign_round(1, 3, 1)
    assert again["proposals"] == 0
    t0 = time.time(); n_tri = len(g.dual_triangles()); print(f"dual triangles {n_tri} in {time.time() - t0:.2f}s (Euler: 2K = {2 * K} on a torus)")
    print(f"C5 OK: {rep['tests'] / (rep['ms_device'] * 1e-3):.3e} tests/s, sizes min/mean/max {sz.min()}/{sz.mean():.1f}/{sz.max()}")
