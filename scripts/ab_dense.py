"""A/B of dense-scan variants on a small mesh, round by round (diagnostic): python scripts/ab_dense.py <mesh> <variant> ..."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from acvd_b200 import capi, meshgen

if __name__ == "__main__":
    mesh = sys.argv[1]
    p, t = meshgen.bipyramid(24, 4) if mesh == "spindle" else meshgen.geodesic_icosphere(24)
    K = 150
    res = {}
    for v in sys.argv[2:]:
        os.environ["ACVD_DENSE_VARIANT"] = v
        rows = []
        for loops in (1, 2, 3, 4, 6, 8, 12, 20, 40, 80):
            g = capi.Context(0)
            g.set_mesh(p, t); g.build_items("qem", 0.0, None); g.set_num_clusters(K); g.initial_sampling()
            rep = g.minimize(unconstrained_init=1, max_loops=loops, bulk_rounds=1000)
            rows.append((loops, rep["rounds"], rep["bulk_rounds"], rep["tests"], rep["proposals"], rep["modifications"], rep["evaluated"],
                         int(np.bitwise_xor.reduce(g.clustering() * np.arange(1, p.shape[0] + 1, dtype=np.int64)))))
            g.close()
        res[v] = rows
    base = sys.argv[2]
    for v in sys.argv[2:]:
        for a, b in zip(res[base], res[v]):
            print(v, b, "" if a == b else "  <-- differs from %s: %s" % (base, a))
