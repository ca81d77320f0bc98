"""Regenerates tests/golden/oracle_ico6_k12.json and oracle_curv_edges_ico4.json from the CPU oracle (oracle/acvd_oracle.cpp).

The reference itself cannot run here (VTK is absent), so this fixture pins the *oracle*, not upstream:
it guards the restatement against accidental change.  Run: python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from acvd_b200 import meshgen  # noqa: E402
from oracle import oracle  # noqa: E402

p, t = meshgen.geodesic_icosphere(6)
K = 12
out = {}
for metric, uncon in (("iso", 0), ("qem", 1)):
    o = oracle.Oracle(p, t)
    o.build_metric(metric)
    o.set_num_clusters(K)
    o.set_params(unconstrained_init=uncon)
    cl0 = o.initial_sampling().copy()
    o.minimize()
    o.recompute_statistics()
    out[metric] = dict(initial_sampling=cl0.tolist(), clustering=o.clustering().tolist(), energy=o.global_energy(),
                       dual_triangles=o.dual_triangles().tolist(), loops=o.report()["loops"])
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_ico6_k12.json"), "w"))
print("written", {k: (v["energy"], v["loops"]) for k, v in out.items()})

# curvature (vtkCurvatureMeasure restatement) and the reference's edge numbering (what Subdivide's midpoints follow)
p, t = meshgen.ridged_ellipsoid(4)             # V = 162
o = oracle.Oracle(p, t)
ind, info = o.curvature(3)
a, b = o.edges()
json.dump(dict(indicator=ind.tolist(), info=info.astype(float).tolist(), edge_v1=a.tolist(), edge_v2=b.tolist()),
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_curv_edges_ico4.json"), "w"))
print("written curvature/edges fixture:", ind.shape[0], "vertices,", a.shape[0], "edges")
