import sys, time
sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = meshgen.workload(name)
g = capi.Context(0)
g.set_mesh(w["points"], w["triangles"]); g.build_items(w["metric"], w["gradation"], w["indicator"]); g.set_num_clusters(w["K"])
g.initial_sampling(); g.save_clustering()
for i in range(2):
    g.restore_clustering()
    print("=== run", i, file=sys.stderr)
    r = g.minimize(unconstrained_init=1)
    print({k: round(v, 3) if isinstance(v, float) else v for k, v in r.items()}, file=sys.stderr)
