"""k_cluster_pass under ncu (not a bench): statistics-only pass, then the pass with the connectivity check.
usage: python scripts/prof_cluster_pass.py [workload]"""
import sys

sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen  # noqa: E402

if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    w = meshgen.workload(wl)
    g = capi.Context(0)
    g.set_mesh(w["points"], w["triangles"])
    g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    g.set_num_clusters(int(w["K"]))
    g.initial_sampling()
    g.fill_holes(0)
    g.recompute_statistics(1, 3)
    print("disconnected", g.clean_clustering())
    g.close()
