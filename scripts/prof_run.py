"""One clustering run for profiling under ncu (not a bench): python scripts/prof_run.py <workload> [key=value ...]
keys are acvd_params fields (unconstrained_init, bulk_rounds, max_loops, max_convergences, connexity ...)."""
import sys

sys.path.insert(0, ".")
from acvd_b200 import capi, meshgen  # noqa: E402

if __name__ == "__main__":
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
    w = meshgen.workload(wl)
    g = capi.Context(0)
    g.set_mesh(w["points"], w["triangles"])
    g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    g.set_num_clusters(int(w["K"]))
    g.initial_sampling()
    kw.setdefault("unconstrained_init", 1 if w["metric"] == "qem" else 0)
    rep = g.minimize(**kw)
    print({k: rep[k] for k in ("rounds", "bulk_rounds", "tests", "modifications", "energy", "ms_device")})
    g.close()
