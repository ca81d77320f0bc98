"""Runs the CPU oracle (restated sequential reference, oracle/acvd_oracle.cpp) TO CONVERGENCE on the BASELINE.json
configurations C1, C2 and C3 and freezes what it found in tests/golden/oracle_baseline_runs.json:

    loops, convergence events, vertex tests, modifications, seconds on this box (1 core), final energy
    (ComputeGlobalEnergy after a fresh ReComputeStatistics), translation-invariant energy sum w |p - c|^2 where it is
    defined (isotropic / QEM), sha256 of the initial sampling and of the final clustering.

The GPU parity tests at the BASELINE sizes (tests/test_gpu_parity.py::test_baseline_size_energy_vs_oracle_fixture)
compare the CUDA path's converged energy with these numbers (1 % bar of BASELINE.json north_star) without having to
re-run 30-200 s of sequential CPU work on the GPU box; C1 is also re-run live there (0.3 s).

The upstream binary cannot run here (VTK is absent): this pins the restatement, not upstream.
Run:  python tests/golden/make_baseline_runs.py [C1 C2 C3]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from acvd_b200 import meshgen  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_baseline_runs.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run(name):
    w = meshgen.workload(name)
    uncon = 1 if w["metric"] == "qem" else 0          # ACVDQ.cxx:325
    o = oracle.Oracle(w["points"], w["triangles"])
    o.build_metric(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    o.set_num_clusters(w["K"])
    o.set_params(unconstrained_init=uncon)
    t0 = time.perf_counter()
    cl0 = o.initial_sampling().copy()
    t_sampling = time.perf_counter() - t0
    t0 = time.perf_counter()
    o.minimize()
    t_min = time.perf_counter() - t0
    r = o.report()
    e_left = o.global_energy()                        # incremental sums, as the reference leaves them
    o.recompute_statistics()
    e = o.global_energy()
    cl = o.clustering()
    out = dict(V=int(o.V), F=int(o.F), K=int(w["K"]), metric=w["metric"], gradation=float(w["gradation"]),
               unconstrained_init=uncon, loops=r["loops"], convergences=r["convergences"], tests=r["tests"],
               modifications=r["mods"], seconds_minimize=t_min, seconds_initial_sampling=t_sampling, cores=1,
               energy=e, energy_incremental=e_left, sha256_initial_sampling=sha(cl0), sha256_clustering=sha(cl))
    if w["metric"] in ("iso", "qem"):
        items = o.items()
        _, cen, _, _ = o.cluster_stats()
        out["true_energy"] = oracle.true_energy(w["points"], items[:, 3], cl, cen)
    print(name, json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or ["C1", "C2", "C3"]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in names:
        res[n] = run(n)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
