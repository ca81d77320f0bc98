#!/usr/bin/env python
"""bench.py — ACVDQ clustering to convergence on synthetic meshes (BASELINE.json metric).

One "step" = one full pass of the hot path: acvd_minimize() (MinimizeEnergy of the reference,
Common/vtkUniformClustering.h:725-830) from the same initial sampling to convergence on the workload.
The JSON line reports vertex tests/s (whole job), ms per step (= time to convergence), the roofline of the
dominant kernel (the reassignment/propose kernel), an end-to-end number through the C ABI with host
buffers, and a CPU baseline (the restated reference from oracle/) timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C2|C1|C4s|C2s] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "acvdq_vertex_tests_per_s"
UNIT = "tests/s"

WORKLOADS = {
    # name: (description, generator kwargs)
    "C4": "ACVDQ -s 100 on a 40,000,002-vertex displaced geodesic sphere -> 400k clusters (BASELINE configs[3])",
    "C4s": "1/100-scale C4: 400,002-vertex displaced sphere -> 4k clusters",
    "C2": "ACVDQ gradation 1.5 (analytic curvature) on a 2.6M-vertex noisy torus -> 100k clusters (configs[1])",
    "C2s": "1/16-scale C2",
    "C1": "ACVD isotropic on a 163,842-vertex icosphere -> 3000 clusters (configs[0])",
    "C3": "AnisotropicRemeshingQ gradation 1.5 (analytic curvature) on a 998,562-vertex ridged ellipsoid -> 10k clusters (configs[2])",
    "C3s": "1/16-scale C3",
    "C5": "ACVD -m 1 -l 3 on a 160M-vertex (after SplitLongEdges) banded thin torus, 8:1 elongated cells -> 1.6M clusters (BASELINE configs[4])",
    "C5s": "1/256-scale C5",
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def config_dict(name, w, world):
    """The `config` object of the JSON line: identical in both arms (ours / --impl reference).  For workloads with a
    SplitLongEdges pre-pass V and F are those of the mesh after the split (what the clustering sees)."""
    V, F, K = int(w["points"].shape[0]), int(w["triangles"].shape[0]), int(w["K"])
    return {"workload": f"{name}: {WORKLOADS[name]}", "V": V, "F": F, "K": K,
            "metric_kind": w["metric"], "gradation": w["gradation"],
            "l2": "inputs larger than L2 (items+CSR >> 126 MB)" if V >= 2000000 else "inputs fit L2 (small workload)",
            "parallelism": f"vertex-range x{world}" if world > 1 else "single GPU"}


def make_workload(name):
    from acvd_b200 import meshgen
    w = meshgen.workload(name)
    w["name"] = name
    w["unconstrained_init"] = 1 if w["metric"] == "qem" else 0   # ACVDQ.cxx:325
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profiled_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture for this workload (profiles/kernel_traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(workload, {}).get(kernel)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------
def cpu_mesh(w):
    """The mesh the clustering sees on the CPU legs: workloads with `-l` go through the restated SplitLongEdges first."""
    if w.get("split_ratio"):
        from oracle import oracle
        p, t, _, _, _ = oracle.split_long_edges(w["points"], w["triangles"], w["split_ratio"])
        w["points"], w["triangles"], w["split_ratio"] = p, t, None
    return w["points"], w["triangles"]


def cpu_sample(w, threads, loops, steps=1, warmup=0):
    """Bounded sample of the restated reference (oracle) on the same workload: `loops` passes of
    ProcessOneLoop from the initial sampling.  Returns per-step (tests, seconds) and setup info."""
    from oracle import oracle
    t0 = time.time()
    o = oracle.Oracle(*cpu_mesh(w))
    o.build_metric(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    o.set_num_clusters(w["K"])
    cl0 = o.initial_sampling().copy()
    setup_s = time.time() - t0
    res = []
    for it in range(warmup + steps):
        o.set_num_clusters(w["K"])        # resets loops / counters / queue
        o.set_clustering(cl0)
        o.set_params(unconstrained_init=w["unconstrained_init"])
        t0 = time.perf_counter()
        if threads > 1:
            o.minimize_threaded(threads, loops)
        else:
            o.minimize(loops)
        dt = time.perf_counter() - t0
        r = o.report()
        if it >= warmup:
            res.append((r["tests"], dt))
    return res, setup_s


def cpu_to_convergence(w, threads=1):
    """The restated reference run TO CONVERGENCE on workload dict `w` (timer where the reference's own sits,
    Common/vtkUniformClustering.h:690-706).  Returns seconds, loops, tests, energy (fresh statistics)."""
    from oracle import oracle
    o = oracle.Oracle(*cpu_mesh(w))
    o.build_metric(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    o.set_num_clusters(w["K"])
    o.initial_sampling()
    o.set_params(unconstrained_init=w["unconstrained_init"])
    t0 = time.perf_counter()
    if threads > 1:
        o.minimize_threaded(threads, 0)
    else:
        o.minimize(0)
    if w.get("force_manifold"):        # the -m 1 loop (vtkDiscreteRemeshing.h:937-950), sequential engine
        for _ in range(200):
            if o.detect_non_manifold(1).size == 0:
                break
            o.set_connexity(0)
            o.minimize(0)
    dt = time.perf_counter() - t0
    r = o.report()
    o.recompute_statistics()
    return {"seconds": dt, "loops": r["loops"], "tests": r["tests"], "convergences": r["convergences"],
            "energy": o.global_energy(), "threads": threads}


# workloads whose sequential CPU run to convergence fits the bench's time bound (seconds on one core: C1 0.4, C2 ~40)
CPU_CONVERGES_LIVE = {"C1", "C2", "C4s", "C2s", "C3s", "C5s"}
SCALED_TWIN = {"C4": ("C4s", 100.0), "C3": ("C3s", 16.0), "C5": ("C5s", 256.0)}


def cpu_time_to_convergence(name, w, rate_full, threads=1):
    """CPU time to convergence for workload `name`: measured live when it fits the bound, otherwise measured on the
    scaled twin of the workload (same generator, same V/K ratio) and extrapolated -- formula in `how`."""
    kind = "sequential" if threads == 1 else f"threaded ({threads} threads)"
    if name in CPU_CONVERGES_LIVE:
        r = cpu_to_convergence(w, threads)
        r.update(extrapolated=False, how=f"{kind} restated reference on the full {name} mesh, run to convergence")
        r["time_to_convergence_s"] = r["seconds"]
        return r
    twin, factor = SCALED_TWIN[name]
    wt = make_workload(twin)
    res, _ = cpu_sample(wt, threads, 3)
    rate_twin = res[0][0] / res[0][1]
    r = cpu_to_convergence(wt, threads)
    vf = w["points"].shape[0] / wt["points"].shape[0]
    slow = rate_twin / rate_full if rate_full else 1.0
    r.update(extrapolated=True, twin=twin, twin_seconds=r["seconds"], twin_V=int(wt["points"].shape[0]),
             how=(f"{kind} restated reference run to convergence on {twin} (1/{factor:g}-scale twin: {r['seconds']:.2f} s, "
                  f"{r['loops']} loops, {r['tests']} tests), scaled by the vertex ratio {vf:.1f} and by the measured drop of the "
                  f"tests/s rate from {twin} to the full mesh over the first 3 loops ({rate_twin:.3g} -> {rate_full:.3g} tests/s)"))
    r["time_to_convergence_s"] = r["seconds"] * vf * slow
    r["tests"] = int(r["tests"] * vf)
    return r


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (restated in oracle/, the
    upstream sources need VTK and cannot be compiled here) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload)
    cpu_mesh(w)
    threads = os.cpu_count() or 1
    loops = args.ref_loops
    res, setup_s = cpu_sample(w, threads, loops, steps=args.steps, warmup=args.warmup)
    tests = sum(r[0] for r in res)
    secs = sum(r[1] for r in res)
    val = tests / secs
    sample = (f"{loops} ProcessOneLoop passes from the initial sampling per step (first, unconstrained phase) of the "
              f"threaded restatement (vtkThreadedClustering scheme, {threads} threads) on the full {args.workload} mesh")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(res)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, w, int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": setup_s,
    }
    conv = cpu_time_to_convergence(args.workload, w, val, threads)
    line["time_to_convergence_s"] = conv["time_to_convergence_s"]
    line["cpu_baseline"].update(time_to_convergence_s=conv["time_to_convergence_s"], time_extrapolated=conv["extrapolated"],
                                loops=conv["loops"], tests_to_convergence=conv["tests"], energy=conv["energy"], how=conv["how"])
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def main():
    # stdout carries exactly one JSON line: anything libraries print there (e.g. NCCL's version banner) goes to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ACVD_BENCH_WORKLOAD", "C4"), choices=list(WORKLOADS))
    ap.add_argument("--ref-loops", type=int, default=3, help="ProcessOneLoop passes per reference step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        log("note: timing rules ask for >= 3 warm-up steps")

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the acvd_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from acvd_b200 import capi
    t0 = time.time()
    w = make_workload(args.workload)
    V, F, K = int(w["points"].shape[0]), int(w["triangles"].shape[0]), int(w["K"])
    log(f"[rank {rank}] workload {args.workload}: V={V} F={F} K={K} generated in {time.time()-t0:.1f}s")

    # pinned host copies of the inputs (e2e copies come from pinned memory)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    _k1, h_xyz = pinned(w["points"])
    _k2, h_tri = pinned(w["triangles"])
    h_ind = h_pd = None
    if w["indicator"] is not None:
        _k3, h_ind = pinned(w["indicator"])
    if w.get("pd") is not None:
        _k6, h_pd = pinned(w["pd"])

    ctx = capi.Context(local_rank)
    if world > 1:
        uid = [capi.Context.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.dist_init(rank, world, uid[0])
    split_s = None
    if w.get("split_ratio"):
        # -l: vtkSurface::SplitLongEdges on the device; the clustering (and V in the config) sees the mesh after the split
        t0 = time.time()
        ctx.set_mesh(h_xyz, h_tri)
        ratio = w["split_ratio"]
        nv2, nf2, n_pass = ctx.split_long_edges(ratio, fetch=False)
        del _k1, _k2, h_xyz, h_tri
        w["points"] = w["triangles"] = None
        _k1 = torch.empty((nv2, 3), dtype=torch.float32).pin_memory()
        _k2 = torch.empty((nf2, 3), dtype=torch.int32).pin_memory()
        h_xyz, h_tri = _k1.numpy(), _k2.numpy()
        ctx._ck(ctx.L.acvd_get_subdivision(ctx.h, capi._p(h_xyz), capi._p(h_tri), None, None))
        w["points"], w["triangles"] = h_xyz, h_tri
        w["split_ratio"] = None           # done: the CPU baseline legs below get the mesh the clustering sees
        V, F = nv2, nf2
        split_s = time.time() - t0
        log(f"[rank {rank}] SplitLongEdges({ratio}): {n_pass} passes -> V={V} F={F} in {split_s:.2f}s (upload, split, download)")
    fm = bool(w.get("force_manifold"))
    t0 = time.time()
    ctx.set_mesh(h_xyz, h_tri)
    t1 = time.time()
    ctx.build_items(w["metric"], w["gradation"], h_ind, h_pd)
    ctx.set_num_clusters(K)
    t2 = time.time()
    # host, sequential; outside the timed region as in the reference (:690).  Across GPUs rank 0 runs it and broadcasts
    # ("run once on host or GPU 0 and scatter")
    if world > 1:
        cl_t = torch.empty(V, dtype=torch.int32, device="cuda")
        if rank == 0:
            ctx.initial_sampling()
            cl_t.copy_(torch.from_numpy(ctx.clustering()))
        dist.broadcast(cl_t, src=0)
        if rank != 0:
            ctx.set_clustering(cl_t.cpu().numpy())
        del cl_t
    else:
        ctx.initial_sampling()
    t3 = time.time()
    ctx.save_clustering()
    _k4, h_cl0 = pinned(ctx.clustering())
    setup = {"split_long_edges_s": split_s, "set_mesh_s": t1 - t0, "build_items_s": t2 - t1, "initial_sampling_s": t3 - t2,
             "note": "first calls of the process (memory pool growth included); initial sampling is sequential by definition "
                     "(vtkUniformClustering.h:1178-1316) and outside the reference's own timer (:690)"}
    log(f"[rank {rank}] setup: set_mesh {t1-t0:.2f}s, items {t2-t1:.2f}s, initial sampling {t3-t2:.2f}s")
    mparams = dict(unconstrained_init=w["unconstrained_init"])

    SUM_KEYS = None

    def run_clustering():
        """MinimizeEnergy, and under -m 1 the loop of vtkDiscreteRemeshing.h:937-950: detect non-manifold output
        vertices, re-enter MinimizeEnergy with the connexity constraint off, until the dual mesh is manifold."""
        rep = ctx.minimize(**mparams)
        if not fm:
            return rep
        rep["m_loop_iterations"] = 0
        rep["ms_detect"] = 0.0
        for _ in range(200):
            t0 = time.perf_counter()
            n_issues = ctx.detect_non_manifold(1)
            dt = 1e3 * (time.perf_counter() - t0)
            rep["ms_detect"] += dt
            rep["ms_device"] += dt          # wall time of the detection step (device kernels + the host list logic)
            if n_issues == 0:
                break
            r2 = ctx.minimize(connexity=0, **mparams)
            rep["m_loop_iterations"] += 1
            for k, v in r2.items():
                if k in ("energy", "disconnected"):
                    rep[k] = v
                else:
                    rep[k] += v
        rep["output_vertices"] = ctx.K
        return rep

    def step():
        if fm:
            ctx.set_num_clusters(K)     # the -m loop of the previous step appended clusters
        ctx.restore_clustering()
        return run_clustering()

    for _ in range(args.warmup):
        step()
    barrier()
    reps = []
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            reps.append(step())
        barrier()
        t_wall = time.perf_counter() - t_wall0
    ms_dev = sum(r["ms_device"] for r in reps)
    lr = reps[-1]
    import hashlib
    final_sha = hashlib.sha256(ctx.clustering().tobytes()).hexdigest()[:16]     # equal at every N: same clustering
    log(f"[rank {rank}] last step: device {lr['ms_device']:.1f} ms = scan {lr['ms_scan']:.1f} (dense bulk {lr['ms_dense_scan']:.1f} in "
        f"{lr['dense_scan_launches']} launches) + evaluate {lr['ms_evaluate']:.1f} + commit {lr['ms_commit']:.1f} + stats/clean/fill {lr['ms_clean']:.1f} "
        f"+ other {lr['ms_device'] - lr['ms_scan'] - lr['ms_evaluate'] - lr['ms_commit'] - lr['ms_clean']:.1f}; rounds {lr['rounds']} "
        f"({lr['bulk_rounds']} bulk), launches {lr['kernel_launches']}")
    if dist is not None:
        tt = torch.tensor([ms_dev], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev = float(tt.item())
    tests = sum(r["tests"] for r in reps)
    value = tests / (ms_dev * 1e-3)
    ms_per_step = ms_dev / args.steps

    # ---- end to end through the C ABI from pinned host buffers (fresh context state every step)
    e2e_tests = 0
    _k5, h_out = pinned(np.zeros(V, dtype=np.int32))
    # result buffers of the caller, pinned and reused (as a C++ caller's would be); under -m 1 the loop may append clusters
    np_pay_out = {"iso": 4, "qem": 13, "aniso": 13, "anisoq": 22}[w["metric"]]
    stats_out = None
    if not fm:
        _k7 = [pinned(np.zeros((K, np_pay_out))), pinned(np.zeros((K, 3))), pinned(np.zeros(K)), pinned(np.zeros(K, dtype=np.int32))]
        stats_out = tuple(a for _, a in _k7)
    # one untimed pass first (when e2e is measured at all): the first pass after the device-resident steps re-grows
    # the memory pool for the mesh-build scratch, a one-off of the process, not of the path
    e2e_passes = args.e2e_steps + (1 if args.e2e_steps else 0)
    barrier()
    t_e2e0 = time.perf_counter()
    for e2e_it in range(e2e_passes):
        if e2e_it == 1:
            barrier()
            t_e2e0 = time.perf_counter()
            e2e_tests = 0
        tt = [time.perf_counter()]
        ctx.set_mesh(h_xyz, h_tri); tt.append(time.perf_counter())
        ctx.build_items(w["metric"], w["gradation"], h_ind, h_pd); tt.append(time.perf_counter())
        ctx.set_num_clusters(K); tt.append(time.perf_counter())
        ctx.set_clustering(h_cl0); tt.append(time.perf_counter())
        r = run_clustering(); tt.append(time.perf_counter())

        ctx.clustering(h_out); tt.append(time.perf_counter())
        sums, cen, en, sz = ctx.cluster_stats(stats_out); tt.append(time.perf_counter())
        e2e_tests += r["tests"]
        log(f"[rank {rank}] e2e step: " + ", ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in zip(
            ("set_mesh", "build_items", "set_num_clusters", "set_clustering", "minimize", "get_clustering", "get_cluster_stats"), tt, tt[1:])))
    barrier()
    t_e2e = time.perf_counter() - t_e2e0
    h2d = h_xyz.nbytes + h_tri.nbytes + h_cl0.nbytes + (h_ind.nbytes if h_ind is not None else 0) + (h_pd.nbytes if h_pd is not None else 0)
    np_pay = {"iso": 4, "qem": 13, "aniso": 13, "anisoq": 22}[w["metric"]]
    d2h = h_out.nbytes + K * (np_pay * 8 + 24 + 8 + 4)
    e2e = {"value": e2e_tests / t_e2e if args.e2e_steps else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "s_per_step": t_e2e / max(1, args.e2e_steps), "steps": args.e2e_steps}

    # ---- roofline of the dominant kernel of the reassignment loop, timed with CUDA events inside the library on the
    # stream the kernels are launched on.  The dominant kernel is the TMA-staged dense bulk scan (k_scan_bulk_dense);
    # the other frontier-scan launches (list-based k_scan) and the candidate evaluation (k_evaluate) are given beside it.
    # Counters are summed over the ranks while every rank times its own share, hence the division by `world`.
    peak, peak_src = measured_peak()
    n_l = sum(r["round_launches"] for r in reps)

    def kernel_roof(name, label, by, ms, launches):
        by = by / world
        ach = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"bound": "hbm", "kernel": label, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": profiled_traffic(args.workload, name) if world == 1 else None, "peak_source": peak_src,
                "bytes_per_launch": by / max(1, launches), "us_per_launch": 1e3 * ms / max(1, launches),
                "launches_per_step": launches / args.steps, "share_of_step": ms / ms_dev}
    S = lambda k: sum(r[k] for r in reps)
    roof_dense = kernel_roof("k_scan_bulk_dense", "dense bulk scan = k_scan_classify + k_bulk_decide (TMA-staged frontier scan over all tiles -> candidate list -> bulk decision)",
                             S("dense_scan_bytes"), S("ms_dense_scan"), S("dense_scan_launches"))
    if S("dense_scan_launches"):
        # roofline.achieved / frac count the HBM-compulsory bytes only: SURVEY 8d's frontier-scan figure, 8 + 8 deg = 56 B per
        # vertex (row_ptr / own cluster id / neighbour ids / neighbour cluster ids), times the vertices one launch scans.
        # The operands of the tests (centroids, sizes: gathers the kernel serves from L1/L2) are reported beside it.
        with_ops = dict(GBps=roof_dense["achieved"], frac=roof_dense["frac"], bytes_per_launch=roof_dense["bytes_per_launch"],
                        note="scan bytes + the tests' operands (per decided vertex 48 B, per test 32 B, per proposal 8 B): L1/L2-resident gathers, not HBM traffic")
        scan_b = S("dense_scan_vertices") / world * 56.0
        ach = scan_b / (S("ms_dense_scan") * 1e-3) / 1e9
        roof_dense.update(achieved=ach, frac=ach / peak, bytes_per_launch=scan_b / S("dense_scan_launches"),
                          bytes_model=f"56 B per vertex (SURVEY 8d frontier scan: 8 + 8 deg, deg = 6) x {int(S('dense_scan_vertices') / world / S('dense_scan_launches'))} vertices per launch",
                          with_test_operands=with_ops)
    roof_scan = kernel_roof("k_scan_avg", "k_scan (all frontier-scan launches of the loop, dense + list-based)", S("scan_bytes"), S("ms_scan"), n_l)
    # (the ncu DRAM figures of these two are per full-activity launch and do not pair with per-launch averages over a
    # whole run, so no `traffic` is attached to them; profiles/README.md has the full-launch captures)
    roof_eval = kernel_roof("k_evaluate_avg", "k_evaluate (candidate energy evaluation of the exact rounds)", S("evaluate_bytes"), S("ms_evaluate"),
                            n_l - S("bulk_rounds"))
    roofline = roof_dense if S("dense_scan_launches") else roof_scan
    roofline["other_kernels"] = [{k: o[k] for k in ("kernel", "achieved", "frac", "bytes_per_launch", "us_per_launch", "launches_per_step", "share_of_step", "traffic")}
                                 for o in (roof_scan, roof_eval)]

    # ---- CPU baseline beside it (rank 0, N=1 only): sequential restated reference, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        loops = args.ref_loops
        res, setup_s = cpu_sample(w, 1, loops)
        rate = res[0][0] / res[0][1]
        conv = cpu_time_to_convergence(args.workload, w, rate)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{loops} sequential ProcessOneLoop passes from the initial sampling on the full {args.workload} mesh "
                         f"({res[0][0]} tests in {res[0][1]:.1f}s; first phase) for the tests/s rate; time to convergence: {conv['how']}",
               "time_to_convergence_s": conv["time_to_convergence_s"], "time_extrapolated": conv["extrapolated"],
               "loops": conv["loops"], "tests_to_convergence": conv["tests"], "energy": conv["energy"],
               "host_threads_available": os.cpu_count()}

    if rank == 0:
        last = reps[-1]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.workload, w, world),
            "time_to_convergence_s": ms_per_step * 1e-3, "wall_s_per_step": t_wall / args.steps,
            "rounds": last["rounds"], "convergences": last["convergences"], "modifications": last["modifications"],
            "energy": last["energy"], "tests_per_step": tests / args.steps, "clustering_sha256_16": final_sha,
            "m_loop": ({"iterations": last.get("m_loop_iterations"), "output_vertices": last.get("output_vertices"),
                        "ms_detect": last.get("ms_detect")} if fm else None),
            "vs_cpu_time_ratio": (cpu["time_to_convergence_s"] / (ms_per_step * 1e-3)) if cpu else None, "setup": setup,
            "tail": {"ms_sparse_rounds": sum(r.get("ms_sparse", 0.0) for r in reps) / args.steps,
                     "sparse_rounds": sum(r.get("sparse_rounds", 0) for r in reps) / args.steps,
                     "share_of_step": sum(r.get("ms_sparse", 0.0) for r in reps) / ms_dev},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(sum(r["kernel_launches"] for r in reps)),
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
