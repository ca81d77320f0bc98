"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer stages bit-exact given an identical cluster assignment;
per-cluster sums and quadrics within 1e-6 relative; converged energy within 1 % of the sequential
reference restatement.
"""
import numpy as np
import pytest

from acvd_b200 import meshgen

pytestmark = pytest.mark.gpu

REL = 1e-6   # sums / quadrics tolerance stated by north_star


def rel_err(a, b):
    scale = np.maximum(np.abs(b).max(axis=0, keepdims=True), 1e-300)
    return float((np.abs(a - b) / scale).max())


@pytest.fixture(scope="module")
def sphere():
    return meshgen.geodesic_icosphere(32)   # V = 10 242


@pytest.fixture(scope="module")
def torus():
    p, t = meshgen.torus_grid(160, 100, noise=0.002, seed=1)
    return p, t, meshgen.torus_curvature_indicator(160, 100)


@pytest.fixture(scope="module")
def spindle():
    return meshgen.bipyramid(24, 4)      # V = 6 146, two vertices of valence 24 (> ELL width)


def make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K, gradation=0.0, cw=None, pd=None):
    o = oracle_mod.Oracle(p, t)
    o.build_metric(metric, gradation, cw, pd)
    o.set_num_clusters(K)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items(metric, gradation, cw, pd)
    g.set_num_clusters(K)
    return o, g


def test_csr_matches_oracle_adjacency(oracle_mod, gpu_ctx_factory, sphere):
    p, t = sphere
    o = oracle_mod.Oracle(p, t)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    assert g.num_edges() == o.E
    rp_o, col_o = o.csr()
    rp_g, col_g = g.csr()
    assert np.array_equal(rp_o, rp_g)          # same degrees
    for v in (0, 1, 11, 12, 500, p.shape[0] - 1):
        assert sorted(col_o[rp_o[v]:rp_o[v + 1]]) == list(col_g[rp_g[v]:rp_g[v + 1]])
    # rows sorted, whole neighbour sets equal
    so = np.concatenate([np.sort(col_o[rp_o[v]:rp_o[v + 1]]) for v in range(p.shape[0])])
    assert np.array_equal(so, col_g)


@pytest.mark.parametrize("metric", ["iso", "qem"])
def test_items_match_oracle(oracle_mod, gpu_ctx_factory, torus, metric):
    p, t, ind = torus
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, 50, 1.5, ind)
    io, ig = o.items(), g.items()
    assert io.shape == ig.shape
    assert rel_err(ig, io) < REL
    assert rel_err(g.vertex_areas()[:, None], o.vertex_areas()[:, None]) < 1e-12


def test_items_match_oracle_aniso(oracle_mod, gpu_ctx_factory, sphere):
    p, t = sphere
    rng = np.random.default_rng(3)
    pd = rng.normal(size=(p.shape[0], 6)).astype(np.float32)
    ind = rng.uniform(0.5, 2.0, size=p.shape[0])
    for metric in ("aniso", "anisoq"):
        o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, 50, 1.5, ind, pd)
        assert rel_err(g.items(), o.items()) < REL


@pytest.mark.parametrize("metric", ["iso", "qem", "aniso", "anisoq"])
def test_recompute_statistics_matches_oracle(oracle_mod, gpu_ctx_factory, sphere, metric):
    p, t = sphere
    K = 200
    rng = np.random.default_rng(5)
    pd = rng.normal(size=(p.shape[0], 6)).astype(np.float32) if metric.startswith("aniso") else None
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K, 0.0, None, pd)
    g.set_items(metric, o.items())            # identical item bytes
    cl = o.initial_sampling()
    g.set_clustering(cl)
    o.recompute_statistics()
    g.recompute_statistics(1, 3)
    so, co, eo, zo = o.cluster_stats()
    sg, cg, eg, zg = g.cluster_stats()
    assert np.array_equal(zo, zg)             # sizes exact
    assert rel_err(sg, so) < REL
    assert np.abs(cg - co).max() < 1e-6 * max(1.0, np.abs(co).max())
    assert np.abs(eg - eo).max() <= REL * np.abs(eo).max()
    assert abs(g.global_energy() - o.global_energy()) <= REL * abs(o.global_energy())


def test_representative_points_match_oracle(oracle_mod, gpu_ctx_factory):
    rng = np.random.default_rng(7)
    n = 4096
    Q = np.zeros((n, 9))
    # sums of 1..3 plane quadrics: full rank, rank-deficient (plane, crease) and noisy cases
    for i in range(n):
        for _ in range(1 + i % 3):
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            d = rng.normal()
            v = np.append(nrm, d)
            q = np.outer(v, v)
            Q[i] += [q[0, 0], q[0, 1], q[0, 2], q[0, 3], q[1, 1], q[1, 2], q[1, 3], q[2, 2], q[2, 3]]
    P = rng.normal(size=(n, 3))
    g = gpu_ctx_factory()
    for level in (3, 2, 1):
        pg, rg = g.representative_points(Q, P, level, 1e-3)
        for i in range(n):
            po, ro = oracle_mod.representative_point(Q[i], P[i], level, 1e-3)
            assert ro == rg[i]
            assert np.abs(po - pg[i]).max() < 1e-8 * max(1.0, np.abs(po).max())


def test_integer_stages_bit_exact(oracle_mod, gpu_ctx_factory, sphere):
    p, t = sphere
    K = 300
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    o.initial_sampling()
    o.minimize()
    cl = o.clustering()
    g.set_clustering(cl)
    assert np.array_equal(g.boundary_flags(), o.boundary_flags())
    assert np.array_equal(g.cluster_adjacency(), o.cluster_adjacency())
    assert np.array_equal(g.dual_triangles(), o.dual_triangles())


def test_clean_clustering_bit_exact(oracle_mod, gpu_ctx_factory, sphere):
    p, t = sphere
    K = 100
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    cl = o.initial_sampling().copy()
    # break clusters apart: transplant scattered vertices into other clusters
    rng = np.random.default_rng(11)
    idx = rng.choice(p.shape[0], 400, replace=False)
    cl[idx] = rng.integers(0, K, size=400)
    o.set_clustering(cl)
    g.set_clustering(cl)
    do = o.clean_clustering()
    dg = g.clean_clustering()
    assert do == dg and do > 0
    assert np.array_equal(o.clustering(), g.clustering())
    # fill: every vertex assigned afterwards, clusters connected after one more clean
    g.fill_holes()
    cg = g.clustering()
    assert cg.min() >= 0 and cg.max() < K


@pytest.mark.parametrize("metric,uncon", [("iso", 0), ("qem", 1), ("qem", 0)])
def test_minimize_energy_within_one_percent(oracle_mod, gpu_ctx_factory, sphere, metric, uncon):
    p, t = sphere
    K = 200
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K)
    g.set_items(metric, o.items())
    cl0 = o.initial_sampling()
    g.set_clustering(cl0)
    o.set_params(unconstrained_init=uncon)
    o.minimize()
    o.recompute_statistics()
    rep = g.minimize(unconstrained_init=uncon, log_energy=1)
    cg = g.clustering()
    # invariants of a converged clustering
    assert cg.min() >= 0 and cg.max() < K
    sizes = np.bincount(cg, minlength=K)
    assert sizes.min() >= 1 and sizes.sum() == p.shape[0]
    assert g.clean_clustering() == 0          # every cluster connected
    log = g.energy_log()
    assert rep["modifications"] > 0 and rep["rounds"] == len(log)
    when = g.energy_times()
    assert len(when) == len(log) and np.all(np.diff(when) >= 0) and when[-1] <= 1e-3 * rep["ms_total"] + 1e-3
    # energy: raw (energy.txt parity number) and translation-invariant sum w |p - c|^2
    it = o.items()
    w = it[:, 3]
    const = float(np.sum((it[:, :3] ** 2).sum(axis=1) / w))
    e_o, e_g = o.global_energy(), rep["energy"]
    assert abs(e_g - e_o) <= 0.01 * abs(e_o)
    if metric == "iso":
        t_o, t_g = const + e_o, const + e_g
        assert t_g <= 1.01 * t_o
    # the oracle finds nothing to improve on the GPU result and vice versa
    o2 = oracle_mod.Oracle(p, t)
    o2.build_metric(metric)
    o2.set_num_clusters(K)
    o2.set_clustering(cg)
    o2.set_connexity(1)
    o2.prime()
    assert o2.process_one_loop() == 0


def test_energy_monotone_per_phase(oracle_mod, gpu_ctx_factory, torus):
    p, t, ind = torus
    K = 400
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K, 1.5, ind)
    cl0 = o.initial_sampling()
    g.set_clustering(cl0)
    g.recompute_statistics()
    e_prev = g.global_energy()
    for _ in range(30):
        r = g.reassign_round(1, 3, 0)
        e = g.global_energy()
        assert e <= e_prev + 1e-12 * abs(e_prev)
        assert (r["modifications"] > 0) == (r["proposals"] > 0)
        e_prev = e


@pytest.mark.parametrize("metric,uncon", [("iso", 0), ("qem", 1)])
def test_bulk_rounds_energy_monotone_and_same_fixed_point(oracle_mod, gpu_ctx_factory, torus, metric, uncon):
    """Bulk (Lloyd-criterion) rounds only ever lower the energy, and with or without them the
    minimisation ends in a state the sequential oracle cannot improve, at the same energy within 1 %."""
    p, t, ind = torus
    K = 500
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K, 1.5, ind)
    cl0 = o.initial_sampling()
    res = {}
    for bulk in (1000, -1):          # forced on (the automatic setting turns them on from 500 k vertices) / off
        g.set_clustering(cl0)
        rep = g.minimize(unconstrained_init=uncon, log_energy=1, bulk_rounds=bulk)
        log = g.energy_log()
        nb = rep["bulk_rounds"]
        assert (nb > 0) == (bulk > 0)
        if nb:
            # energy never rises from one bulk round to the next; the only places it may rise are the phase
            # boundaries, where CleanClustering / FillHoles re-assign broken-off components
            e = log[:nb]
            rises = np.diff(e) > 1e-12 * np.abs(e[:-1])
            assert rises.sum() <= rep["convergences"] - 1
            assert np.all(np.diff(e)[rises] <= 1e-5 * np.abs(e[:-1][rises]))
        res[bulk] = rep["energy"]
        assert g.clean_clustering() == 0
    assert abs(res[1000] - res[-1]) <= 0.01 * abs(res[-1])
    o.set_params(unconstrained_init=uncon)
    o.minimize()
    o.recompute_statistics()
    assert abs(res[1000] - o.global_energy()) <= 0.01 * abs(o.global_energy())


@pytest.mark.parametrize("metric,uncon", [("iso", 0), ("qem", 1)])
def test_high_valence_rows(oracle_mod, gpu_ctx_factory, spindle, metric, uncon):
    """Adjacency rows longer than the ELL width take the CSR overflow paths (scan, bulk decision, components)."""
    p, t = spindle
    K = 150
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K)
    rp, col = g.csr()
    assert (np.diff(rp) > 8).sum() == 2 and np.diff(rp).max() == 24
    rpo, colo = o.csr()
    assert np.array_equal(rp, rpo) and np.array_equal(col, np.concatenate([np.sort(colo[rpo[v]:rpo[v + 1]]) for v in range(p.shape[0])]))
    g.set_items(metric, o.items())
    cl0 = o.initial_sampling().copy()
    # bit-exact stages at the same clustering, including the clean-up of broken clusters around the apexes
    cl = cl0.copy()
    cl[cl == K] = 0
    apex = int(np.argmax(np.diff(rp)))
    cl[col[rp[apex]:rp[apex + 1]][::2]] = (cl[apex] + 1) % K     # every other neighbour of the apex -> foreign cluster
    o.set_clustering(cl)
    g.set_clustering(cl)
    assert np.array_equal(g.boundary_flags(), o.boundary_flags())
    assert o.clean_clustering() == g.clean_clustering()
    assert np.array_equal(o.clustering(), g.clustering())
    # full minimisation from the same start
    o.set_clustering(cl0)
    g.set_clustering(cl0)
    o.set_params(unconstrained_init=uncon)
    o.minimize()
    o.recompute_statistics()
    rep = g.minimize(unconstrained_init=uncon)
    cg = g.clustering()
    assert cg.min() >= 0 and cg.max() < K and np.bincount(cg, minlength=K).min() >= 1
    assert g.clean_clustering() == 0
    assert abs(rep["energy"] - o.global_energy()) <= 0.01 * abs(o.global_energy())
    o2 = oracle_mod.Oracle(p, t)
    o2.build_metric(metric)
    o2.set_num_clusters(K)
    o2.set_clustering(cg)
    o2.set_connexity(1)
    o2.prime()
    assert o2.process_one_loop() == 0


def test_round_trip_determinism(gpu_ctx_factory, sphere):
    p, t = sphere
    res = []
    for _ in range(2):
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        g.build_items("qem")
        g.set_num_clusters(150)
        g.initial_sampling()
        g.minimize(unconstrained_init=1)
        res.append(g.clustering())
    assert np.array_equal(res[0], res[1])


def test_initial_sampling_matches_oracle(oracle_mod, gpu_ctx_factory, sphere):
    p, t = sphere
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", 123)
    g.set_items("iso", o.items())
    g.initial_sampling()
    assert np.array_equal(g.clustering(), o.initial_sampling())


@pytest.mark.parametrize("mesh", ["torus", "spindle"])
def test_tma_staged_dense_scan_equals_list_scan(gpu_ctx_factory, torus, spindle, mesh, monkeypatch):
    """The TMA-staged dense bulk scan (every generation: fused k_scan_bulk_dense*, split k_scan_classify + k_bulk_decide) and
    the list-based k_scan<W, true> take the same decisions: identical clustering, tests, proposals and rounds on the same start."""
    if mesh == "torus":
        p, t, ind = torus
        K, grad = 400, 1.5
    else:
        (p, t), ind, K, grad = spindle, None, 150, 0.0
    res = []
    # variants: 0 = fused first generation; "1"/"0" = the list-based scan; 20, 22, 25 = third generation (min/max candidates,
    # prefetch, static assignment); 40+ = split form (classify + decide) with 1 / 2 / 4 tiles per ticket -- 42 is the shipped default
    for no_dense, variant in (("", "0"), ("1", "0"), ("", "20"), ("", "22"), ("", "25"), ("", "40"), ("", "42"), ("", "46")):
        monkeypatch.setenv("ACVD_DENSE_VARIANT", variant)
        if no_dense:
            monkeypatch.setenv("ACVD_NO_DENSE_SCAN", "1")
        else:
            monkeypatch.delenv("ACVD_NO_DENSE_SCAN", raising=False)
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        g.build_items("qem", grad, ind)
        g.set_num_clusters(K)
        g.initial_sampling()
        rep = g.minimize(unconstrained_init=1, bulk_rounds=1000)      # forced on: these meshes are below the automatic threshold
        res.append((g.clustering().copy(), rep))
    c0, r0 = res[0]
    assert r0["bulk_rounds"] > 0 and r0["dense_scan_launches"] > 0 and res[1][1]["dense_scan_launches"] == 0
    for c1, r1 in res[1:]:
        for k in ("rounds", "bulk_rounds", "tests", "proposals", "modifications", "evaluated"):
            assert r0[k] == r1[k], k
        assert np.array_equal(c0, c1)
        assert r0["energy"] == r1["energy"]


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_full_size_properties(gpu_ctx_factory, name):
    """BASELINE-size configurations (C2: ACVDQ gradation 1.5, 2.6 M vertices -> 100 k clusters; C3: AnisotropicRemeshingQ
    1.5, 1 M vertices -> 10 k; C4: ACVDQ, 40 M vertices -> 400 k, the bench workload) through size-independent
    properties: every cluster non-empty and connected, sizes sum to V, energy below the initial one, a further round
    finds no improving move (idempotence), two runs identical (C4: one run; scripts/c5_check.py does the same at 160 M)."""
    w = meshgen.workload(name)
    p, t, K = w["points"], w["triangles"], int(w["K"])
    uncon = 1 if w["metric"] == "qem" else 0
    runs = []
    for _ in range(1 if name == "C4" else 2):
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
        g.set_num_clusters(K)
        g.initial_sampling()
        g.fill_holes()
        g.recompute_statistics(0 if uncon else 1, 3)
        e0 = g.global_energy()
        rep = g.minimize(unconstrained_init=uncon)
        cl = g.clustering()
        runs.append((cl.copy(), rep))
        assert cl.min() >= 0 and cl.max() < K
        sz = np.bincount(cl, minlength=K)
        assert sz.min() >= 1 and sz.sum() == p.shape[0]
        _, _, _, zg = g.cluster_stats()
        assert np.array_equal(zg, sz)
        assert rep["disconnected"] == 0 and g.clean_clustering() == 0
        if not uncon:
            assert rep["energy"] < e0            # same energy functional before and after
        again = g.reassign_round(1, 3, 1)
        assert again["proposals"] == 0 and again["modifications"] == 0
        g.close()
    if len(runs) == 2:
        assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1]["energy"] == runs[1][1]["energy"]


def test_bench_kernel_hook(gpu_ctx_factory, torus):
    """acvd_bench_kernel times the dense bulk scan variants on the current state and leaves the context usable."""
    p, t, ind = torus
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items("qem", 1.5, ind)
    g.set_num_clusters(400)
    g.initial_sampling()
    g.save_clustering()
    g.minimize(unconstrained_init=1, max_loops=4)
    for variant in (-1, 0, 1):
        assert g.bench_kernel(0, variant, 0, 3) > 0.0
    g.restore_clustering()
    rep = g.minimize(unconstrained_init=1)
    cl = g.clustering()
    assert rep["disconnected"] == 0 and np.bincount(cl, minlength=400).min() >= 1


@pytest.mark.parametrize("metric", ["aniso", "anisoq"])
def test_minimize_anisotropic_within_one_percent(oracle_mod, gpu_ctx_factory, metric):
    """The anisotropic metrics (vtkAnisotropicMetricForClustering / vtkQuadricAnisotropicMetricForClustering) through the
    whole minimisation against the sequential oracle, on a small ridged ellipsoid with analytic principal directions."""
    p, t = meshgen.ridged_ellipsoid(32)
    pd, ind = meshgen.ellipsoid_principal_directions(p)
    K = 150
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, metric, K, 1.5, ind, pd)
    g.set_items(metric, o.items())
    cl0 = o.initial_sampling()
    g.set_clustering(cl0)
    o.minimize()
    o.recompute_statistics()
    rep = g.minimize()
    cg = g.clustering()
    sizes = np.bincount(cg, minlength=K)
    assert cg.min() >= 0 and cg.max() < K and sizes.min() >= 1 and sizes.sum() == p.shape[0]
    assert g.clean_clustering() == 0
    e_o, e_g = o.global_energy(), rep["energy"]
    assert abs(e_g - e_o) <= 0.01 * abs(e_o)
    o2 = oracle_mod.Oracle(p, t)
    o2.build_metric(metric, 1.5, ind, pd)
    o2.set_num_clusters(K)
    o2.set_clustering(cg)
    o2.set_connexity(1)
    o2.prime()
    assert o2.process_one_loop() == 0


@pytest.mark.parametrize("mesh", ["torus", "ellipsoid", "fan"])
def test_curvature_matches_oracle(oracle_mod, gpu_ctx_factory, torus, mesh):
    """acvd_curvature (vtkCurvatureMeasure: polynomial fitting over the 3-ring, vertices) against the oracle's restatement:
    indicator within 1e-6 relative, principal-direction vectors within 1e-5 of their scale where the two curvatures are
    separated (where they coincide the directions are not defined).  "fan" has a vertex of valence 120, whose
    neighbourhood exceeds the kernel's local list (global-scratch pass)."""
    if mesh == "torus":
        p, t, _ = torus
    elif mesh == "ellipsoid":
        p, t = meshgen.ridged_ellipsoid(24)
    else:
        p, t = meshgen.bipyramid(120, 2)
    o = oracle_mod.Oracle(p, t)
    io, fo = o.curvature(3)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    ig, fg = g.curvature(3)
    scale = np.median(io[io > 0])
    assert np.array_equal(io == 0, ig == 0)          # same degenerate / small-neighbourhood vertices
    assert np.abs(ig - io).max() <= REL * max(scale, np.abs(io).max())
    ka = (fo[:, :3].astype(np.float64) ** 2).sum(axis=1)
    kb = (fo[:, 3:].astype(np.float64) ** 2).sum(axis=1)
    sep = (io > 0) & (np.abs(ka - kb) > 1e-3 * (ka + kb))
    assert sep.mean() > 0.5
    fs = np.abs(fo).max()
    for blk in (slice(0, 3), slice(3, 6)):
        d = np.minimum(np.abs(fg[sep][:, blk] - fo[sep][:, blk]).max(axis=1), np.abs(fg[sep][:, blk] + fo[sep][:, blk]).max(axis=1))
        assert d.max() <= 1e-5 * fs
    # indicator only (no principal directions requested)
    ig2, none = g.curvature(3, principal_directions=False)
    assert none is None and np.array_equal(ig2, ig)


def test_curvature_feeds_gradation_run(gpu_ctx_factory, torus):
    """The measured indicator drives a gradation-1.5 ACVDQ run (the reference's SamplingPreProcessing path): clusters
    shrink where the curvature is high (weights w = area * indicator^1.5 are equalised, vtkQEMetricForClustering.h:333-334),
    which gradation 0 does not do."""
    p, t, _ = torus
    res = {}
    for grad in (0.0, 1.5):
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        ind, _ = g.curvature(3, principal_directions=False)
        g.build_items("qem", grad, ind if grad > 0 else None)
        g.set_num_clusters(400)
        g.initial_sampling()
        rep = g.minimize(unconstrained_init=1)
        assert rep["disconnected"] == 0
        cl = g.clustering()
        size = np.bincount(cl, minlength=400).astype(np.float64)
        mean_ind = np.bincount(cl, weights=ind, minlength=400) / size
        hi = mean_ind > np.median(mean_ind)
        res[grad] = size[hi].mean() / size[~hi].mean()        # size ratio: high-curvature clusters over low-curvature ones
    assert res[1.5] < 0.9 * res[0.0], res


@pytest.mark.parametrize("mesh", ["sphere", "spindle", "torus"])
def test_subdivide_bit_exact(oracle_mod, gpu_ctx_factory, sphere, spindle, torus, mesh):
    """acvd_subdivide (vtkSurface::Subdivide, Common/vtkSurface.cxx:605-677) against the numbering the reference produces:
    old points, then one midpoint per edge in the oracle's (first-seen) edge order, four faces per face; parents = edge
    end points in first-seen orientation.  Bit-exact, and the result is a valid closed mesh with 4 F faces."""
    p, t = {"sphere": sphere, "spindle": spindle, "torus": torus[:2]}[mesh]
    o = oracle_mod.Oracle(p, t)
    a, b = o.edges()
    V, E = p.shape[0], a.shape[0]
    mid = (0.5 * (p[a].astype(np.float64) + p[b].astype(np.float64))).astype(np.float32)
    key = np.minimum(a, b).astype(np.int64) * V + np.maximum(a, b)
    order = np.argsort(key)

    def eid(u, v):
        k = np.minimum(u, v).astype(np.int64) * V + np.maximum(u, v)
        return order[np.searchsorted(key[order], k)]
    v4, v5, v6 = V + eid(t[:, 0], t[:, 1]), V + eid(t[:, 1], t[:, 2]), V + eid(t[:, 2], t[:, 0])
    exp_t = np.stack([t[:, 0], v4, v6, v4, t[:, 1], v5, v5, t[:, 2], v6, v4, v5, v6], axis=1).reshape(-1, 3).astype(np.int32)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    ps, ts, p1, p2 = g.subdivide()
    assert ps.shape == (V + E, 3) and ts.shape == (4 * t.shape[0], 3)
    assert np.array_equal(ps[:V], p) and np.array_equal(ps[V:], mid)
    assert np.array_equal(ts, exp_t)
    assert np.array_equal(p1[:V], np.arange(V)) and np.array_equal(p2[:V], np.arange(V))
    assert np.array_equal(p1[V:], a) and np.array_equal(p2[V:], b)
    # the subdivided mesh is a mesh: feed it back
    g2 = gpu_ctx_factory()
    g2.set_mesh(ps, ts)
    assert g2.num_edges() == 2 * E + 3 * t.shape[0]


def test_against_committed_golden_fixtures(gpu_ctx_factory):
    """The CUDA path against the fixtures committed under tests/golden/ (frozen oracle outputs, generator alongside):
    initial sampling, dual triangles and energy at the frozen clustering (ico6, K = 12); curvature and the subdivision's
    edge numbering (ridged ellipsoid, 162 vertices)."""
    import json
    import os
    gold_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(gold_dir, "oracle_ico6_k12.json")))
    p, t = meshgen.geodesic_icosphere(6)
    for metric in ("iso", "qem"):
        gm = gold[metric]
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        g.build_items(metric)
        g.set_num_clusters(12)
        g.initial_sampling()
        assert g.clustering().tolist() == gm["initial_sampling"]
        g.set_clustering(np.asarray(gm["clustering"], dtype=np.int32))
        assert g.dual_triangles().tolist() == gm["dual_triangles"]
        g.recompute_statistics(1, 3)
        assert abs(g.global_energy() - gm["energy"]) <= REL * abs(gm["energy"])
    gold = json.load(open(os.path.join(gold_dir, "oracle_curv_edges_ico4.json")))
    p, t = meshgen.ridged_ellipsoid(4)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    ind, info = g.curvature(3)
    gi = np.asarray(gold["indicator"])
    assert np.abs(ind - gi).max() <= REL * np.abs(gi).max()
    ps, ts, p1, p2 = g.subdivide()
    V = p.shape[0]
    assert p1[V:].tolist() == gold["edge_v1"] and p2[V:].tolist() == gold["edge_v2"]


# ---------------------------------------------------------------------------------------------------------------
# round 2: pins at the BASELINE sizes, the connexity predicate itself, FillHoles in the reference's FIFO order

def _scrambled_clustering(o, p, K, seed, n_scatter):
    cl = o.initial_sampling().copy()
    o.fill_holes()
    cl = o.clustering().copy()
    rng = np.random.default_rng(seed)
    idx = rng.choice(p.shape[0], n_scatter, replace=False)
    cl[idx] = rng.integers(0, K, size=n_scatter)
    return cl


@pytest.mark.parametrize("mesh", ["sphere", "spindle"])
def test_connexity_predicate_matches_oracle(oracle_mod, gpu_ctx_factory, sphere, spindle, mesh):
    """vtkVerticesProcessing::ConnexityConstraintProblemLocal (DiscreteRemeshing/vtkVerticesProcessing.h:168-237):
    the device predicate -- ring bit matrix for rows <= 8, generic walk for longer rows -- against the oracle's
    restatement on (item, cluster) pairs: every vertex with its own cluster and with the clusters of its neighbours,
    on a clustering with many fragmented clusters; the spindle has two rows of length 24."""
    p, t = sphere if mesh == "sphere" else spindle
    K = 60
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    cl = _scrambled_clustering(o, p, K, 5, p.shape[0] // 6)
    o.set_clustering(cl)
    g.set_clustering(cl)
    o.set_connexity(1)
    rp, col = g.csr()
    V = p.shape[0]
    items = np.concatenate([np.arange(V, dtype=np.int32), np.repeat(np.arange(V, dtype=np.int32), np.diff(rp))])
    clusters = np.concatenate([cl, cl[col]]).astype(np.int32)
    want = np.array([o.connexity_problem(int(i), int(c)) for i, c in zip(items, clusters)], dtype=np.uint8)
    assert want.sum() > 50 and (want == 0).sum() > 50          # both outcomes are exercised
    for mode in (0, 1):
        got = g.connexity_problem(items, clusters, mode)
        assert np.array_equal(got, want), (mode, int((got != want).sum()))
    long_rows = np.flatnonzero(np.diff(rp) > 8)
    if mesh == "spindle":
        assert long_rows.size == 2


def _punch_holes(p, cl, K, seed, n_holes, radius):
    """NULL (id K) patches several rings deep."""
    rng = np.random.default_rng(seed)
    cl = cl.copy()
    centres = p[rng.choice(p.shape[0], n_holes, replace=False)]
    for c in centres:
        cl[np.linalg.norm(p - c, axis=1) < radius] = K
    return cl


@pytest.mark.parametrize("mesh", ["sphere", "torus", "spindle"])
@pytest.mark.parametrize("connexity", [0, 1])
def test_fill_holes_bit_exact(oracle_mod, gpu_ctx_factory, sphere, torus, spindle, mesh, connexity):
    """FillHolesInClustering (Common/vtkUniformClustering.h:552-633) in the reference's FIFO order, including the
    ConnexityConstraintProblem guard of :606-607: same clustering as the oracle's sequential queue, bit for bit --
    (a) the NULL vertices the initial sampling leaves, (b) holes several rings deep, (c) scattered single holes."""
    p, t = {"sphere": sphere, "torus": torus[:2], "spindle": spindle}[mesh]
    K = 80
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    cl0 = o.initial_sampling().copy()
    cases = [cl0]
    o.fill_holes()
    full = o.clustering().copy()
    cases.append(_punch_holes(p, full, K, 3, 12, 0.12))
    sc = full.copy()
    sc[np.random.default_rng(4).choice(p.shape[0], 300, replace=False)] = K
    cases.append(sc)
    cases.append(_punch_holes(p, sc, K, 6, 3, 0.3))
    n_null = 0
    for cl in cases:
        o.set_connexity(connexity)
        o.set_clustering(cl)
        g.set_clustering(cl)
        o.fill_holes()
        g.fill_holes(connexity)
        co, cg = o.clustering(), g.clustering()
        n_null += int((cl == K).sum())
        assert np.array_equal(co, cg), int((co != cg).sum())
        if not connexity:
            assert cg.max() < K
    assert n_null > 500


def test_negative_ids_are_null(gpu_ctx_factory, sphere):
    """Ids outside [0, K) are "not assigned" (vtkUniformClustering.h:560-561): normalised to the NULL id K on upload."""
    p, t = sphere
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items("iso")
    g.set_num_clusters(50)
    g.initial_sampling()
    g.fill_holes()
    cl = g.clustering()
    cl2 = cl.copy()
    cl2[::7] = -1
    cl2[3::11] = 50 + 17
    g.set_clustering(cl2)
    back = g.clustering()
    assert np.array_equal(back == 50, (cl2 < 0) | (cl2 >= 50))
    g.fill_holes()
    assert g.clustering().max() < 50 and g.clustering().min() >= 0


def test_context_reuse_across_meshes_of_different_scale(oracle_mod, gpu_ctx_factory, sphere):
    """One context, two meshes whose coordinates differ by 1e3: the bulk rounds' fixed-point scale follows the items
    (it used to be cached for the lifetime of the context), and a new mesh needs a new cluster count."""
    p, t = sphere
    g = gpu_ctx_factory()
    energies = []
    for scale in (1.0, 1000.0, 0.01):
        ps = (p * scale).astype(np.float32)
        g.set_mesh(ps, t)
        with pytest.raises(Exception):
            g.clean_clustering()                     # the old clustering died with the old mesh
        g.build_items("iso")
        g.set_num_clusters(150)
        g.initial_sampling()
        rep = g.minimize(bulk_rounds=1000)
        assert rep["bulk_rounds"] > 0 and rep["disconnected"] == 0
        o = oracle_mod.Oracle(ps, t)
        o.build_metric("iso")
        o.set_num_clusters(150)
        o.initial_sampling()
        o.minimize()
        o.recompute_statistics()
        assert abs(rep["energy"] - o.global_energy()) <= 0.01 * abs(o.global_energy())
        energies.append(rep["energy"])
    assert abs(energies[1]) > 1e9 * abs(energies[0]) > 1e9 * abs(energies[2])    # E ~ scale^4


def test_set_mesh_rejects_bad_indices(gpu_ctx_factory, sphere):
    p, t = sphere
    g = gpu_ctx_factory()
    bad = t.copy()
    bad[5, 1] = p.shape[0] + 3
    with pytest.raises(Exception):
        g.set_mesh(p, bad)
    bad[5, 1] = -2
    with pytest.raises(Exception):
        g.set_mesh(p, bad)
    g.set_mesh(p, t)                                  # the context is still usable


def _baseline_fixture():
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_baseline_runs.json")
    return json.load(open(path))


@pytest.mark.parametrize("name", ["C1", "C2", "C3"])
def test_baseline_size_energy_vs_oracle(oracle_mod, gpu_ctx_factory, name):
    """BASELINE.json configs[0..2] at full size: the converged GPU energy is within 1 % of the sequential oracle's
    (tests/golden/oracle_baseline_runs.json, produced by tests/golden/make_baseline_runs.py from the same generator and
    the same initial sampling -- its sha256 is checked; C1 is also re-run live), and the oracle's ProcessOneLoop, primed
    on the GPU's final clustering with the connexity constraint on, finds no improving move."""
    import hashlib
    fx = _baseline_fixture()[name]
    w = meshgen.workload(name)
    p, t, K = w["points"], w["triangles"], int(w["K"])
    assert (p.shape[0], K) == (fx["V"], fx["K"])
    uncon = fx["unconstrained_init"]
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    g.set_num_clusters(K)
    g.initial_sampling()
    cl0 = g.clustering()
    assert hashlib.sha256(cl0.tobytes()).hexdigest() == fx["sha256_initial_sampling"]
    rep = g.minimize(unconstrained_init=uncon)
    cg = g.clustering()
    e_o = fx["energy"]
    assert abs(rep["energy"] - e_o) <= 0.01 * abs(e_o), (rep["energy"], e_o)
    o = oracle_mod.Oracle(p, t)
    o.build_metric(w["metric"], w["gradation"], w["indicator"], w.get("pd"))
    o.set_num_clusters(K)
    if name == "C1":                                  # live re-run of the oracle (0.3 s): the fixture is not stale
        o.set_clustering(cl0)
        o.set_params(unconstrained_init=uncon)
        o.minimize()
        o.recompute_statistics()
        assert abs(o.global_energy() - e_o) <= 1e-12 * abs(e_o)
        assert o.report()["tests"] == fx["tests"] and o.report()["loops"] == fx["loops"]
        o.set_num_clusters(K)
    if w["metric"] in ("iso", "qem"):                 # stricter: translation-invariant energy sum w |p - c|^2
        items = g.items()
        _, cen, _, sz = g.cluster_stats()
        assert sz.min() >= 1
        te = oracle_mod.true_energy(p, items[:, 3], cg, cen)
        assert te <= 1.01 * fx["true_energy"], (te, fx["true_energy"])
    o.set_clustering(cg)
    o.set_connexity(1)
    o.prime()
    assert o.process_one_loop() == 0
    g.close()


@pytest.mark.parametrize("case", ["torus-qem", "sphere-iso", "spindle-qem", "ellipsoid-anisoq", "sphere-qem-passes4"])
def test_sparse_rounds_equal_tile_filter_rounds(gpu_ctx_factory, sphere, torus, spindle, case, monkeypatch):
    """The persistent sparse-round kernel (dirty set enumerated from the member arrays of the modified clusters,
    convergence decided on the device) takes the decisions of the tile-filter path round for round: same clustering,
    same number of rounds, vertex tests, proposals and moves, same energy -- the "recently modified" rule of
    Common/vtkUniformClustering.h:909-920 is preserved."""
    kw = {}
    if case == "torus-qem":
        p, t, ind = torus
        args, K, uncon = ("qem", 1.5, ind, None), 500, 1
    elif case == "sphere-iso":
        p, t = sphere
        args, K, uncon = ("iso", 0.0, None, None), 300, 0
    elif case == "spindle-qem":
        p, t = spindle
        args, K, uncon = ("qem", 0.0, None, None), 150, 1
    elif case == "sphere-qem-passes4":
        p, t = sphere
        args, K, uncon = ("qem", 0.0, None, None), 250, 1
        kw = dict(commit_passes=4)
    else:
        p, t = meshgen.ridged_ellipsoid(40)
        pd, ind = meshgen.ellipsoid_principal_directions(p)
        args, K, uncon = ("anisoq", 1.5, ind, pd), 200, 0
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items(*args)
    g.set_num_clusters(K)
    g.initial_sampling()
    g.save_clustering()
    out = {}
    # 0: sparse rounds, launch shape chosen by the work (grid / one cluster); "grid": cluster form off; -1: tile-filter path
    for sparse in (0, "grid", -1):
        if sparse == "grid":
            monkeypatch.setenv("ACVD_NO_SPARSE_CLUSTER", "1")
        else:
            monkeypatch.delenv("ACVD_NO_SPARSE_CLUSTER", raising=False)
        g.restore_clustering()
        rep = g.minimize(unconstrained_init=uncon, sparse_rounds=-1 if sparse == -1 else 0, **kw)
        out[sparse] = (g.clustering().copy(), rep)
    (c0, r0), (cg, rg), (c1, r1) = out[0], out["grid"], out[-1]
    assert r0["sparse_rounds"] > 0 and rg["sparse_rounds"] > 0 and r1["sparse_rounds"] == 0
    assert rg["sparse_cluster_rounds"] == 0
    if case in ("torus-qem", "spindle-qem"):
        assert r0["sparse_cluster_rounds"] > 0      # these meshes are small: the tail runs in the one-cluster form
    assert np.array_equal(c0, c1) and np.array_equal(cg, c1)
    for k in ("rounds", "tests", "proposals", "modifications", "convergences", "evaluated", "energy"):
        assert r0[k] == r1[k] == rg[k], (k, r0[k], rg[k], r1[k])
    assert g.clean_clustering() == 0


def test_connectivity_check_of_modified_clusters_only(gpu_ctx_factory, sphere, torus, monkeypatch):
    """CleanClustering inside acvd_minimize checks the connectivity of the clusters modified since the previous check only
    (the others were connected then and have not changed).  Same clustering, events and counts as checking every cluster at
    every event (ACVD_CC_ALL=1), from a start with disconnected clusters (a random labelling) and from the usual sampling."""
    p, t, ind = torus
    rng = np.random.default_rng(5)
    for (pp, tt), K, start in (((p, t), 400, "sampling"), (sphere, 60, "random")):
        res = []
        for cc_all in ("", "1"):
            if cc_all:
                monkeypatch.setenv("ACVD_CC_ALL", "1")
            else:
                monkeypatch.delenv("ACVD_CC_ALL", raising=False)
            g = gpu_ctx_factory()
            g.set_mesh(pp, tt)
            g.build_items("qem")
            g.set_num_clusters(K)
            if start == "sampling":
                g.initial_sampling()
            else:
                g.set_clustering(np.random.default_rng(5).integers(0, K, pp.shape[0]).astype(np.int32))
            rep = g.minimize(unconstrained_init=1)
            res.append((g.clustering().copy(), rep, g.clean_clustering()))
        (c0, r0, d0), (c1, r1, d1) = res
        assert np.array_equal(c0, c1) and d0 == d1 == 0
        for k in ("rounds", "convergences", "tests", "modifications", "disconnected", "energy"):
            assert r0[k] == r1[k], (start, k, r0[k], r1[k])


def test_cluster_stats_into_caller_buffers(gpu_ctx_factory, sphere):
    p, t = sphere
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    g.build_items("qem")
    g.set_num_clusters(50)
    g.initial_sampling()
    g.minimize(unconstrained_init=1, max_loops=5)
    a = g.cluster_stats()
    out = (np.full((50, 13), -1.0), np.full((50, 3), -1.0), np.full(50, -1.0), np.full(50, -1, dtype=np.int32))
    b = g.cluster_stats(out)
    assert all(x is y for x, y in zip(b, out))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


# ---------------------------------------------------------------------------------------------------------------
# the -m 1 row: IsVertexManifold on the device, one DetectNonManifoldOutputVertices step, the ACVD post-process sums

def _pinched_open_mesh(sphere):
    """sphere with a few faces removed (boundary edges) and two distant vertices merged (a pinch)"""
    p, t = sphere
    t = np.delete(t, [5, 400, 401, 2000], axis=0)
    a, b = int(t[0, 0]), int(t[-1, 0])
    t = t.copy()
    t[t == b] = a
    return p, t


def test_input_manifold_flags_match_oracle(oracle_mod, gpu_ctx_factory, sphere, spindle):
    for p, t in (sphere, spindle, _pinched_open_mesh(sphere)):
        o = oracle_mod.Oracle(p, t)
        g = gpu_ctx_factory()
        g.set_mesh(p, t)
        want, got = o.input_vertex_manifold(), g.input_manifold_flags()
        assert np.array_equal(want, got), int((want != got).sum())
    assert (want == 0).sum() >= 8


@pytest.mark.parametrize("force", [0, 1])
def test_output_manifold_flags_match_oracle(oracle_mod, gpu_ctx_factory, sphere, force):
    """vtkSurfaceBase::IsVertexManifold on the dual mesh (with and without the -m edges of vtkDiscreteRemeshing.h:1114-1133):
    a converged clustering, a fragmented one, and caps + band (no dual face at all)."""
    p, t = sphere
    K = 120
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    cl0 = o.initial_sampling().copy()
    o.minimize()
    cases = [o.clustering().copy(), _scrambled_clustering(o, p, K, 9, 600)]
    n_bad = 0
    for cl in cases:
        o.set_clustering(cl)
        g.set_clustering(cl)
        want, got = o.output_vertex_manifold(force), g.output_manifold_flags(force)
        assert np.array_equal(want, got), int((want != got).sum())
        n_bad += int((want == 0).sum())
    assert cases[0] is not None and n_bad > 5
    band = np.where(p[:, 2] > 0.4, 0, np.where(p[:, 2] < -0.4, 2, 1)).astype(np.int32)
    o3, g3 = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", 3)
    o3.set_clustering(band)
    g3.set_clustering(band)
    assert np.array_equal(o3.output_vertex_manifold(force), g3.output_manifold_flags(force))
    assert not g3.output_manifold_flags(force).any()


def test_detect_non_manifold_step_matches_oracle(oracle_mod, gpu_ctx_factory, sphere):
    """One DetectNonManifoldOutputVertices step (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383) from the same
    clustering: same issues, same grown cluster count, same edited clustering, same frozen flags -- then the -m loop
    (re-enter MinimizeEnergy with the connexity constraint off, detect again) ends on a manifold dual mesh."""
    p, t = sphere
    K = 150
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    o.initial_sampling()
    cl = _scrambled_clustering(o, p, K, 21, 500)
    # include a one-item cluster among the offenders (exercises the "take a ring neighbour" branch)
    lone = int(np.flatnonzero(np.bincount(cl, minlength=K) > 3)[0])
    keep = int(np.flatnonzero(cl == lone)[0])
    rp, col = g.csr()
    other = cl[col[rp[keep]]]
    cl[(cl == lone) & (np.arange(p.shape[0]) != keep)] = other if other != lone else (lone + 1) % K
    o.set_clustering(cl)
    g.set_clustering(cl)
    issues_o = o.detect_non_manifold(1)
    n_g = g.detect_non_manifold(1)
    assert n_g == issues_o.size and n_g > 3
    assert g.K == o.K and g.K > K
    assert np.array_equal(g.clustering(), o.clustering())
    assert np.array_equal(g.frozen(), o.frozen())
    # the loop of vtkDiscreteRemeshing.h:937-950 on the device side
    for _ in range(40):
        g.minimize(connexity=0)
        if g.detect_non_manifold(1) == 0:
            break
    else:
        raise AssertionError("-m loop did not end")
    assert g.output_manifold_flags(1).all()
    cg = g.clustering()
    assert cg.min() >= 0 and cg.max() < g.K and np.bincount(cg, minlength=g.K).min() >= 1
    n_tri = g.dual_triangles().shape[0]
    assert n_tri == 2 * g.K - 4                     # closed genus-0 manifold output


def test_cluster_quadrics_match_item_sums(oracle_mod, gpu_ctx_factory, sphere):
    """ACVD's quadric post-process (Examples/ACVD.cxx:237-262) sums, per cluster, the quadrics of the faces around every
    item: that is the sum of the QEM item quadrics (vtkQEMetricForClustering.h:151-167) over the cluster's items."""
    p, t = sphere
    K = 90
    o, g = make_pair(oracle_mod, gpu_ctx_factory, p, t, "iso", K)
    cl = o.initial_sampling().copy()
    o.fill_holes()
    cl = o.clustering().copy()
    g.set_clustering(cl)
    q = g.cluster_quadrics()
    oq = oracle_mod.Oracle(p, t)
    oq.build_metric("qem")
    items = oq.items()[:, 4:13]
    want = np.zeros((K, 9))
    np.add.at(want, cl, items)
    assert rel_err(q, want) <= REL
    assert np.array_equal(g.cluster_quadrics(10), q[:10])


@pytest.mark.parametrize("mesh", ["banded-torus", "sphere", "spindle"])
def test_split_long_edges_bit_exact(oracle_mod, gpu_ctx_factory, sphere, spindle, mesh):
    """vtkSurface::SplitLongEdges (Common/vtkSurface.cxx:444-604, option -l) on the device against the restated pass
    structure (threshold from the mesh as given; per pass: midpoints in edge-id order, Split2 / Split3 / 1 -> 4 patterns):
    points, faces and parents identical; behaviour: no edge above the threshold is left, the surface stays closed and
    manifold, old points are untouched."""
    if mesh == "banded-torus":
        w = meshgen.workload("C5s")
        p, t, ratio = w["points"], w["triangles"], w["split_ratio"]
    elif mesh == "sphere":
        (p, t), ratio = sphere, 0.9             # cuts most edges, several passes
    else:
        (p, t), ratio = spindle, 1.2
    po, to, p1o, p2o, passes_o = oracle_mod.split_long_edges(p, t, ratio)
    g = gpu_ctx_factory()
    g.set_mesh(p, t)
    pg, tg, p1g, p2g, passes_g = g.split_long_edges(ratio)
    assert passes_o == passes_g and passes_g >= 1
    assert pg.shape == po.shape and tg.shape == to.shape and pg.shape[0] > p.shape[0]
    assert np.array_equal(pg, po) and np.array_equal(tg, to)
    assert np.array_equal(p1g, p1o) and np.array_equal(p2g, p2o)
    assert np.array_equal(pg[:p.shape[0]], p)
    e0 = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    ue = np.unique(np.sort(e0, axis=1), axis=0)
    thr = ratio * np.linalg.norm(p[ue[:, 0]].astype(np.float64) - p[ue[:, 1]].astype(np.float64), axis=1).mean()
    e = np.concatenate([tg[:, [0, 1]], tg[:, [1, 2]], tg[:, [2, 0]]])
    L = np.linalg.norm(pg[e[:, 0]].astype(np.float64) - pg[e[:, 1]].astype(np.float64), axis=1)
    assert L.max() <= thr * (1 + 1e-12)
    key = np.minimum(e[:, 0], e[:, 1]).astype(np.int64) * pg.shape[0] + np.maximum(e[:, 0], e[:, 1])
    _, cnt = np.unique(key, return_counts=True)
    assert (cnt == 2).all()                       # closed, edge-manifold
    g2 = gpu_ctx_factory()
    g2.set_mesh(pg, tg)
    assert g2.input_manifold_flags().all()
