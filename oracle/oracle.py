"""ctypes wrapper over the CPU oracle (oracle/acvd_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(acvd_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ISO, QEM, ANISO, ANISOQ = 0, 1, 2, 3
METRICS = {"iso": ISO, "qem": QEM, "aniso": ANISO, "anisoq": ANISOQ}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libacvd_oracle.so")
    src = os.path.join(_HERE, "acvd_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libacvd_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i, i, vp, vp]
        for name, args, res in [
            ("orc_destroy", [vp], None), ("orc_num_edges", [vp], i), ("orc_get_edges", [vp, vp, vp], None),
            ("orc_get_csr", [vp, vp, vp], None), ("orc_vertex_areas", [vp, vp], None),
            ("orc_build_metric", [vp, i, d, vp, vp], None), ("orc_payload_size", [vp], i),
            ("orc_get_items", [vp, vp], None), ("orc_set_num_clusters", [vp, i], None),
            ("orc_set_params", [vp, i, i, i, i, i, i], None), ("orc_set_constrained", [vp, i], None),
            ("orc_set_fixed", [vp, vp, i], None), ("orc_set_frozen", [vp, vp], None),
            ("orc_initial_sampling", [vp], None), ("orc_set_clustering", [vp, vp], None),
            ("orc_get_clustering", [vp, vp], None), ("orc_minimize", [vp, i], None),
            ("orc_minimize_threaded", [vp, i, i], None), ("orc_process_one_loop", [vp], i),
            ("orc_prime", [vp], None), ("orc_recompute_statistics", [vp], None),
            ("orc_clean_clustering", [vp], i), ("orc_fill_holes", [vp], None),
            ("orc_set_connexity", [vp, i], None), ("orc_connexity_problem", [vp, i, i], i),
            ("orc_global_energy", [vp], d), ("orc_get_cluster_stats", [vp, vp, vp, vp, vp], None),
            ("orc_get_report", [vp, vp], None), ("orc_energy_log", [vp, vp, i], i),
            ("orc_representative_point", [vp, vp, i, d], i), ("orc_triangle_quadric", [vp, vp, vp, vp], None),
            ("orc_triangle_area", [vp, vp, vp], d), ("orc_mt19937_first", [vp, i], None),
            ("orc_dual_triangles", [vp, vp, i], i), ("orc_boundary_flags", [vp, vp], None),
            ("orc_cluster_adjacency", [vp, vp, C.c_int64], C.c_int64),
            ("orc_curvature", [vp, i, vp, vp], None),
            ("orc_split_long_edges", [i, i, vp, vp, d, vp, vp, vp, vp, vp, vp], i),
            ("orc_output_vertex_manifold", [vp, i, vp], None), ("orc_input_vertex_manifold", [vp, vp], None),
            ("orc_detect_non_manifold", [vp, i, vp, i], i), ("orc_num_clusters", [vp], i), ("orc_get_frozen", [vp, vp], None),
        ]:
            f = getattr(L, name)
            f.argtypes = args
            f.restype = res
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def representative_point(Q9, P3, level=3, thr=1e-3):
    Q = np.ascontiguousarray(Q9, dtype=np.float64)
    P = np.array(P3, dtype=np.float64)
    rd = lib().orc_representative_point(_p(Q), _p(P), level, thr)
    return P, rd


def triangle_quadric(x1, x2, x3):
    a, b, c = (np.ascontiguousarray(x, dtype=np.float64) for x in (x1, x2, x3))
    q = np.zeros(10)
    lib().orc_triangle_quadric(_p(a), _p(b), _p(c), _p(q))
    return q


def mt19937_first(n=3):
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_mt19937_first(_p(out), n)
    return out


def split_long_edges(points, triangles, ratio):
    """vtkSurface::SplitLongEdges restated (Common/vtkSurface.cxx:444-604): (points, triangles, parent1, parent2, passes)."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    nv, nf = C.c_int(), C.c_int()
    passes = lib().orc_split_long_edges(p.shape[0], t.shape[0], _p(p), _p(t), float(ratio), C.byref(nv), C.byref(nf), None, None, None, None)
    po = np.zeros((nv.value, 3), dtype=np.float32)
    to = np.zeros((nf.value, 3), dtype=np.int32)
    p1 = np.zeros(nv.value, dtype=np.int32)
    p2 = np.zeros(nv.value, dtype=np.int32)
    lib().orc_split_long_edges(0, 0, None, None, 0.0, C.byref(nv), C.byref(nf), _p(po), _p(to), _p(p1), _p(p2))
    return po, to, p1, p2, passes


class Oracle:
    """Sequential (and threaded) restatement of vtkUniformClustering over a triangle mesh."""

    def __init__(self, points, triangles):
        self.points = np.ascontiguousarray(points, dtype=np.float32)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32)
        self.V, self.F = self.points.shape[0], self.triangles.shape[0]
        self.h = lib().orc_create(self.V, self.F, _p(self.points), _p(self.triangles))
        self.E = lib().orc_num_edges(self.h)
        self.K = 0
        self.np = 4

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    # mesh
    def edges(self):
        a = np.zeros(self.E, dtype=np.int32)
        b = np.zeros(self.E, dtype=np.int32)
        lib().orc_get_edges(self.h, _p(a), _p(b))
        return a, b

    def csr(self):
        rp = np.zeros(self.V + 1, dtype=np.int32)
        col = np.zeros(2 * self.E, dtype=np.int32)
        lib().orc_get_csr(self.h, _p(rp), _p(col))
        return rp, col

    def vertex_areas(self):
        out = np.zeros(self.V)
        lib().orc_vertex_areas(self.h, _p(out))
        return out

    def curvature(self, ring_size=3):
        """vtkCurvatureMeasure (polynomial fitting, vertices, n-ring): (indicator[V] float64, info[V, 6] float32)."""
        ind = np.zeros(self.V)
        info = np.zeros((self.V, 6), dtype=np.float32)
        lib().orc_curvature(self.h, int(ring_size), _p(ind), _p(info))
        return ind, info

    # metric
    def build_metric(self, metric="iso", gradation=0.0, custom_weights=None, principal_dirs=None):
        m = METRICS[metric] if isinstance(metric, str) else metric
        cw = None if custom_weights is None else np.ascontiguousarray(custom_weights, dtype=np.float64)
        pd = None if principal_dirs is None else np.ascontiguousarray(principal_dirs, dtype=np.float32)
        lib().orc_build_metric(self.h, m, float(gradation), _p(cw), _p(pd))
        self.np = lib().orc_payload_size(self.h)

    def items(self):
        out = np.zeros((self.V, self.np))
        lib().orc_get_items(self.h, _p(out))
        return out

    # engine
    def set_num_clusters(self, K):
        self.K = int(K)
        lib().orc_set_num_clusters(self.h, self.K)

    def set_params(self, unconstrained_init=0, qlevel=3, max_loops=0, max_conv=0, connexity=0, log_energy=0):
        lib().orc_set_params(self.h, unconstrained_init, qlevel, max_loops, max_conv, connexity, log_energy)

    def set_constrained(self, on):
        lib().orc_set_constrained(self.h, int(on))

    def set_fixed(self, items):
        a = np.ascontiguousarray(items, dtype=np.int64)
        lib().orc_set_fixed(self.h, _p(a), a.size)

    def set_frozen(self, flags):
        a = np.ascontiguousarray(flags, dtype=np.uint8)
        lib().orc_set_frozen(self.h, _p(a))

    def initial_sampling(self):
        lib().orc_initial_sampling(self.h)
        return self.clustering()

    def set_clustering(self, cl):
        a = np.ascontiguousarray(cl, dtype=np.int32)
        assert a.size == self.V
        lib().orc_set_clustering(self.h, _p(a))

    def clustering(self):
        out = np.zeros(self.V, dtype=np.int32)
        lib().orc_get_clustering(self.h, _p(out))
        return out

    def minimize(self, loop_budget=0):
        lib().orc_minimize(self.h, loop_budget)

    def minimize_threaded(self, threads, loop_budget=0):
        lib().orc_minimize_threaded(self.h, threads, loop_budget)

    def prime(self):
        lib().orc_prime(self.h)

    def process_one_loop(self):
        return lib().orc_process_one_loop(self.h)

    def recompute_statistics(self):
        lib().orc_recompute_statistics(self.h)

    def clean_clustering(self):
        return lib().orc_clean_clustering(self.h)

    def fill_holes(self):
        lib().orc_fill_holes(self.h)

    def set_connexity(self, on):
        lib().orc_set_connexity(self.h, int(on))

    def connexity_problem(self, item, cluster):
        return lib().orc_connexity_problem(self.h, int(item), int(cluster))

    def global_energy(self):
        return lib().orc_global_energy(self.h)

    def cluster_stats(self):
        sums = np.zeros((self.K, self.np))
        cen = np.zeros((self.K, 3))
        en = np.zeros(self.K)
        sz = np.zeros(self.K, dtype=np.int32)
        lib().orc_get_cluster_stats(self.h, _p(sums), _p(cen), _p(en), _p(sz))
        return sums, cen, en, sz

    def report(self):
        out = np.zeros(5)
        lib().orc_get_report(self.h, _p(out))
        return dict(loops=int(out[0]), convergences=int(out[1]), tests=int(out[2]), mods=int(out[3]), seconds=float(out[4]))

    def energy_log(self):
        n = lib().orc_energy_log(self.h, None, 0)
        out = np.zeros(n)
        lib().orc_energy_log(self.h, _p(out), n)
        return out

    def dual_triangles(self):
        cap = 4 * self.K + 64
        out = np.zeros((cap, 3), dtype=np.int32)
        n = lib().orc_dual_triangles(self.h, _p(out), cap)
        if n > cap:
            out = np.zeros((n, 3), dtype=np.int32)
            n = lib().orc_dual_triangles(self.h, _p(out), n)
        return out[:n].copy()

    def boundary_flags(self):
        out = np.zeros(self.V, dtype=np.uint8)
        lib().orc_boundary_flags(self.h, _p(out))
        return out

    # ---- the -m 1 loop (DiscreteRemeshing/vtkDiscreteRemeshing.h:166-383, Common/vtkSurfaceBase.cxx:259-317)
    def output_vertex_manifold(self, force_manifold=1):
        out = np.zeros(self.K, dtype=np.uint8)
        lib().orc_output_vertex_manifold(self.h, int(force_manifold), _p(out))
        return out

    def input_vertex_manifold(self):
        out = np.zeros(self.V, dtype=np.uint8)
        lib().orc_input_vertex_manifold(self.h, _p(out))
        return out

    def detect_non_manifold(self, force_manifold=1):
        """One DetectNonManifoldOutputVertices step: returns the offending cluster ids; K, clustering and the frozen
        flags of the oracle are updated (one new cluster per issue that found an item to take)."""
        cap = self.K + 1
        fl = np.zeros(cap, dtype=np.int32)
        n = lib().orc_detect_non_manifold(self.h, int(force_manifold), _p(fl), cap)
        self.K = lib().orc_num_clusters(self.h)
        return fl[:n].copy()

    def frozen(self):
        out = np.zeros(self.K, dtype=np.uint8)
        lib().orc_get_frozen(self.h, _p(out))
        return out

    def cluster_adjacency(self):
        n = lib().orc_cluster_adjacency(self.h, None, 0)
        out = np.zeros(n, dtype=np.int64)
        lib().orc_cluster_adjacency(self.h, _p(out), n)
        return np.stack([out >> 32, out & 0xFFFFFFFF], axis=1).astype(np.int32)


def true_energy(points, items_w, clustering, centroids):
    """Translation-invariant energy sum_i w_i |p_i - c|^2 (SURVEY §7 'Energy definition')."""
    p = points.astype(np.float64)
    d = p - centroids[clustering]
    return float(np.sum(items_w * np.einsum("ij,ij->i", d, d)))
