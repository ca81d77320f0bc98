"""ctypes binding of the C ABI in include/acvd_b200.h.

This is plumbing only: every call goes straight into ``libacvd_b200.so``.  There is no CPU
fallback — if the library is missing or no CUDA device is present the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacvd_b200.so")

ISO, QEM, ANISO, ANISOQ = 0, 1, 2, 3
METRICS = {"iso": ISO, "qem": QEM, "aniso": ANISO, "anisoq": ANISOQ}
NCCL_ID_BYTES = 128


class AcvdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"acvd_b200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [
        ("unconstrained_init", C.c_int32), ("quadrics_level", C.c_int32), ("connexity", C.c_int32),
        ("max_loops", C.c_int32), ("max_convergences", C.c_int32), ("early_stop_div", C.c_int32),
        ("log_energy", C.c_int32), ("rounds_per_sync", C.c_int32), ("sv_threshold", C.c_double),
        ("bulk_rounds", C.c_int32), ("commit_passes", C.c_int32), ("sparse_rounds", C.c_int32),
    ]


class Report(C.Structure):
    _fields_ = [
        ("rounds", C.c_int64), ("convergences", C.c_int64), ("tests", C.c_int64), ("modifications", C.c_int64),
        ("proposals", C.c_int64), ("disconnected", C.c_int64), ("energy", C.c_double), ("ms_total", C.c_double),
        ("ms_scan", C.c_double), ("ms_evaluate", C.c_double), ("ms_commit", C.c_double), ("ms_clean", C.c_double),
        ("round_launches", C.c_int64), ("scan_bytes", C.c_int64), ("evaluate_bytes", C.c_int64),
        ("evaluated", C.c_int64), ("ms_device", C.c_double),
        ("kernel_launches", C.c_int64), ("bulk_rounds", C.c_int64),
        ("dense_scan_launches", C.c_int64), ("ms_dense_scan", C.c_double), ("dense_scan_bytes", C.c_int64),
        ("dense_scan_vertices", C.c_int64), ("bulk_rollbacks", C.c_int64), ("sparse_rounds", C.c_int64),
        ("ms_sparse", C.c_double), ("sparse_cluster_rounds", C.c_int64),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/acvd_b200.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _d = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SYMBOLS = [
    ("acvd_payload_size", C.c_int, [C.c_int]),
    ("acvd_create", C.c_int, [C.POINTER(_vp), C.c_int]),
    ("acvd_destroy", C.c_int, [_vp]),
    ("acvd_last_error", C.c_char_p, [_vp]),
    ("acvd_abi_version", C.c_int, []),
    ("acvd_set_mesh", C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    ("acvd_get_num_edges", C.c_int, [_vp, C.POINTER(_i64)]),
    ("acvd_get_csr", C.c_int, [_vp, _vp, _vp]),
    ("acvd_subdivide", C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32)]),
    ("acvd_get_subdivision", C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ("acvd_split_long_edges", C.c_int, [_vp, _d, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    ("acvd_curvature", C.c_int, [_vp, _i32, _vp, _vp]),
    ("acvd_build_items", C.c_int, [_vp, C.c_int, _d, _vp, _vp]),
    ("acvd_set_items", C.c_int, [_vp, C.c_int, _vp]),
    ("acvd_get_items", C.c_int, [_vp, _vp]),
    ("acvd_get_vertex_areas", C.c_int, [_vp, _vp]),
    ("acvd_set_num_clusters", C.c_int, [_vp, _i32]),
    ("acvd_set_clustering", C.c_int, [_vp, _vp]),
    ("acvd_get_clustering", C.c_int, [_vp, _vp]),
    ("acvd_save_clustering", C.c_int, [_vp]),
    ("acvd_restore_clustering", C.c_int, [_vp]),
    ("acvd_set_frozen", C.c_int, [_vp, _vp]),
    ("acvd_set_fixed_clusters", C.c_int, [_vp, _vp, _i32]),
    ("acvd_initial_sampling", C.c_int, [_vp]),
    ("acvd_minimize", C.c_int, [_vp, C.POINTER(Params), C.POINTER(Report)]),
    ("acvd_recompute_statistics", C.c_int, [_vp, C.c_int, C.c_int]),
    ("acvd_clean_clustering", C.c_int, [_vp, C.POINTER(_i32)]),
    ("acvd_fill_holes", C.c_int, [_vp, C.c_int]),
    ("acvd_connexity_problem", C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    ("acvd_reassign_round", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    ("acvd_get_cluster_stats", C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ("acvd_global_energy", C.c_int, [_vp, C.POINTER(_d)]),
    ("acvd_get_energy_log", C.c_int, [_vp, _vp, _i32, C.POINTER(_i32)]),
    ("acvd_get_energy_times", C.c_int, [_vp, _vp, _i32, C.POINTER(_i32)]),
    ("acvd_representative_points", C.c_int, [_vp, _i32, _vp, _vp, _i32, _d, _vp]),
    ("acvd_cluster_quadrics", C.c_int, [_vp, _i32, _vp]),
    ("acvd_boundary_flags", C.c_int, [_vp, _vp]),
    ("acvd_cluster_adjacency", C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    ("acvd_dual_triangles", C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    ("acvd_input_manifold_flags", C.c_int, [_vp, _vp]),
    ("acvd_output_manifold_flags", C.c_int, [_vp, _i32, _vp]),
    ("acvd_detect_non_manifold", C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    ("acvd_get_frozen", C.c_int, [_vp, _vp]),
    ("acvd_bench_kernel", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    ("acvd_dist_unique_id", C.c_int, [_vp]),
    ("acvd_dist_init", C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    ("acvd_dist_partition", C.c_int, [_i64, _i64, _i32, _i32, _i32, _vp]),
]

_LIB = None


def load_library():
    """Load libacvd_b200.so (built in-tree by acvd_b200.build).  Raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise AcvdError(-2, f"{LIB_PATH} not built: run `python -m acvd_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dist_partition(V, F, K, rank, world):
    """acvd_dist_partition: the ranges rank works on -- dict(tiles, clusters, points, faces), each (begin, end).  Host
    arithmetic inside the library: needs no CUDA device."""
    out = np.zeros(8, dtype=np.int64)
    rc = load_library().acvd_dist_partition(int(V), int(F), int(K), int(rank), int(world), _p(out))
    if rc != 0:
        raise AcvdError(rc, "acvd_dist_partition: bad arguments")
    o = [int(x) for x in out]
    return dict(tiles=(o[0], o[1]), clusters=(o[2], o[3]), points=(o[4], o[5]), faces=(o[6], o[7]))


class Context:
    """One clustering context on one CUDA device (acvd_ctx)."""

    def __init__(self, device: int = -1):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.acvd_create(C.byref(h), device)
        if rc != 0:
            raise AcvdError(rc, self.L.acvd_last_error(None).decode())
        self.h = h
        self.V = self.F = self.K = 0
        self.metric = None

    def close(self):
        if getattr(self, "h", None):
            self.L.acvd_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise AcvdError(rc, self.L.acvd_last_error(self.h).decode())

    # ---- mesh / items
    def set_mesh(self, points, triangles):
        p = np.ascontiguousarray(points, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        self.V, self.F = p.shape[0], t.shape[0]
        self._ck(self.L.acvd_set_mesh(self.h, self.V, self.F, _p(p), _p(t)))

    def num_edges(self):
        e = C.c_int64()
        self._ck(self.L.acvd_get_num_edges(self.h, C.byref(e)))
        return e.value

    def csr(self):
        E = self.num_edges()
        rp = np.zeros(self.V + 1, dtype=np.int32)
        col = np.zeros(2 * E, dtype=np.int32)
        self._ck(self.L.acvd_get_csr(self.h, _p(rp), _p(col)))
        return rp, col

    def subdivide(self):
        """vtkSurface::Subdivide of the context's mesh: (points float32 [nv,3], triangles int32 [nf,3], parent1, parent2)."""
        nv, nf = _i32(), _i32()
        self._ck(self.L.acvd_subdivide(self.h, C.byref(nv), C.byref(nf)))
        p = np.zeros((nv.value, 3), dtype=np.float32)
        t = np.zeros((nf.value, 3), dtype=np.int32)
        p1 = np.zeros(nv.value, dtype=np.int32)
        p2 = np.zeros(nv.value, dtype=np.int32)
        self._ck(self.L.acvd_get_subdivision(self.h, _p(p), _p(t), _p(p1), _p(p2)))
        return p, t, p1, p2

    def split_long_edges(self, ratio, fetch=True):
        """vtkSurface::SplitLongEdges of the context's mesh: (points, triangles, parent1, parent2, passes), or the sizes."""
        nv, nf, npass = _i32(), _i32(), _i32()
        self._ck(self.L.acvd_split_long_edges(self.h, float(ratio), C.byref(nv), C.byref(nf), C.byref(npass)))
        if not fetch:
            return nv.value, nf.value, npass.value
        p = np.zeros((nv.value, 3), dtype=np.float32)
        t = np.zeros((nf.value, 3), dtype=np.int32)
        p1 = np.zeros(nv.value, dtype=np.int32)
        p2 = np.zeros(nv.value, dtype=np.int32)
        self._ck(self.L.acvd_get_subdivision(self.h, _p(p), _p(t), _p(p1), _p(p2)))
        return p, t, p1, p2, npass.value

    def curvature(self, ring_size=3, principal_directions=True):
        """vtkCurvatureMeasure (polynomial fitting, vertices, n-ring): (indicator[V] float64, info[V, 6] float32 or None)."""
        ind = np.zeros(self.V)
        info = np.zeros((self.V, 6), dtype=np.float32) if principal_directions else None
        self._ck(self.L.acvd_curvature(self.h, int(ring_size), _p(ind), _p(info)))
        return ind, info

    def build_items(self, metric="iso", gradation=0.0, custom_weights=None, principal_dirs=None):
        m = METRICS[metric] if isinstance(metric, str) else int(metric)
        cw = None if custom_weights is None else np.ascontiguousarray(custom_weights, dtype=np.float64)
        pd = None if principal_dirs is None else np.ascontiguousarray(principal_dirs, dtype=np.float32)
        self._ck(self.L.acvd_build_items(self.h, m, float(gradation), _p(cw), _p(pd)))
        self.metric = m

    def set_items(self, metric, payload):
        m = METRICS[metric] if isinstance(metric, str) else int(metric)
        a = np.ascontiguousarray(payload, dtype=np.float64)
        assert a.shape == (self.V, self.L.acvd_payload_size(m))
        self._ck(self.L.acvd_set_items(self.h, m, _p(a)))
        self.metric = m

    def items(self):
        out = np.zeros((self.V, self.L.acvd_payload_size(self.metric)))
        self._ck(self.L.acvd_get_items(self.h, _p(out)))
        return out

    def vertex_areas(self):
        out = np.zeros(self.V)
        self._ck(self.L.acvd_get_vertex_areas(self.h, _p(out)))
        return out

    # ---- clusters
    def set_num_clusters(self, K):
        self._ck(self.L.acvd_set_num_clusters(self.h, int(K)))
        self.K = int(K)

    def set_clustering(self, cl):
        a = np.ascontiguousarray(cl, dtype=np.int32)
        assert a.size == self.V
        self._ck(self.L.acvd_set_clustering(self.h, _p(a)))

    def clustering(self, out=None):
        if out is None:
            out = np.zeros(self.V, dtype=np.int32)
        self._ck(self.L.acvd_get_clustering(self.h, _p(out)))
        return out

    def save_clustering(self):
        self._ck(self.L.acvd_save_clustering(self.h))

    def restore_clustering(self):
        self._ck(self.L.acvd_restore_clustering(self.h))

    def set_frozen(self, flags):
        a = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
        self._ck(self.L.acvd_set_frozen(self.h, _p(a)))

    def set_fixed_clusters(self, items):
        a = np.ascontiguousarray(items, dtype=np.int64)
        self._ck(self.L.acvd_set_fixed_clusters(self.h, _p(a), a.size))

    def initial_sampling(self):
        self._ck(self.L.acvd_initial_sampling(self.h))

    # ---- hot path
    def minimize(self, **kw) -> dict:
        p = Params()
        p.quadrics_level = 3
        for k, v in kw.items():
            setattr(p, k, v)
        r = Report()
        self._ck(self.L.acvd_minimize(self.h, C.byref(p), C.byref(r)))
        return r.asdict()

    def recompute_statistics(self, constrained=1, quadrics_level=3):
        self._ck(self.L.acvd_recompute_statistics(self.h, int(constrained), int(quadrics_level)))

    def clean_clustering(self):
        d = C.c_int32()
        self._ck(self.L.acvd_clean_clustering(self.h, C.byref(d)))
        return d.value

    def fill_holes(self, connexity=0):
        self._ck(self.L.acvd_fill_holes(self.h, int(connexity)))

    def connexity_problem(self, items, clusters, mode=0):
        """Device connexity predicate on (item, cluster) pairs against the current clustering (uint8 array)."""
        it = np.ascontiguousarray(items, dtype=np.int32)
        cl = np.ascontiguousarray(clusters, dtype=np.int32)
        assert it.shape == cl.shape
        out = np.zeros(it.size, dtype=np.uint8)
        self._ck(self.L.acvd_connexity_problem(self.h, it.size, _p(it), _p(cl), int(mode), _p(out)))
        return out

    def reassign_round(self, constrained=1, quadrics_level=3, connexity=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.L.acvd_reassign_round(self.h, constrained, quadrics_level, connexity, C.byref(a), C.byref(b), C.byref(c)))
        return dict(proposals=a.value, modifications=b.value, tests=c.value)

    def cluster_stats(self, out=None):
        """acvd_get_cluster_stats.  `out` = (sums, centroids, energies, sizes) buffers of the caller (e.g. pinned, reused
        across calls: fresh pageable arrays cost more in page faults than the copy itself); default: new arrays."""
        np_ = self.L.acvd_payload_size(self.metric)
        if out is not None:
            sums, cen, en, sz = out
            assert sums.shape == (self.K, np_) and cen.shape == (self.K, 3) and en.shape == (self.K,) and sz.shape == (self.K,)
            assert sums.dtype == np.float64 and cen.dtype == np.float64 and en.dtype == np.float64 and sz.dtype == np.int32
        else:
            sums = np.zeros((self.K, np_))
            cen = np.zeros((self.K, 3))
            en = np.zeros(self.K)
            sz = np.zeros(self.K, dtype=np.int32)
        self._ck(self.L.acvd_get_cluster_stats(self.h, _p(sums), _p(cen), _p(en), _p(sz)))
        return sums, cen, en, sz

    def global_energy(self):
        e = C.c_double()
        self._ck(self.L.acvd_global_energy(self.h, C.byref(e)))
        return e.value

    def energy_log(self):
        n = C.c_int32()
        self._ck(self.L.acvd_get_energy_log(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        self._ck(self.L.acvd_get_energy_log(self.h, _p(out), n.value, C.byref(n)))
        return out

    def energy_times(self):
        """seconds since the start of the last minimize() at which each entry of energy_log() was taken"""
        n = C.c_int32()
        self._ck(self.L.acvd_get_energy_times(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        self._ck(self.L.acvd_get_energy_times(self.h, _p(out), n.value, C.byref(n)))
        return out

    def representative_points(self, quadrics9, points3, max_sv=3, sv_threshold=1e-3):
        q = np.ascontiguousarray(quadrics9, dtype=np.float64).reshape(-1, 9)
        p = np.array(points3, dtype=np.float64).reshape(-1, 3).copy()
        rd = np.zeros(q.shape[0], dtype=np.int32)
        self._ck(self.L.acvd_representative_points(self.h, q.shape[0], _p(q), _p(p), max_sv, sv_threshold, _p(rd)))
        return p, rd

    def cluster_quadrics(self, n_clusters=None):
        """Per-cluster sums of the quadrics of the input faces around every item (ACVD post-process): [n, 9]."""
        n = self.K if n_clusters is None else int(n_clusters)
        out = np.zeros((n, 9))
        self._ck(self.L.acvd_cluster_quadrics(self.h, n, _p(out)))
        return out

    # ---- integer stages
    def boundary_flags(self):
        out = np.zeros(self.V, dtype=np.uint8)
        self._ck(self.L.acvd_boundary_flags(self.h, _p(out)))
        return out

    def cluster_adjacency(self):
        n = C.c_int64()
        self._ck(self.L.acvd_cluster_adjacency(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        self._ck(self.L.acvd_cluster_adjacency(self.h, _p(out), n.value, C.byref(n)))
        return np.stack([out >> 32, out & 0xFFFFFFFF], axis=1).astype(np.int32)

    def dual_triangles(self):
        n = C.c_int64()
        self._ck(self.L.acvd_dual_triangles(self.h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), dtype=np.int32)
        if n.value:
            self._ck(self.L.acvd_dual_triangles(self.h, _p(out), n.value, C.byref(n)))
        return out

    def input_manifold_flags(self):
        out = np.zeros(self.V, dtype=np.uint8)
        self._ck(self.L.acvd_input_manifold_flags(self.h, _p(out)))
        return out

    def output_manifold_flags(self, force_manifold_edges=1):
        out = np.zeros(self.K, dtype=np.uint8)
        self._ck(self.L.acvd_output_manifold_flags(self.h, int(force_manifold_edges), _p(out)))
        return out

    def detect_non_manifold(self, force_manifold_edges=1):
        """One DetectNonManifoldOutputVertices step; returns the number of issues, self.K follows the grown count."""
        n, k = _i32(), _i32()
        self._ck(self.L.acvd_detect_non_manifold(self.h, int(force_manifold_edges), C.byref(n), C.byref(k)))
        self.K = k.value
        return n.value

    def frozen(self):
        out = np.zeros(self.K, dtype=np.uint8)
        self._ck(self.L.acvd_get_frozen(self.h, _p(out)))
        return out

    # ---- multi-GPU
    def bench_kernel(self, kernel=0, variant=0, stage=0, reps=20) -> float:
        """ms per launch of one kernel on the current state (CUDA events on the library stream)."""
        ms = C.c_float()
        self._ck(self.L.acvd_bench_kernel(self.h, kernel, variant, stage, reps, C.byref(ms)))
        return float(ms.value)

    @staticmethod
    def dist_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(NCCL_ID_BYTES)
        rc = L.acvd_dist_unique_id(buf)
        if rc != 0:
            raise AcvdError(rc, "ncclGetUniqueId failed")
        return buf.raw

    def dist_init(self, rank, world, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, NCCL_ID_BYTES)
        self._ck(self.L.acvd_dist_init(self.h, rank, world, buf))
