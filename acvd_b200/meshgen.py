"""Synthetic input surfaces for the ACVD clustering path.

Every generator returns ``(points float32 [V,3], triangles int32 [F,3])`` with
outward (counter-clockwise) orientation, fixed seeds, and closed topology.
Points are float32 because the reference stores mesh points in a default
``vtkPoints`` (float32) and widens to double on read
(reference Common/vtkSurfaceBase.cxx:1389-1391).

Workloads of BASELINE.json map to:
  C1  geodesic_icosphere(128)                       V = 163 842
  C2  noisy_torus(2000, 1300)                       V = 2 600 000
  C3  ridged_ellipsoid(316)                         V = 998 562
  C4  displaced_sphere(2000)                        V = 40 000 002
  C5  thin_torus(16000, 10000)                      V = 160 000 000
"""
from __future__ import annotations

import numpy as np

_T = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = np.array(
    [[-1, _T, 0], [1, _T, 0], [-1, -_T, 0], [1, -_T, 0],
     [0, -1, _T], [0, 1, _T], [0, -1, -_T], [0, 1, -_T],
     [_T, 0, -1], [_T, 0, 1], [-_T, 0, -1], [-_T, 0, 1]], dtype=np.float64)
_ICO_F = np.array(
    [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
     [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
     [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
     [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def geodesic_icosphere(n: int, dtype=np.float32):
    """Class-I geodesic subdivision of the icosahedron with frequency ``n``.

    V = 10 n^2 + 2, F = 20 n^2, E = 30 n^2.  ``n = 2**L`` gives the same vertex
    count as L rounds of 1->4 subdivision (reference Common/vtkSurface.cxx:605).
    Vertex ids: 12 corners, then 30 x (n-1) edge points, then per-face interior
    points row by row, so a face's interior is contiguous in memory.
    """
    assert n >= 1
    corners = _ICO_V / np.linalg.norm(_ICO_V[0])
    edges = {}
    for f in _ICO_F:
        for a, b in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            key = (min(a, b), max(a, b))
            if key not in edges:
                edges[key] = len(edges)
    n_edge_pts = n - 1
    n_int = (n - 1) * (n - 2) // 2
    V = 12 + 30 * n_edge_pts + 20 * n_int
    pts = np.empty((V, 3), dtype=np.float64)
    pts[:12] = corners
    k = np.arange(1, n, dtype=np.float64)[:, None]
    for (a, b), e in edges.items():
        base = 12 + e * n_edge_pts
        pts[base:base + n_edge_pts] = (corners[a] * (n - k) + corners[b] * k) / n
    tris = np.empty((20 * n * n, 3), dtype=np.int64)
    ii, jj = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    tpos = 0
    for fi, (A, B, C) in enumerate(_ICO_F):
        T = np.full((n + 1, n + 1), -1, dtype=np.int64)
        T[0, 0], T[n, 0], T[0, n] = A, B, C
        if n > 1:
            kk = np.arange(1, n)

            def edge_ids(a, b):
                e = edges[(min(a, b), max(a, b))]
                base = 12 + e * n_edge_pts
                return base + (kk - 1) if a < b else base + (n - kk - 1)

            T[kk, 0] = edge_ids(A, B)          # j = 0, i = k along A->B
            T[0, kk] = edge_ids(A, C)          # i = 0, j = k along A->C
            T[n - kk, kk] = edge_ids(B, C)     # i + j = n, j = k along B->C
        if n_int > 0:
            m = (ii >= 1) & (jj >= 1) & (ii + jj <= n - 1)
            i_in, j_in = ii[m], jj[m]
            # rows j = 1..n-2 hold (n-1-j) interior points each
            row_off = (j_in - 1) * (n - 1) - (j_in - 1) * j_in // 2
            ids = 12 + 30 * n_edge_pts + fi * n_int + row_off + (i_in - 1)
            T[i_in, j_in] = ids
            w = (corners[A][None, :] * (n - i_in - j_in)[:, None]
                 + corners[B][None, :] * i_in[:, None]
                 + corners[C][None, :] * j_in[:, None]) / n
            pts[ids] = w
        up = (ii + jj <= n - 1)
        iu, ju = ii[up], jj[up]
        nu = iu.size
        tris[tpos:tpos + nu, 0] = T[iu, ju]
        tris[tpos:tpos + nu, 1] = T[iu + 1, ju]
        tris[tpos:tpos + nu, 2] = T[iu, ju + 1]
        tpos += nu
        dn = (ii + jj <= n - 2)
        idn, jdn = ii[dn], jj[dn]
        nd = idn.size
        tris[tpos:tpos + nd, 0] = T[idn + 1, jdn]
        tris[tpos:tpos + nd, 1] = T[idn + 1, jdn + 1]
        tris[tpos:tpos + nd, 2] = T[idn, jdn + 1]
        tpos += nd
    assert tpos == tris.shape[0]
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    return pts.astype(dtype), tris.astype(np.int32)


def torus_grid(nu: int, nv: int, R: float = 1.0, r: float = 0.35, noise: float = 0.0,
               seed: int = 1, dtype=np.float32, return_uv: bool = False):
    """Closed torus on an ``nu`` (major) x ``nv`` (minor) grid, quads split on a fixed
    diagonal; V = nu*nv, F = 2V, E = 3V.  ``noise`` is the sigma of a radial
    (tube-normal) Gaussian perturbation."""
    u = (np.arange(nu, dtype=np.float64) * (2 * np.pi / nu))[:, None]
    v = (np.arange(nv, dtype=np.float64) * (2 * np.pi / nv))[None, :]
    rr = np.full((nu, nv), r, dtype=np.float64)
    if noise > 0:
        rng = np.random.Generator(np.random.PCG64(seed))
        rr = rr + rng.normal(0.0, noise, size=(nu, nv))
    x = (R + rr * np.cos(v)) * np.cos(u)
    y = (R + rr * np.cos(v)) * np.sin(u)
    z = rr * np.sin(v) + 0 * u
    pts = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(dtype)
    i = np.arange(nu, dtype=np.int64)[:, None]
    j = np.arange(nv, dtype=np.int64)[None, :]
    i1 = (i + 1) % nu
    j1 = (j + 1) % nv
    a = (i * nv + j).ravel()
    b = (i1 * nv + j).ravel()
    c = (i1 * nv + j1).ravel()
    d = (i * nv + j1).ravel()
    tris = np.empty((2 * nu * nv, 3), dtype=np.int32)
    tris[0::2, 0], tris[0::2, 1], tris[0::2, 2] = a, b, c
    tris[1::2, 0], tris[1::2, 1], tris[1::2, 2] = a, c, d
    if return_uv:
        uu = np.broadcast_to(u, (nu, nv)).ravel()
        vv = np.broadcast_to(v, (nu, nv)).ravel()
        return pts, tris, uu, vv
    return pts, tris


def torus_curvature_indicator(nu: int, nv: int, R: float = 1.0, r: float = 0.35):
    """Analytic curvature indicator sqrt(k1^2 + k2^2) of the *smooth* torus on the same grid
    (k1 = 1/r, k2 = cos v / (R + r cos v)).  Stands in for vtkCurvatureMeasure
    (reference DiscreteRemeshing/vtkDiscreteRemeshing.h:640-653) until §8f-2 is built;
    runs that use it are labelled "analytic-curvature"."""
    v = (np.arange(nv, dtype=np.float64) * (2 * np.pi / nv))[None, :]
    k1 = 1.0 / r
    k2 = np.cos(v) / (R + r * np.cos(v))
    ind = np.sqrt(k1 * k1 + k2 * k2)
    return np.broadcast_to(ind, (nu, nv)).ravel().copy()


def noisy_torus(nu: int = 2000, nv: int = 1300, dtype=np.float32):
    """C2: R=1, r=0.35, radial Gaussian noise sigma=0.002, seed 1."""
    return torus_grid(nu, nv, 1.0, 0.35, noise=0.002, seed=1, dtype=dtype)


def thin_torus(nu: int = 16000, nv: int = 10000, dtype=np.float32):
    """C5: R=1, r=0.078125 -> 8:1 elongated cells along the major circle."""
    return torus_grid(nu, nv, 1.0, 0.078125, dtype=dtype)


def banded_thin_torus(nu: int = 16000, nv: int = 10000, band: float = 0.1, stretch: int = 8, dtype=np.float32):
    """C5 input: closed thin torus R = 1, r = 0.078125 with 8:1 elongated cells along the major circle, where a
    band of the major-circle columns (`band` of the circumference) is coarsened `stretch` times: its edges exceed 3 x the
    mean edge length, so `-l 3` (vtkSurface::SplitLongEdges) really cuts there.  `nu` is the number of major-circle columns
    the *uniform* torus would have; the band keeps every `stretch`-th of its columns.  Built column block by column block
    straight into float32 / int32 (the full-size mesh is 146 M vertices, 291 M triangles)."""
    n_band = int(round(nu * band))
    cols = np.concatenate([np.arange(0, n_band, stretch), np.arange(n_band, nu)]).astype(np.float64)
    nc = cols.size
    R, r = 1.0, 0.078125
    v = np.arange(nv) * (2 * np.pi / nv)
    ring_r = (R + r * np.cos(v))                       # distance of the minor-circle points from the axis
    ring_z = (r * np.sin(v)).astype(dtype)
    pts = np.empty((nc * nv, 3), dtype=dtype)
    tris = np.empty((2 * nc * nv, 3), dtype=np.int32)
    jn = ((np.arange(nv) + 1) % nv).astype(np.int64)
    j = np.arange(nv, dtype=np.int64)
    block = 512
    for c0 in range(0, nc, block):
        c1 = min(nc, c0 + block)
        u = cols[c0:c1] * (2 * np.pi / nu)
        sl = slice(c0 * nv, c1 * nv)
        pts[sl, 0] = (np.cos(u)[:, None] * ring_r[None, :]).reshape(-1)
        pts[sl, 1] = (np.sin(u)[:, None] * ring_r[None, :]).reshape(-1)
        pts[sl, 2] = np.broadcast_to(ring_z[None, :], (c1 - c0, nv)).reshape(-1)
        i = np.arange(c0, c1, dtype=np.int64)[:, None]
        i1 = (i + 1) % nc
        a = (i * nv + j[None, :]).reshape(-1)
        b = (i1 * nv + j[None, :]).reshape(-1)
        c = (i1 * nv + jn[None, :]).reshape(-1)
        d = (i * nv + jn[None, :]).reshape(-1)
        # faces (a, b, c) of all cells first, then (a, c, d): the order torus_grid uses (fixed diagonal)
        tris[sl, 0] = a; tris[sl, 1] = b; tris[sl, 2] = c
        s2 = slice(nc * nv + c0 * nv, nc * nv + c1 * nv)
        tris[s2, 0] = a; tris[s2, 1] = c; tris[s2, 2] = d
    return pts, tris


def displaced_sphere(n: int = 2000, amp: float = 0.05, seed: int = 2, dtype=np.float32):
    """C4: geodesic icosphere radially displaced by amp * sum_k a_k sin(f_k d_k.p + phi_k)."""
    pts, tris = geodesic_icosphere(n, dtype=np.float64)
    rng = np.random.Generator(np.random.PCG64(seed))
    nk = 6
    d = rng.normal(size=(nk, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    f = rng.uniform(2.0, 12.0, size=nk)
    phi = rng.uniform(0, 2 * np.pi, size=nk)
    a = rng.uniform(0.5, 1.0, size=nk)
    a /= a.sum()
    disp = np.zeros(pts.shape[0], dtype=np.float64)
    for k in range(nk):
        disp += a[k] * np.sin(f[k] * (pts @ d[k]) + phi[k])
    pts *= (1.0 + amp * disp)[:, None]
    return pts.astype(dtype), tris


def ridged_ellipsoid(n: int = 316, axes=(1.0, 0.6, 0.4), ridge: float = 0.03, freq: int = 24,
                     dtype=np.float32):
    """C3: geodesic icosphere mapped to an ellipsoid with sinusoidal ridges r(1 + ridge sin(freq theta))."""
    pts, tris = geodesic_icosphere(n, dtype=np.float64)
    theta = np.arctan2(pts[:, 1], pts[:, 0])
    rad = 1.0 + ridge * np.sin(freq * theta)
    pts = pts * np.asarray(axes, dtype=np.float64)[None, :] * rad[:, None]
    return pts.astype(dtype), tris


def ellipsoid_principal_directions(points, axes=(1.0, 0.6, 0.4)):
    """Analytic curvature of the ellipsoid x^2/a^2 + y^2/b^2 + z^2/c^2 = 1 at the radial projections of ``points``
    (the C3 inputs the reference takes from vtkCurvatureMeasure: "analytic-curvature" runs).

    Returns ``(pd, indicator)``: ``pd[v] = (sqrt|k1| d1, sqrt|k2| d2)`` as float32 with the larger |k| first
    (the PrincipalDirections layout of Common/vtkCurvatureMeasure.cxx:445-484) and ``indicator = sqrt(k1^2 + k2^2)``.
    The shape operator of the level set F = 1 is P H P / |grad F| with P = I - n n^T and H = diag(2/a^2, 2/b^2, 2/c^2)."""
    ax = np.asarray(axes, dtype=np.float64)
    q = np.asarray(points, dtype=np.float64)
    q = q / np.sqrt(((q / ax) ** 2).sum(axis=1))[:, None]          # onto the ellipsoid
    grad = 2.0 * q / ax ** 2
    gn = np.linalg.norm(grad, axis=1)
    n = grad / gn[:, None]
    P = np.eye(3)[None, :, :] - n[:, :, None] * n[:, None, :]
    H = np.diag(2.0 / ax ** 2)
    S = P @ H[None, :, :] @ P / gn[:, None, None]
    w, vec = np.linalg.eigh(S)                                      # ascending; one eigenvalue ~0 along n
    order = np.argsort(-np.abs(w), axis=1)                          # |k| descending: k1, k2, ~0
    idx = np.arange(q.shape[0])
    k1, k2 = w[idx, order[:, 0]], w[idx, order[:, 1]]
    d1, d2 = vec[idx, :, order[:, 0]], vec[idx, :, order[:, 1]]
    pd = np.concatenate([np.sqrt(np.abs(k1))[:, None] * d1, np.sqrt(np.abs(k2))[:, None] * d2], axis=1)
    return pd.astype(np.float32), np.sqrt(k1 * k1 + k2 * k2)


def subdivide(points, triangles):
    """One 1->4 midpoint subdivision (old points first, then one midpoint per edge)."""
    p = np.asarray(points, dtype=np.float64)
    t = np.asarray(triangles, dtype=np.int64)
    V = p.shape[0]
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    key = np.minimum(e[:, 0], e[:, 1]) * V + np.maximum(e[:, 0], e[:, 1])
    uniq, inv = np.unique(key, return_inverse=True)
    mid = 0.5 * (p[uniq // V] + p[uniq % V])
    m = (V + inv).reshape(3, -1).T          # midpoints of (v0v1, v1v2, v2v0) per face
    a, b, c = t[:, 0], t[:, 1], t[:, 2]
    m01, m12, m20 = m[:, 0], m[:, 1], m[:, 2]
    nt = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1), np.stack([m12, c, m20], 1),
                         np.stack([m01, m12, m20], 1)])
    return np.concatenate([p, mid]).astype(np.float32), nt.astype(np.int32)


def bipyramid(n: int = 24, levels: int = 4):
    """Closed surface with two vertices of valence ``n`` (the apexes) -- exercises adjacency rows longer than
    the ELL width -- subdivided ``levels`` times (the apex valence is preserved by midpoint subdivision)."""
    ang = np.arange(n) * (2 * np.pi / n)
    ring = np.stack([np.cos(ang), np.sin(ang), 0 * ang], axis=1)
    pts = np.concatenate([ring, [[0, 0, 0.8]], [[0, 0, -0.8]]])
    top, bot = n, n + 1
    tris = []
    for i in range(n):
        j = (i + 1) % n
        tris.append([i, j, top])
        tris.append([j, i, bot])
    p, t = pts.astype(np.float32), np.asarray(tris, dtype=np.int32)
    for _ in range(levels):
        p, t = subdivide(p, t)
    return p, t


def workload(name: str):
    """Named workloads used by bench.py and the tests."""
    if name == "C1":
        p, t = geodesic_icosphere(128)
        return dict(points=p, triangles=t, K=3000, metric="iso", gradation=0.0, indicator=None)
    if name == "C2":
        p, t = noisy_torus()
        return dict(points=p, triangles=t, K=100000, metric="qem", gradation=1.5,
                    indicator=torus_curvature_indicator(2000, 1300))
    if name == "C2s":  # 1/16-size C2 for CPU-side tests
        p, t = torus_grid(500, 325, noise=0.002, seed=1)
        return dict(points=p, triangles=t, K=6250, metric="qem", gradation=1.5,
                    indicator=torus_curvature_indicator(500, 325))
    if name in ("C3", "C3s"):  # AnisotropicRemeshingQ 1.5 on the ridged ellipsoid (C3s: 1/16 size)
        p, t = ridged_ellipsoid(316 if name == "C3" else 79)
        pd, ind = ellipsoid_principal_directions(p)
        return dict(points=p, triangles=t, K=10000 if name == "C3" else 625, metric="anisoq", gradation=1.5,
                    indicator=ind, pd=pd)
    if name in ("C5", "C5s"):
        # ACVD m K 0 -m 1 -l 3 on the banded thin torus: V = vertex count AFTER SplitLongEdges (what the clustering sees);
        # the split itself is part of the run (acvd_split_long_edges), `split_ratio` tells the caller to apply it
        if name == "C5":
            p, t = banded_thin_torus(16000, 10000)
            K = 1600000
        else:   # 1/256 scale twin
            p, t = banded_thin_torus(1000, 625)
            K = 6250
        return dict(points=p, triangles=t, K=K, metric="iso", gradation=0.0, indicator=None, split_ratio=3.0, force_manifold=1)
    if name == "C4":
        p, t = displaced_sphere(2000)
        return dict(points=p, triangles=t, K=400000, metric="qem", gradation=0.0, indicator=None)
    if name == "C4s":  # 1/100-size C4
        p, t = displaced_sphere(200)
        return dict(points=p, triangles=t, K=4000, metric="qem", gradation=0.0, indicator=None)
    raise KeyError(name)
