"""Minimal binary PLY reader/writer (float32 xyz, uchar+int32 triangle lists): the on-disk format of the
front-ends (simplification.ply, Remeshing.ply)."""
from __future__ import annotations

import numpy as np


def write_ply(path, points, triangles):
    p = np.ascontiguousarray(points, dtype="<f4")
    t = np.ascontiguousarray(triangles, dtype="<i4")
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {p.shape[0]}\nproperty float x\nproperty float y\n"
                 f"property float z\nelement face {t.shape[0]}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        f.write(p.tobytes())
        rec = np.empty(t.shape[0], dtype=[("n", "u1"), ("v", "<i4", (3,))])
        rec["n"] = 3
        rec["v"] = t
        f.write(rec.tobytes())


def read_ply(path):
    with open(path, "rb") as f:
        nv = nf = 0
        fmt = None
        while True:
            line = f.readline().decode().strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element vertex"):
                nv = int(line.split()[2])
            elif line.startswith("element face"):
                nf = int(line.split()[2])
            elif line == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError("read_ply: only binary_little_endian files written by this package are supported")
        p = np.frombuffer(f.read(12 * nv), dtype="<f4").reshape(nv, 3).copy()
        rec = np.frombuffer(f.read(13 * nf), dtype=[("n", "u1"), ("v", "<i4", (3,))])
        return p, rec["v"].copy()
