"""Drop-in surface on the GPU: the three CLIs with the reference's argument grammar and output files."""
import os
import subprocess

import numpy as np
import pytest

from acvd_b200 import build, meshgen, meshio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bins():
    exes = build.build_host()
    return {os.path.basename(e): e for e in exes}


def edge_manifold_closed(t):
    e = np.sort(np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    return bool((cnt == 2).all())


def run(exe, args, cwd):
    r = subprocess.run([exe] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_acvd_cli_outputs(bins, tmp_path):
    p, t = meshgen.geodesic_icosphere(32)
    meshio.write_ply(tmp_path / "in.ply", p, t)
    out = run(bins["ACVD"], ["in.ply", 300, 0, "-m", 1], tmp_path)
    assert "The clustering took" in out
    ps, ts = meshio.read_ply(tmp_path / "smooth_simplification.ply")
    pq, tq = meshio.read_ply(tmp_path / "simplification.ply")
    assert ps.shape[0] >= 300 and pq.shape == ps.shape and np.array_equal(ts, tq)
    assert edge_manifold_closed(tq)                                   # -m 1: manifold output
    assert ts.shape[0] == 2 * ps.shape[0] - 4                         # closed genus-0 triangulation
    # smooth_* holds the cluster centroids (inside the sphere); the quadric post-process moves every vertex
    # outwards, to the point that minimises the distance to the tangent planes of its cluster (outside a convex cap)
    rs, rq = np.linalg.norm(ps, axis=1), np.linalg.norm(pq, axis=1)
    assert (rs < 1).all() and (rq > rs).all() and np.abs(rq - 1).max() < 0.01


def test_acvd_cli_subdivides_to_ratio(bins, tmp_path):
    p, t = meshgen.geodesic_icosphere(8)          # 642 vertices; -s 10 with 200 clusters needs >= 2000
    meshio.write_ply(tmp_path / "in.ply", p, t)
    out = run(bins["ACVD"], ["in.ply", 200, 0, "-q", 0, "-of", "out.ply"], tmp_path)
    assert out.count("Subdividing mesh") == 1
    po, to = meshio.read_ply(tmp_path / "out.ply")
    assert po.shape[0] == 200 and not (tmp_path / "smooth_out.ply").exists()


def test_acvdq_cli_with_gradation_and_energy_log(bins, tmp_path):
    p, t = meshgen.torus_grid(160, 100, noise=0.002, seed=1)
    meshio.write_ply(tmp_path / "torus.ply", p, t)
    out = run(bins["ACVDQ"], ["torus.ply", 400, 1.5, "-w", 1], tmp_path)
    assert "Performing unconstrained initialization" in out
    po, to = meshio.read_ply(tmp_path / "simplification.ply")
    assert po.shape[0] == 400 and to.shape[0] > 700
    lines = open(tmp_path / "energy.txt").read().strip().splitlines()
    assert lines[-1].startswith("Final Energy :") and len(lines) > 5


def test_acvdq_fixed_vertices_are_kept(bins, tmp_path):
    p, t = meshgen.geodesic_icosphere(24)
    meshio.write_ply(tmp_path / "in.ply", p, t)
    fixed = [5, 77, 1234, 4000]
    (tmp_path / "fixed.txt").write_text("\n".join(map(str, fixed)))
    out = run(bins["ACVDQ"], ["in.ply", 100, 0, "-fv", "fixed.txt"], tmp_path)
    assert "Constraints on vertices have been checked" in out
    po, _ = meshio.read_ply(tmp_path / "simplification.ply")
    assert po.shape[0] == 104 and np.array_equal(po[:4], p[fixed])


def test_anisotropic_cli(bins, tmp_path):
    p, t = meshgen.ridged_ellipsoid(40)
    meshio.write_ply(tmp_path / "ell.ply", p, t)
    run(bins["AnisotropicRemeshingQ"], ["ell.ply", 300, 1.5], tmp_path)
    po, to = meshio.read_ply(tmp_path / "Remeshing.ply")
    assert po.shape[0] == 300 and to.shape[0] > 500
