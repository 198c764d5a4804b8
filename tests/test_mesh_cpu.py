"""Iso-surface extraction, CPU side: the generated case table, the numpy oracle (oracle/mesh.py) on
shapes with known area / volume / topology, and the PLY byte layout of the reference's writer (utils/mesh_util.py:57-72)."""
import os

import numpy as np
import pytest

from neurecon_b200 import mc_tables
from neurecon_b200.utils import mesh_util
from oracle import mesh


def _grid(N):
    g = np.linspace(-1, 1, N, dtype=np.float32)
    return np.meshgrid(g, g, g, indexing="ij")


def test_case_table_properties():
    tri, cnt = mc_tables.tables()
    assert tri.shape == (256, 32) and cnt[0] == 0 and cnt[255] == 0 and cnt.max() == 5 and int(cnt.sum()) == 820
    for idx in range(256):
        row = tri[idx]
        n = int(cnt[idx])
        assert (row[: 3 * n] >= 0).all() and (row[3 * n:] == -1).all()
        # exactly the edges with one inside endpoint carry vertices
        used = set(row[: 3 * n].tolist())
        want = {e for e in range(12) if ((idx >> mc_tables.edge_corners(e)[0]) & 1) != ((idx >> mc_tables.edge_corners(e)[1]) & 1)}
        assert used == want, idx
        # inside the cube every triangle edge is either a loop edge (lies in a cube face, used once) or a fan diagonal (twice,
        # opposite directions): the case's patch is a union of discs
        he = {}
        for t in range(n):
            a, b, c = (int(v) for v in row[3 * t: 3 * t + 3])
            for u, v in ((a, b), (b, c), (c, a)):
                he[(u, v)] = he.get((u, v), 0) + 1
        assert all(v == 1 for v in he.values())


def test_oracle_sphere_and_torus():
    N = 48
    X, Y, Z = _grid(N)
    h = 2.0 / (N - 1)
    v, f = mesh.marching_cubes((np.sqrt(X * X + Y * Y + Z * Z) - 0.6).astype(np.float32), 0.0, (h, h, h))
    closed, euler, area, vol = mesh.mesh_stats(v, f)
    assert closed and euler == 2
    assert abs(area - 4 * np.pi * 0.36) / (4 * np.pi * 0.36) < 5e-3
    assert abs(-vol - 4 / 3 * np.pi * 0.6 ** 3) / (4 / 3 * np.pi * 0.6 ** 3) < 5e-3      # 'descent': normals towards the inside of an sdf
    r = np.linalg.norm(v - 1.0, axis=1)                                                  # array origin = (-1, -1, -1)
    assert abs(r - 0.6).max() < 0.6 * h * h                                             # linear interpolation: O(h^2 / r)
    va, fa = mesh.marching_cubes((np.sqrt(X * X + Y * Y + Z * Z) - 0.6).astype(np.float32), 0.0, (h, h, h), "ascent")
    assert np.array_equal(va, v) and np.array_equal(fa, f[:, [0, 2, 1]])
    R, r0 = 0.55, 0.2
    v, f = mesh.marching_cubes((np.sqrt((np.sqrt(X * X + Y * Y) - R) ** 2 + Z * Z) - r0).astype(np.float32), 0.0, (h, h, h))
    closed, euler, area, vol = mesh.mesh_stats(v, f)
    assert closed and euler == 0 and abs(area - 4 * np.pi ** 2 * R * r0) / (4 * np.pi ** 2 * R * r0) < 1e-2


def test_oracle_closed_on_noise_with_every_ambiguous_case():
    rs = np.random.RandomState(0)
    tri, cnt = mc_tables.tables()
    seen = np.zeros(256, dtype=bool)
    for trial in range(4):
        vol = rs.normal(size=(22, 19, 25)).astype(np.float32)
        vol[0] = vol[-1] = 5
        vol[:, 0] = vol[:, -1] = 5
        vol[:, :, 0] = vol[:, :, -1] = 5
        v, f = mesh.marching_cubes(vol, 0.1)
        closed, _, _, _ = mesh.mesh_stats(v, f)
        assert closed, "crack between neighbouring cells"
        inside = vol < 0.1
        case = np.zeros((21, 18, 24), dtype=np.int64)
        for c in range(8):
            dx, dy, dz = mc_tables.corner_offset(c)
            case |= inside[dx:21 + dx, dy:18 + dy, dz:24 + dz].astype(np.int64) << c
        seen[np.unique(case)] = True
    assert seen.sum() >= 250          # nearly all of the 256 cases occur (corner 0-planes are forced outside)


def test_degenerate_volumes():
    v, f = mesh.marching_cubes(np.ones((5, 4, 3), np.float32), 0.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    vol = np.ones((2, 2, 2), np.float32)
    vol[0, 0, 0] = -1
    v, f = mesh.marching_cubes(vol, 0.0, (2.0, 3.0, 4.0))
    assert f.shape == (1, 3) and sorted(map(tuple, v.tolist())) == [(0.0, 0.0, 2.0), (0.0, 1.5, 0.0), (1.0, 0.0, 0.0)]


def test_ply_layout_matches_plyfile(tmp_path):
    """header and records as plyfile writes them for the reference's two elements (mesh_util.py:57-72)"""
    rs = np.random.RandomState(1)
    verts = rs.normal(size=(7, 3)).astype(np.float32)
    faces = rs.randint(0, 7, size=(5, 3)).astype(np.int32)
    path = os.path.join(tmp_path, "m.ply")
    mesh_util.write_ply(path, verts, faces)
    raw = open(path, "rb").read()
    head = (b"ply\nformat binary_little_endian 1.0\nelement vertex 7\nproperty float x\nproperty float y\nproperty float z\n"
            b"element face 5\nproperty list uchar int vertex_indices\nend_header\n")
    assert raw.startswith(head) and len(raw) == len(head) + 7 * 12 + 5 * 13
    header, v, n, idx = mesh.read_ply(path)
    assert np.array_equal(v, verts) and np.array_equal(idx, faces) and (n == 3).all()
