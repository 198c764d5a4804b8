"""Iso-surface extraction on the device (csrc/marching_cubes.cu) against the numpy oracle, bit for bit, and the
``utils/mesh_util.py`` drop-ins (``convert_sigma_samples_to_ply``, ``extract_mesh``) end to end."""
import os

import numpy as np
import pytest
import torch

import neurecon_b200
from neurecon_b200.utils import mesh_util
from oracle import mesh
from conftest import build_neus

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _volumes():
    N = 40
    g = np.linspace(-1, 1, N, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    rs = np.random.RandomState(3)
    yield "sphere", (np.sqrt(X * X + Y * Y + Z * Z) - 0.6).astype(np.float32), 0.0, (2 / 39,) * 3
    yield "torus_level", (np.sqrt((np.sqrt(X * X + Y * Y) - 0.55) ** 2 + Z * Z)).astype(np.float32), 0.2, (0.05, 0.07, 0.11)
    yield "noise", rs.normal(size=(23, 37, 18)).astype(np.float32), 0.1, (1.0, 1.0, 1.0)
    yield "noise_open", rs.normal(size=(9, 2, 31)).astype(np.float32), -0.3, (1.0, 2.0, 0.5)
    yield "ties", np.round(rs.normal(size=(16, 16, 16)) * 2).astype(np.float32), 0.0, (1.0, 1.0, 1.0)   # many values == level
    yield "single_cell", np.array([[[-1, 1], [1, 1]], [[1, 1], [1, -2]]], np.float32), 0.0, (1.0, 1.0, 1.0)
    yield "empty", np.ones((6, 5, 4), np.float32), 0.0, (1.0, 1.0, 1.0)


@pytest.mark.parametrize("direction", ["descent", "ascent"])
def test_marching_cubes_bit_exact_vs_oracle(direction):
    for name, vol, level, spacing in _volumes():
        want_v, want_f = mesh.marching_cubes(vol, level, spacing, direction)
        v, f = mesh_util.marching_cubes(torch.from_numpy(vol).to(DEV), level=level, spacing=spacing, gradient_direction=direction)
        assert v.dtype == torch.float32 and f.dtype == torch.int32 and v.is_cuda
        assert v.shape == want_v.shape and f.shape == want_f.shape, name
        assert np.array_equal(v.cpu().numpy().view(np.uint32), want_v.view(np.uint32)), name
        assert np.array_equal(f.cpu().numpy(), want_f), name


def test_marching_cubes_large_grid_properties():
    """256^3 (larger than the oracle handles in seconds): closed genus-1 surface, area of the torus, one launch pair"""
    N = 256
    g = torch.linspace(-1, 1, N, device=DEV)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    R, r0 = 0.55, 0.2
    vol = torch.sqrt((torch.sqrt(X * X + Y * Y) - R) ** 2 + Z * Z) - r0
    h = 2.0 / (N - 1)
    v, f = mesh_util.marching_cubes(vol, 0.0, (h, h, h))
    closed, euler, area, volume = mesh.mesh_stats(v.cpu().numpy(), f.cpu().numpy())
    assert closed and euler == 0
    assert abs(area - 4 * np.pi ** 2 * R * r0) / (4 * np.pi ** 2 * R * r0) < 1e-3
    assert abs(-volume - 2 * np.pi ** 2 * R * r0 ** 2) / (2 * np.pi ** 2 * R * r0 ** 2) < 1e-3


def test_extract_mesh_end_to_end(tmp_path):
    """mesh_util.extract_mesh (reference :82-111): sdf grid from the network, iso-surface, PLY -- against the oracle run on
    the same grid, and the file parsed back"""
    neurecon_b200.set_precision("fp32")
    try:
        m = build_neus(seed=1, device=DEV)
        N, s = 64, 2.0
        path = os.path.join(tmp_path, "surface.ply")
        pts, faces = mesh_util.extract_mesh(m.implicit_surface, volume_size=s, level=0.0, N=N, filepath=path, show_progress=False)
        sdf = mesh_util.query_sdf_grid(m.implicit_surface, N=N, volume_size=s, plane_range=(0, N))
    finally:
        neurecon_b200.set_precision("fp16")
    want_v, want_f = mesh.marching_cubes(sdf.cpu().numpy(), 0.0, (s / N,) * 3)
    want_pts = (np.float32(-s / 2) + want_v).astype(np.float32)                 # mesh_util.py:39-42
    assert np.array_equal(faces, want_f) and np.array_equal(pts, want_pts)
    header, v, n, idx = mesh.read_ply(path)
    assert np.array_equal(v, pts) and np.array_equal(idx, faces) and (n == 3).all()
    closed, euler, area, volume = mesh.mesh_stats(pts, faces)
    assert closed and euler == 2 and len(faces) > 1000                          # the sphere-initialised surface (radius_init 0.5)
    # convert_sigma_samples_to_ply with a numpy grid, scale and offset (mesh_util.py:44-48)
    p2, f2 = mesh_util.convert_sigma_samples_to_ply(sdf.cpu().numpy(), [-1.0, -1.0, -1.0], [s / N] * 3, path, level=0.0,
                                                    offset=np.array([0.1, 0.2, 0.3], np.float32), scale=2.0)
    assert np.array_equal(f2, faces) and np.allclose(p2, pts / 2.0 - np.array([0.1, 0.2, 0.3], np.float32), atol=1e-7)
