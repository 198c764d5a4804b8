"""GPU: dense SDF-grid query (BASELINE config 5) against the oracle on the reference's lattice."""
import numpy as np
import pytest
import torch

import neurecon_b200
from conftest import build_neus, cpu_state_dict, rel_err
from oracle import nets
from neurecon_b200.utils import mesh_util

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference_lattice(N, s):
    """utils/mesh_util.py:83-100 restated (np.int -> np.int64; the true division is the reference's)."""
    idx = np.arange(0, N ** 3, 1).astype(np.int64)
    xyz = np.zeros([N ** 3, 3])
    xyz[:, 2] = idx % N
    xyz[:, 1] = (idx / N) % N
    xyz[:, 0] = ((idx / N) / N) % N
    org = -s / 2.0
    xyz[:, 0] = xyz[:, 0] * (s / (N - 1)) + org
    xyz[:, 1] = xyz[:, 1] * (s / (N - 1)) + org
    xyz[:, 2] = xyz[:, 2] * (s / (N - 1)) + org
    return torch.from_numpy(xyz).float()


@pytest.mark.parametrize("tier,tol", [("fp32", 1e-5), ("fp16", 1e-2)])
def test_grid_query_matches_oracle(tier, tol):
    neurecon_b200.set_precision(tier)
    try:
        N, s = 24, 2.0
        m = build_neus(seed=1, device=DEV)
        L = nets.layers_from_state_dict(cpu_state_dict(m), "implicit_surface.surface_fc_layers", 9)
        pts = reference_lattice(N, s)
        want_sdf, want_nab, _ = nets.sdf_forward_with_nablas(pts, L)
        sdf, nab = mesh_util.query_sdf_grid(m.implicit_surface, N=N, volume_size=s, with_nablas=True, plane_range=(0, N),
                                            chunk=5000)
        assert sdf.shape == (N, N, N) and nab.shape == (N, N, N, 3)
        assert rel_err(sdf.reshape(-1), want_sdf) < tol and rel_err(nab.reshape(-1, 3), want_nab) < max(tol, 1e-4)
        # slabs concatenate to the whole grid (how ranks shard it), sdf-only path
        a = mesh_util.query_sdf_grid(m.implicit_surface, N=N, volume_size=s, plane_range=(0, 10))
        b = mesh_util.query_sdf_grid(m.implicit_surface, N=N, volume_size=s, plane_range=(10, N))
        whole = mesh_util.query_sdf_grid(m.implicit_surface, N=N, volume_size=s, plane_range=(0, N))
        assert torch.equal(torch.cat([a, b]), whole) and rel_err(whole.reshape(-1), want_sdf) < tol
    finally:
        neurecon_b200.set_precision("fp16")


def test_lattice_bit_exact():
    from neurecon_b200 import _lib
    N, s = 17, 2.0
    pts = torch.empty(N ** 3, 3, device=DEV)
    _lib.check(_lib.get_lib().nr_grid_points(0, N ** 3, N, s, 1, _lib.ptr(pts), _lib.stream_ptr()), "grid")
    assert torch.equal(pts.cpu(), reference_lattice(N, s))
