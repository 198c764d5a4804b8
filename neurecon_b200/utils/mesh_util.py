"""Dense SDF-grid query: the hot part of the reference's ``utils/mesh_util.py:extract_mesh`` (:82-111).
Marching cubes and the PLY writer (:13-80) are CPU post-processing and stay with the reference."""
import torch

from .. import _lib
from . import dist_util


def query_sdf_grid(implicit_surface, N=512, volume_size=2.0, with_nablas=False, plane_range=None, chunk=1 << 21,
                   faithful_lattice=True, device=None):
    """SDF (and optionally its analytic gradient) on x-planes ``plane_range = (lo, hi)`` of the N^3 lattice
    of ``extract_mesh`` (default: this rank's share of the planes, `dist_util.shard_range`).  Lattice
    points are generated on the device; nothing crosses PCIe except the result you choose to copy.
    Returns sdf [hi-lo, N, N] (and nablas [hi-lo, N, N, 3]).  ``faithful_lattice`` keeps the reference's
    true-division lattice (mesh_util.py:92-94)."""
    lib = _lib.get_lib()
    dev = torch.device(device) if device is not None else next(implicit_surface.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("neurecon_b200: the grid query runs only on CUDA (there is no CPU fallback)")
    lo, hi = dist_util.shard_range(N) if plane_range is None else plane_range
    count = (hi - lo) * N * N
    sdf = torch.empty(count, dtype=torch.float32, device=dev)
    nab = torch.empty(count, 3, dtype=torch.float32, device=dev) if with_nablas else None
    pts = torch.empty(min(chunk, max(count, 1)), 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), torch.no_grad():
        st = _lib.stream_ptr(dev)
        for c0 in range(0, count, chunk):
            m = min(chunk, count - c0)
            _lib.check(lib.nr_grid_points(lo * N * N + c0, m, N, float(volume_size), int(faithful_lattice), _lib.ptr(pts), st),
                       "grid_points")
            s_, n_, _ = implicit_surface._run(pts[:m], want_nablas=with_nablas, want_feat=False)
            sdf[c0:c0 + m] = s_
            if with_nablas:
                nab[c0:c0 + m] = n_
    sdf = sdf.reshape(hi - lo, N, N)
    return (sdf, nab.reshape(hi - lo, N, N, 3)) if with_nablas else sdf
