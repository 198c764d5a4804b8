"""Drop-in for the hot-path part of the reference's ``utils/rend_util.py``:
``near_far_from_sphere`` (:167-185), ``get_sphere_intersection`` (:188-210),
``get_dvals_from_radius`` (:213-234), ``sample_pdf`` (:255-292), ``sample_cdf`` (:294-327).
All of them run in the CUDA library; CPU tensors raise."""
import torch

from .. import _lib


def near_far_from_sphere(ray_origins, ray_directions, r=1.0, keepdim=True):
    """rend_util.py:167-185.  ``ray_directions`` already normalised."""
    _lib.require_cuda(ray_origins, ray_directions)
    lib = _lib.get_lib()
    shape = ray_origins.shape[:-1]
    o = _lib.f32c(ray_origins.reshape(-1, 3))
    d = _lib.f32c(ray_directions.expand_as(ray_origins).reshape(-1, 3))
    R = o.shape[0]
    near = torch.empty(R, dtype=torch.float32, device=o.device)
    far = torch.empty(R, dtype=torch.float32, device=o.device)
    with torch.cuda.device(o.device):
        _lib.check(lib.nr_near_far_from_sphere(_lib.ptr(o), _lib.ptr(d), R, float(r), _lib.ptr(near),
                                               _lib.ptr(far), _lib.stream_ptr(o.device)), "near_far_from_sphere")
    out_shape = (*shape, 1) if keepdim else shape
    return near.reshape(out_shape), far.reshape(out_shape)


def _invert(bins, w_or_cdf, N_importance, det, eps, cdf_given, u=None, return_details=False):
    _lib.require_cuda(bins, w_or_cdf)
    lib = _lib.get_lib()
    prefix = bins.shape[:-1]
    M = bins.shape[-1]
    assert w_or_cdf.shape[-1] == M - 1, "weights/cdf must have one entry per interval"
    b = _lib.f32c(bins.detach().reshape(-1, M))
    w = _lib.f32c(w_or_cdf.detach().reshape(-1, M - 1))
    R, dev = b.shape[0], b.device
    if u is None and not det:
        # same RNG call as the reference: torch.rand(prefix + [N]) on the device (rend_util.py:271)
        u = torch.rand(list(prefix) + [N_importance], device=dev)
    if u is not None:
        u = _lib.f32c(u.reshape(-1, N_importance))
    samples = torch.empty(R, N_importance, dtype=torch.float32, device=dev)
    below = above = cdf = None
    if return_details:
        below = torch.empty(R, N_importance, dtype=torch.int32, device=dev)
        above = torch.empty(R, N_importance, dtype=torch.int32, device=dev)
        cdf = torch.empty(R, M, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nr_sample_pdf(_lib.ptr(b), _lib.ptr(w), _lib.ptr(u), R, M, N_importance, int(cdf_given),
                                     float(eps), _lib.ptr(samples), _lib.ptr(below), _lib.ptr(above), _lib.ptr(cdf),
                                     _lib.stream_ptr(dev)), "sample_pdf")
    samples = samples.reshape(*prefix, N_importance)
    if return_details:
        return (samples, below.reshape(*prefix, N_importance), above.reshape(*prefix, N_importance),
                cdf.reshape(*prefix, M))
    return samples


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None, return_details=False):
    """rend_util.py:255-292.  Extra keyword ``u`` supplies the uniforms explicitly (parity
    tests); ``return_details`` also returns (below, above, cdf)."""
    return _invert(bins, weights, N_importance, det, eps, False, u, return_details)


def sample_cdf(bins, cdf, N_importance, det=False, eps=1e-5, u=None, return_details=False):
    """rend_util.py:294-327 (``cdf`` is the un-normalised CDF with one entry per interval)."""
    return _invert(bins, cdf, N_importance, det, eps, True, u, return_details)


def get_sphere_intersection(ray_origins, ray_directions, r=1.0):
    """rend_util.py:188-210: exact ray / sphere (near, far, mask_intersect), each [..., 1]; ``nr_sphere_intersection``."""
    _lib.require_cuda(ray_origins, ray_directions)
    lib = _lib.get_lib()
    prefix = ray_origins.shape[:-1]
    o, d = _lib.f32c(ray_origins.reshape(-1, 3)), _lib.f32c(ray_directions.expand_as(ray_origins).reshape(-1, 3))
    R, dev = o.shape[0], o.device
    near, far = torch.empty(R, dtype=torch.float32, device=dev), torch.empty(R, dtype=torch.float32, device=dev)
    mask = torch.empty(R, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nr_sphere_intersection(_lib.ptr(o), _lib.ptr(d), R, float(r), _lib.ptr(near), _lib.ptr(far),
                                              _lib.ptr(mask), _lib.stream_ptr(dev)), "sphere_intersection")
    return near.reshape(*prefix, 1), far.reshape(*prefix, 1), mask.bool().reshape(*prefix, 1)


def get_dvals_from_radius(ray_origins, ray_directions, rs, far_end=True, strict=False):
    """rend_util.py:213-234: depth at which the ray is ``rs`` away from the origin; rs [..., N] -> d_vals [..., N];
    ``nr_dvals_from_radius``.  The reference asserts ``rs^2 > |o|^2 - (o.d)^2`` with a host read of the whole tensor
    (:225); here violating entries come back NaN (what ``torch.sqrt`` gives) without a sync, and ``strict=True`` performs
    the reference's assertion (one host read of a device counter)."""
    _lib.require_cuda(ray_origins, ray_directions, rs)
    lib = _lib.get_lib()
    prefix = ray_origins.shape[:-1]
    o, d = _lib.f32c(ray_origins.reshape(-1, 3)), _lib.f32c(ray_directions.expand_as(ray_origins).reshape(-1, 3))
    R, dev = o.shape[0], o.device
    N = rs.shape[-1]
    rs_f = _lib.f32c(rs.expand(*prefix, N).reshape(R, N))
    out = torch.empty(R, N, dtype=torch.float32, device=dev)
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if strict else None
    with torch.cuda.device(dev):
        _lib.check(lib.nr_dvals_from_radius(_lib.ptr(o), _lib.ptr(d), _lib.ptr(rs_f), R, N, int(bool(far_end)),
                                            _lib.ptr(out), _lib.ptr(bad), _lib.stream_ptr(dev)), "dvals_from_radius")
    if strict:
        assert int(bad.item()) == 0, "get_dvals_from_radius: rs inside the ray's closest approach (rend_util.py:225)"
    return out.reshape(*prefix, N)


def quat_to_rot(q):
    """rend_util.py:76-93 (tiny host-side composition; only used by get_rays for 7-vector poses)."""
    import torch.nn.functional as F
    q = F.normalize(q, dim=-1)
    qr, qi, qj, qk = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = torch.stack([
        1 - 2 * (qj ** 2 + qk ** 2), 2 * (qj * qi - qk * qr), 2 * (qi * qk + qr * qj),
        2 * (qj * qi + qk * qr), 1 - 2 * (qi ** 2 + qk ** 2), 2 * (qj * qk - qi * qr),
        2 * (qk * qi - qj * qr), 2 * (qj * qk + qi * qr), 1 - 2 * (qi ** 2 + qj ** 2)], dim=-1)
    return R.reshape(*q.shape[:-1], 3, 3)


def get_rays(c2w, intrinsics, H, W, N_rays=-1):
    """rend_util.py:112-164: (rays_o, rays_d, select_inds) for [..., 4, 4] (or [..., 7] quaternion + position)
    poses.  Pixel selection draws the same two ``torch.randint`` calls on the CPU generator as the reference."""
    _lib.require_cuda(c2w)
    lib = _lib.get_lib()
    dev = c2w.device
    if c2w.shape[-1] == 7:
        p = torch.eye(4, device=dev).repeat([*c2w.shape[:-1], 1, 1]).float()
        p[..., :3, :3] = quat_to_rot(c2w[..., :4])
        p[..., :3, 3] = c2w[..., 4:]
    else:
        p = c2w
    prefix = p.shape[:-2]
    B = 1
    for s in prefix:
        B *= s
    pose = _lib.f32c(p.reshape(B, 4, 4))
    K = intrinsics.to(dev).float().expand(*prefix, *intrinsics.shape[-2:]).reshape(B, *intrinsics.shape[-2:])
    intr = torch.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 0, 1]], dim=-1).contiguous()
    if N_rays > 0:
        N_rays = min(N_rays, H * W)
        select_hs = torch.randint(0, H, size=[N_rays]).to(dev)
        select_ws = torch.randint(0, W, size=[N_rays]).to(dev)
        select_inds = (select_hs * W + select_ws).expand([*prefix, N_rays])
        N = N_rays
        sel = select_inds.reshape(B, N).contiguous()
    else:
        N = H * W
        select_inds = torch.arange(N, device=dev).expand([*prefix, N])
        sel = None
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nr_get_rays(_lib.ptr(pose), _lib.ptr(intr), _lib.ptr(sel), B, W, N, _lib.ptr(rays_o), _lib.ptr(rays_d),
                                   _lib.stream_ptr(dev)), "get_rays")
    return rays_o.reshape(*prefix, N, 3), rays_d.reshape(*prefix, N, 3), select_inds
