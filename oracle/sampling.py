"""Oracle restatement of the ray-sampling utilities (TEST INFRASTRUCTURE ONLY).

Follows utils/rend_util.py:167-234 (sphere geometry) and :255-327 (inverse-CDF
sampling).  ``linspace01`` restates the fp32 formula of ``torch.linspace`` so the
CUDA kernels can be checked for bit-equality of the deterministic ``u``.
"""
import numpy as np
import torch


def linspace01(n, dtype=torch.float32, device=None):
    """torch.linspace(0, 1, n) restated: step = fp32(1/(n-1)); the first half is
    ``step*i``, the second half ``1 - step*(n-1-i)``."""
    return _linspace01_cpu(n, dtype).to(device) if device is not None else _linspace01_cpu(n, dtype)


def _linspace01_cpu(n, dtype):
    if n == 1:
        return torch.zeros(1, dtype=dtype)
    i = np.arange(n, dtype=np.int64)
    if dtype == torch.float32:
        # fp32 FMA emulated exactly: the fp64 product of an fp32 step and a small
        # integer is exact, so one rounding to fp32 at the end equals fmaf().
        step = np.float64(np.float32(1.0) / np.float32(n - 1))
        lo = (step * i).astype(np.float32)
        hi = (1.0 - step * (n - 1 - i)).astype(np.float32)
        return torch.from_numpy(np.where(i < n // 2, lo, hi).astype(np.float32))
    step = 1.0 / (n - 1)
    return torch.from_numpy(np.where(i < n // 2, step * i, 1.0 - step * (n - 1 - i))).to(dtype)


def near_far_from_sphere(o, d, r=1.0, keepdim=True):
    """utils/rend_util.py:167-185."""
    mid = -(o * d).sum(-1, keepdim=keepdim)
    return (mid - r).clamp_min(0.0), (mid + r).clamp_min(r)


def get_sphere_intersection(o, d, r=1.0):
    """utils/rend_util.py:188-210."""
    o2 = (o * o).sum(-1, keepdim=True)
    od = (o * d).sum(-1, keepdim=True)
    under = od * od + r * r - o2
    mask = under > 0
    sq = torch.sqrt(under.clamp_min(0))
    near = torch.where(mask, -sq - od, torch.zeros_like(od)).clamp_min(0.0)
    far = torch.where(mask, sq - od, torch.zeros_like(od)).clamp_min(0.0)
    return near, far, mask


def get_dvals_from_radius(o, d, rs, far_end=True):
    """utils/rend_util.py:213-234."""
    o2 = (o * o).sum(-1, keepdim=True)
    od = (o * d).sum(-1, keepdim=True)
    under = rs * rs - (o2 - od * od)
    assert (under > 0).all()
    sq = torch.sqrt(under)
    return (-od + sq) if far_end else (-od - sq).clamp_min(0.0)


def search_lower_bound(cdf, u):
    """torch.searchsorted(cdf, u, right=False): first i with cdf[i] >= u."""
    return torch.searchsorted(cdf.contiguous(), u.contiguous(), right=False)


def invert_cdf(bins, cdf, u, eps=1e-5, return_inds=False):
    """Shared tail of sample_pdf / sample_cdf (rend_util.py:275-292 / :310-327)."""
    M = cdf.shape[-1]
    inds = search_lower_bound(cdf, u)
    below = (inds - 1).clamp_min(0)
    above = inds.clamp_max(M - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    samples = bin_b + t * (bin_a - bin_b)
    return (samples, below, above) if return_inds else samples


def make_u(prefix, n, det, dtype, generator=None, device=None):
    if det:
        return linspace01(n, dtype, device).expand(*prefix, n).contiguous()
    return torch.rand(*prefix, n, dtype=dtype, generator=generator, device=device)


def pdf_to_cdf(weights):
    """rend_util.py:258-264."""
    w = weights + 1e-5
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)


def sample_pdf(bins, weights, n, det=False, eps=1e-5, u=None, return_inds=False):
    """utils/rend_util.py:255-292."""
    cdf = pdf_to_cdf(weights)
    if u is None:
        u = make_u(cdf.shape[:-1], n, det, cdf.dtype, device=cdf.device)
    return invert_cdf(bins, cdf, u, eps, return_inds)


def sample_cdf(bins, cdf, n, det=False, eps=1e-5, u=None, return_inds=False):
    """utils/rend_util.py:294-327 (CDF given, not re-normalised)."""
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = make_u(cdf.shape[:-1], n, det, cdf.dtype, device=cdf.device)
    return invert_cdf(bins, cdf, u, eps, return_inds)


def get_rays(c2w, intrinsics, H, W, select_inds=None):
    """utils/rend_util.py:95-164 for [..., 4, 4] poses; ``select_inds`` [..., N] (pixel = h*W + w) or None = all."""
    prefix = c2w.shape[:-2]
    ii, jj = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i = ii.t().reshape([*[1] * len(prefix), H * W]).expand([*prefix, H * W])
    j = jj.t().reshape([*[1] * len(prefix), H * W]).expand([*prefix, H * W])
    if select_inds is not None:
        i, j = torch.gather(i, -1, select_inds), torch.gather(j, -1, select_inds)
    fx, fy = intrinsics[..., 0, 0].unsqueeze(-1), intrinsics[..., 1, 1].unsqueeze(-1)
    cx, cy, sk = intrinsics[..., 0, 2].unsqueeze(-1), intrinsics[..., 1, 2].unsqueeze(-1), intrinsics[..., 0, 1].unsqueeze(-1)
    z = torch.ones_like(i)
    x_lift = (i - cx + cy * sk / fy - sk * j / fy) / fx * z
    y_lift = (j - cy) / fy * z
    cam = torch.stack((x_lift, y_lift, z, torch.ones_like(z)), dim=-1).transpose(-1, -2)
    world = torch.matmul(c2w, cam).transpose(-1, -2)[..., :3]
    cam_loc = c2w[..., :3, 3]
    rays_d = world - cam_loc[..., None, :]
    return cam_loc[..., None, :].expand_as(rays_d), rays_d
