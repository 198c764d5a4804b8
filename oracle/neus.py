"""Oracle restatement of NeuS volume rendering (TEST INFRASTRUCTURE ONLY).

Follows models/frameworks/neus.py:21-70 (cdf / alpha / weights) and :118-397
(volume_render, 'official_solution' up-sampling), inference semantics.
"""
import math
from collections import OrderedDict
import torch
import torch.nn.functional as F

from . import nets, sampling


def sdf_to_alpha(sdf, s):
    """neus.py:28-35."""
    cdf = torch.sigmoid(sdf * s)
    alpha = ((cdf[..., :-1] - cdf[..., 1:]) / (cdf[..., :-1] + 1e-10)).clamp_min(0)
    return cdf, alpha


def alpha_to_w(alpha):
    """neus.py:57-70: w_i = alpha_i * prod_{j<i} (1 - alpha_j + 1e-10)."""
    shifted = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], dim=-1)
    return alpha * torch.cumprod(shifted, dim=-1)[..., :-1]


def upsample_weights(d, sdf, it):
    """One 'official_solution' iteration up to the weights (neus.py:252-269)."""
    prev_sdf, next_sdf = sdf[..., :-1], sdf[..., 1:]
    prev_z, next_z = d[..., :-1], d[..., 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    dot = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-5)
    prev_dot = torch.cat([torch.zeros_like(dot[..., :1]), dot[..., :-1]], dim=-1)
    dot = torch.minimum(prev_dot, dot).clamp(-10.0, 0.0)
    dist = next_z - prev_z
    prev_esti = mid_sdf - dot * dist * 0.5
    next_esti = mid_sdf + dot * dist * 0.5
    s = 64 * (2 ** it)
    prev_cdf = torch.sigmoid(prev_esti * s)
    next_cdf = torch.sigmoid(next_esti * s)
    alpha = (prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)
    return alpha_to_w(alpha)


def upsample(sdf_fn, rays_o, rays_d, d_coarse, n_importance=64, n_iters=4, perturb=False):
    """neus.py:249-277.  ``sdf_fn(pts[...,3]) -> sdf[...]``."""
    d = d_coarse
    sdf = sdf_fn(rays_o.unsqueeze(-2) + d.unsqueeze(-1) * rays_d.unsqueeze(-2))
    for it in range(n_iters):
        w = upsample_weights(d, sdf, it)
        d_fine = sampling.sample_pdf(d, w, n_importance // n_iters, det=not perturb)
        sdf_fine = sdf_fn(rays_o.unsqueeze(-2) + d_fine.unsqueeze(-1) * rays_d.unsqueeze(-2))
        d = torch.cat([d, d_fine], dim=-1)
        sdf = torch.cat([sdf, sdf_fine], dim=-1)
        d, idx = torch.sort(d, dim=-1)
        sdf = torch.gather(sdf, -1, idx)
    return d, sdf


def composite(sdf, nablas, radiances, d_all, s, white_bkgd=False, calc_normal=True):
    """neus.py:296,346-381 without the NeRF++ branch."""
    d_mid = 0.5 * (d_all[..., 1:] + d_all[..., :-1])
    cdf, alpha = sdf_to_alpha(sdf, s)
    w = alpha_to_w(alpha)
    rgb = (w[..., None] * radiances).sum(-2)
    depth = (w / (w.sum(-1, keepdim=True) + 1e-10) * d_mid).sum(-1)
    acc = w.sum(-1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    ret = OrderedDict(rgb=rgb, depth_volume=depth, mask_volume=acc)
    if calc_normal:
        n = F.normalize(nablas, dim=-1)
        N = min(w.shape[-1], n.shape[-2])
        ret["normals_volume"] = (n[..., :N, :] * w[..., :N, None]).sum(-2)
    ret.update(alpha=alpha, cdf=cdf, visibility_weights=w, d_final=d_mid)
    return ret


def composite_bg(sdf, nablas, radiances, d_all, s, rays_o, rays_d, far, sd, radius, N_outside, white_bkgd=False,
                 calc_normal=True):
    """neus.py:303-381 with the NeRF++ background (perturb=False)."""
    d_mid = 0.5 * (d_all[..., 1:] + d_all[..., :-1])
    pts_mid = rays_o[..., None, :] + rays_d[..., None, :] * d_mid[..., :, None]
    cdf, alpha_in = sdf_to_alpha(sdf, s)
    t = sampling.linspace01(N_outside + 2, sdf.dtype, sdf.device)[1:-1]
    d_out = far / torch.flip(t, dims=[-1])
    d_vals = torch.cat([d_mid, d_out], dim=-1)
    pts_out = rays_o[..., None, :] + rays_d[..., None, :] * d_vals[..., :, None]
    r = pts_out.norm(dim=-1, keepdim=True)
    x_out = torch.cat([pts_out / r, 1.0 / r], dim=-1)
    sigma_out, radiance_out = nets.nerf_forward(x_out, rays_d.unsqueeze(-2).expand_as(pts_out), sd)
    dists = torch.cat([d_vals[..., 1:] - d_vals[..., :-1], 1e10 * torch.ones_like(d_vals[..., :1])], dim=-1)
    alpha_out = 1 - torch.exp(-F.softplus(sigma_out) * dists)
    n1 = d_mid.shape[-1]
    inside = (pts_mid.norm(dim=-1) <= radius).to(sdf.dtype)
    alpha = torch.cat([alpha_in * inside + alpha_out[..., :n1] * (1 - inside), alpha_out[..., n1:]], dim=-1)
    rad = torch.cat([radiances * inside[..., None] + radiance_out[..., :n1, :] * (1 - inside)[..., None],
                     radiance_out[..., n1:, :]], dim=-2)
    w = alpha_to_w(alpha)
    rgb = (w[..., None] * rad).sum(-2)
    depth = (w / (w.sum(-1, keepdim=True) + 1e-10) * d_vals).sum(-1)
    acc = w.sum(-1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    ret = OrderedDict(rgb=rgb, depth_volume=depth, mask_volume=acc)
    if calc_normal:
        n = F.normalize(nablas, dim=-1)
        N = min(w.shape[-1], n.shape[-2])
        ret["normals_volume"] = (n[..., :N, :] * w[..., :N, None]).sum(-2)
    ret.update(alpha=alpha, cdf=cdf, visibility_weights=w, d_final=d_vals, radiance=rad, sigma_out=sigma_out,
               radiance_out=radiance_out)
    return ret


def volume_render(rays_o, rays_d, sd, cfg, obj_bounding_radius=1.0, calc_normal=True,
                  white_bkgd=False, perturb=False, N_samples=64, N_importance=64,
                  N_upsample_iters=4, near_bypass=None, far_bypass=None, dtype=torch.float32, N_outside=0):
    """neus.py:118-397 ('official_solution', no NeRF++), one ray chunk, inference.

    ``sd``: reference-layout state_dict; ``cfg``: dict(multires, multires_view,
    rad_multires, skips, D, D_rad, speed_factor).
    """
    rays_o = rays_o.reshape(-1, 3).to(dtype)
    rays_d = F.normalize(rays_d.reshape(-1, 3).to(dtype), dim=-1)
    sdf_layers = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", cfg["D"] + 1, dtype=dtype)
    rad_layers = nets.layers_from_state_dict(sd, "radiance_net.layers", cfg["D_rad"] + 1, dtype=dtype)
    mr, skips = cfg["multires"], tuple(cfg["skips"])
    s = torch.exp(sd["ln_s"].to(dtype) * cfg["speed_factor"])

    near, far = sampling.near_far_from_sphere(rays_o, rays_d, r=obj_bounding_radius)
    if near_bypass is not None:
        near = near_bypass * torch.ones_like(near)
    if far_bypass is not None:
        far = far_bypass * torch.ones_like(far)
    t = sampling.linspace01(N_samples, dtype, rays_o.device)
    d_coarse = near * (1 - t) + far * t

    sdf_fn = lambda p: nets.sdf_forward(p, sdf_layers, mr, skips)
    d_all, _ = upsample(sdf_fn, rays_o, rays_d, d_coarse, N_importance, N_upsample_iters, perturb)

    pts = rays_o[..., None, :] + rays_d[..., None, :] * d_all[..., :, None]
    d_mid = 0.5 * (d_all[..., 1:] + d_all[..., :-1])
    pts_mid = rays_o[..., None, :] + rays_d[..., None, :] * d_mid[..., :, None]
    sdf, nablas, _ = nets.sdf_forward_with_nablas(pts, sdf_layers, mr, skips)
    _, nab_mid, feat_mid = nets.sdf_forward_with_nablas(pts_mid, sdf_layers, mr, skips)
    views = rays_d.unsqueeze(-2).expand_as(pts_mid)
    radiances = nets.radiance_forward(pts_mid, views, nab_mid, feat_mid, rad_layers,
                                      cfg["rad_multires"], cfg["multires_view"])
    if N_outside > 0:
        ret = composite_bg(sdf, nablas, radiances, d_all, s, rays_o, rays_d, far, sd, obj_bounding_radius, N_outside,
                           white_bkgd, calc_normal)
        ret.update(implicit_nablas=nablas, implicit_surface=sdf, d_all=d_all)
        return ret["rgb"], ret["depth_volume"], ret
    ret = composite(sdf, nablas, radiances, d_all, s, white_bkgd, calc_normal)
    ret.update(implicit_nablas=nablas, implicit_surface=sdf, radiance=radiances, d_all=d_all)
    return ret["rgb"], ret["depth_volume"], ret
