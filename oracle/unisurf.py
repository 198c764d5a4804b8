"""Oracle restatement of UNISURF volume rendering (TEST INFRASTRUCTURE ONLY).

Follows models/ray_casting.py:11-30 (secant), :35-160 (root finding), and
models/frameworks/unisurf.py:34-62 (model queries, occupancy<->logit) and :64-283
(volume_render), inference semantics, ``batched=True`` layout with the leading B dropped
(B = 1): every tensor here is [R, ...].
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import nets, sampling


def get_opacity_from_surface(x):
    """unisurf.py:54-62."""
    odds = torch.exp(-1.0 * x)
    return odds / (1 + odds)


def root_finding(sdf_fn, rays_o, rays_d, near, far, N_steps=256, logit_tau=0.0, N_secant_steps=8):
    """ray_casting.py:35-160 with fill_inf=False.  Returns d_pred_out [R], pt_pred [R,3],
    mask [R], mask_sign_change [R]."""
    R = rays_o.shape[0]
    dt = rays_o.dtype
    t = sampling.linspace01(N_steps, dt)[None, :]
    d_prop = near[:, None] * (1 - t) + far[:, None] * t
    pts = rays_o.unsqueeze(-2) + d_prop.unsqueeze(-1) * rays_d.unsqueeze(-2)
    val = sdf_fn(pts) - logit_tau
    mask_0_not_occupied = val[..., 0] > 0
    sign_matrix = torch.cat([torch.sign(val[..., :-1] * val[..., 1:]), torch.ones(R, 1, dtype=dt)], dim=-1)
    cost = sign_matrix * torch.arange(N_steps, 0, -1, dtype=dt)
    values, indices = torch.min(cost, -1)
    mask_sign_change = values < 0
    ar = torch.arange(R)
    mask_pos_to_neg = val[ar, indices] > 0
    mask = mask_sign_change & mask_pos_to_neg & mask_0_not_occupied
    d_high, f_high = d_prop[ar, indices], val[ar, indices]
    ind2 = torch.clamp(indices + 1, max=N_steps - 1)
    d_low, f_low = d_prop[ar, ind2], val[ar, ind2]
    # secant (ray_casting.py:11-30) on every ray, results used where mask is set
    d_low, f_low, d_high, f_high = d_low.clone(), f_low.clone(), d_high.clone(), f_high.clone()
    d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    for _ in range(N_secant_steps):
        p_mid = rays_o + d_pred.unsqueeze(-1) * rays_d
        f_mid = sdf_fn(p_mid) - logit_tau
        low = f_mid < 0
        d_low = torch.where(low, d_pred, d_low)
        f_low = torch.where(low, f_mid, f_low)
        d_high = torch.where(low, d_high, d_pred)
        f_high = torch.where(low, f_high, f_mid)
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    pt_pred = torch.where(mask[:, None], rays_o + d_pred.unsqueeze(-1) * rays_d, torch.ones(R, 3, dtype=dt))
    d_out = torch.where(mask, d_pred, far)
    d_out = torch.where(mask_0_not_occupied, d_out, torch.zeros_like(d_out))
    return d_out, pt_pred, mask, mask_sign_change


def composite(logits, radiances, nablas, d_all, white_bkgd=False, calc_normal=True):
    """unisurf.py:216-240."""
    alpha = get_opacity_from_surface(logits)
    shifted = torch.cat([torch.ones_like(alpha[..., :1]), 1.0 - alpha + 1e-10], dim=-1)
    w = alpha * torch.cumprod(shifted, dim=-1)[..., :-1]
    rgb = (w[..., None] * radiances).sum(-2)
    depth = (w / (w.sum(-1, keepdim=True) + 1e-10) * d_all).sum(-1)
    acc = w.sum(-1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    ret = OrderedDict(rgb=rgb, depth_volume=depth, mask_volume=acc)
    if calc_normal:
        n = F.normalize(nablas, dim=-1)
        ret["normals_volume"] = (n * w[..., None]).sum(-2)
    ret.update(alpha=alpha, visibility_weights=w)
    return ret


def volume_render(rays_o, rays_d, sd, cfg, calc_normal=True, logit_tau=0.0, white_bkgd=False,
                  radius_of_interest=4.0, interval=1.0, too_close_threshold=0.1, N_query=64, N_freespace=32,
                  near_bypass=None, far_bypass=None, dtype=torch.float32):
    """unisurf.py:64-283 (perturb=False), one ray chunk == one net chunk, B = 1 dropped."""
    rays_o = rays_o.reshape(-1, 3).to(dtype)
    rays_d = F.normalize(rays_d.reshape(-1, 3).to(dtype), dim=-1)
    sdf_layers = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", cfg["D"] + 1, dtype=dtype)
    rad_layers = nets.layers_from_state_dict(sd, "radiance_net.layers", cfg["D_rad"] + 1, dtype=dtype)
    mr, skips = cfg["multires"], tuple(cfg["skips"])
    sdf_fn = lambda p: nets.sdf_forward(p, sdf_layers, mr, skips)

    near, far = sampling.near_far_from_sphere(rays_o, rays_d, r=radius_of_interest, keepdim=False)
    if near_bypass is not None:
        near = near_bypass * torch.ones_like(near)
    if far_bypass is not None:
        far = far_bypass * torch.ones_like(far)
    d_threshold = near + (far - near) * too_close_threshold
    d_pred_out, pt_pred, mask, mask_sign_change = root_finding(sdf_fn, rays_o, rays_d, near, far, logit_tau=logit_tau)
    d_pred = torch.max(torch.min(d_pred_out, far), near)
    d_upper = torch.min(d_pred + interval, far)
    d_lower = torch.max(d_pred - interval, near)
    t = sampling.linspace01(N_query, dtype)
    d_int = d_lower.unsqueeze(-1) * (1 - t) + d_upper.unsqueeze(-1) * t
    d_lower = torch.max(d_lower, d_threshold)
    d_lower = torch.where(mask_sign_change == 0, far, d_lower)
    d_lower = torch.where(d_lower < 1e-10, far, d_lower)
    t = sampling.linspace01(N_freespace, dtype)
    d_free = near[..., None] * (1 - t) + d_lower.unsqueeze(-1) * t
    d_all, _ = torch.sort(torch.cat([d_free, d_int], dim=-1), dim=-1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * d_all[..., :, None]
    logits, nablas, feat = nets.sdf_forward_with_nablas(pts, sdf_layers, mr, skips)
    # UNISURF.forward (unisurf.py:34-38): F.normalize(nablas) has no dim -> dim=1, which in the
    # batched [B, points, 3] layout is the POINT axis of the whole net chunk
    normals = F.normalize(nablas.reshape(1, -1, 3)).reshape(nablas.shape)
    views = rays_d.unsqueeze(-2).expand_as(pts)
    radiances = nets.radiance_forward(pts, views, normals, feat, rad_layers, cfg["rad_multires"], cfg["multires_view"])
    ret = composite(logits, radiances, nablas, d_all, white_bkgd, calc_normal)
    ret.update(surface_points=pt_pred, mask_surface=mask, depth_surface=d_pred, radiance=radiances,
               implicit_surface=logits, implicit_nablas=nablas, d_all=d_all, mask_sign_change=mask_sign_change)
    return ret["rgb"], ret["depth_volume"], ret


def sphere_tracing(sdf_fn, rays_o, rays_d, near=0.0, far=6.0, N_iters=20):
    """ray_casting.py:163-184."""
    d = torch.ones(rays_o.shape[:-1], dtype=rays_o.dtype) * near
    mask = torch.ones_like(d, dtype=torch.bool)
    for _ in range(N_iters):
        val = sdf_fn(rays_o + rays_d * d[..., None])
        d = torch.where(mask, d + val, d)
        mask = mask & ~(d > far) & ~(d < 0)
    return d, rays_o + rays_d * d[..., None], mask
