"""Oracle restatement of VolSDF volume rendering (TEST INFRASTRUCTURE ONLY).

Follows models/frameworks/volsdf.py:16-35 (sdf_to_sigma), :38-74 (error_bound), :77-272
(fine_sample: error-bounded beta iteration with bisection), :306-331 (model queries) and
:334-551 (volume_render), inference semantics.  The per-ray state machine of fine_sample is
restated ray-parallel with explicit masks instead of boolean-index gathers; results are the
same because every ray's computation is independent of the others.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import nets, sampling


def sdf_to_sigma(sdf, alpha, beta):
    """volsdf.py:16-35: Laplace-CDF density."""
    e = 0.5 * torch.exp(-torch.abs(sdf) / beta)
    return alpha * torch.where(sdf >= 0, e, 1 - e)


def error_bound(d_vals, sdf, alpha, beta):
    """volsdf.py:38-74."""
    sigma = sdf_to_sigma(sdf, alpha, beta)
    sabs = torch.abs(sdf)
    delta = d_vals[..., 1:] - d_vals[..., :-1]
    R_t = torch.cat([torch.zeros_like(sdf[..., :1]), torch.cumsum(sigma[..., :-1] * delta, dim=-1)], dim=-1)[..., :-1]
    d_star = torch.clamp_min(0.5 * (sabs[..., :-1] + sabs[..., 1:] - delta), 0.0)
    errors = alpha / (4 * beta) * (delta ** 2) * torch.exp(-d_star / beta)
    errors_t = torch.cumsum(errors, dim=-1)
    bounds = torch.exp(-R_t) * (torch.exp(errors_t) - 1.0)
    return torch.where(torch.isnan(bounds), torch.full_like(bounds, float("inf")), bounds)


def opacity_invert_cdf_sample(d_vals, sdf, alpha, beta, n, det=True, u=None):
    """volsdf.py:102-116."""
    sigma = sdf_to_sigma(sdf, alpha, beta)
    delta = d_vals[..., 1:] - d_vals[..., :-1]
    R_t = torch.cat([torch.zeros_like(sdf[..., :1]), torch.cumsum(sigma[..., :-1] * delta, dim=-1)], dim=-1)[..., :-1]
    opacity = 1 - torch.exp(-R_t)
    return sampling.sample_cdf(d_vals, opacity, n, det=det, u=u)


def fine_sample(sdf_fn, init_dvals, rays_o, rays_d, alpha_net, beta_net, far, eps=0.1, max_iter=5,
                max_bisection=10, final_N_importance=64, N_up=128, perturb=False):
    """volsdf.py:77-272, restated per ray (loop over rays; small cases only).  Returns
    (d_fine [R,N], beta [R,1], iter_usage [R])."""
    R, N0 = init_dvals.shape
    dt = init_dvals.dtype
    if not torch.is_tensor(far):
        far = far * torch.ones(R, 1, dtype=dt)
    out_d = torch.zeros(R, final_N_importance, dtype=dt)
    out_beta = torch.zeros(R, 1, dtype=dt)
    out_iter = torch.zeros(R, dtype=dt)
    alpha_net = alpha_net.reshape(()) if torch.is_tensor(alpha_net) else torch.tensor(alpha_net, dtype=dt)
    beta_net = beta_net.reshape(()) if torch.is_tensor(beta_net) else torch.tensor(beta_net, dtype=dt)

    def query(d, o, dr):
        return sdf_fn(o[None, :] + dr[None, :] * d[:, None])

    for r in range(R):
        o, dr = rays_o[r], rays_d[r]
        d = init_dvals[r]
        beta = torch.sqrt((far[r, 0] ** 2) / (4 * (N0 - 1) * np.log(1 + eps)))
        alpha = 1.0 / beta
        sdf = query(d, o, dr)
        net_max = error_bound(d, sdf, alpha_net, beta_net).max()
        if not (net_max > eps):
            out_d[r] = opacity_invert_cdf_sample(d, sdf, alpha_net, beta_net, final_N_importance, det=not perturb)
            out_iter[r] = 0
            out_beta[r, 0] = beta_net
            continue
        bounds = error_bound(d, sdf, alpha, beta)
        converged = False
        it = 0
        while it < max_iter:
            it += 1
            up = sampling.sample_pdf(d[None], bounds[None], N_up + 2, det=True)[0, 1:-1]
            sdf_up = query(up, o, dr)
            d_cat = torch.cat([d, up])
            d, idx = torch.sort(d_cat)
            sdf = torch.cat([sdf, sdf_up])[idx]
            net_max = error_bound(d, sdf, alpha_net, beta_net).max()
            if not (net_max > eps):
                out_d[r] = opacity_invert_cdf_sample(d, sdf, alpha_net, beta_net, final_N_importance, det=not perturb)
                out_iter[r] = it
                converged = True
                break
            b_right, b_left = beta.clone(), beta_net.clone()
            for _ in range(max_bisection):
                b_tmp = 0.5 * (b_left + b_right)
                bmax = error_bound(d, sdf, 1.0 / b_tmp, b_tmp).max()
                if bmax <= eps:
                    b_right = b_tmp
                if bmax > eps:
                    b_left = b_tmp
            beta = b_right
            alpha = 1.0 / beta
            bounds = torch.clamp(error_bound(d, sdf, alpha, beta), 0, 1e5)
        if converged:
            out_beta[r, 0] = beta_net
        else:
            out_d[r] = opacity_invert_cdf_sample(d, sdf, 1.0 / beta, beta, final_N_importance, det=not perturb)
            out_iter[r] = -1
            out_beta[r, 0] = beta
    return out_d, out_beta, out_iter


def composite(sigma, radiances, nablas, d_all, white_bkgd=False, calc_normal=True):
    """volsdf.py:482-503."""
    delta = d_all[..., 1:] - d_all[..., :-1]
    p_i = torch.exp(-F.relu(sigma[..., :-1] * delta))
    tau = (1 - p_i + 1e-10) * torch.cumprod(torch.cat([torch.ones_like(p_i[..., :1]), p_i], dim=-1), dim=-1)[..., :-1]
    rgb = (tau[..., None] * radiances[..., :-1, :]).sum(-2)
    depth = (tau / (tau.sum(-1, keepdim=True) + 1e-10) * d_all[..., :-1]).sum(-1)
    acc = tau.sum(-1)
    if white_bkgd:
        rgb = rgb + (1.0 - acc[..., None])
    ret = OrderedDict(rgb=rgb, depth_volume=depth, mask_volume=acc)
    if calc_normal:
        n = F.normalize(nablas, dim=-1)
        N = min(tau.shape[-1], n.shape[-2])
        ret["normals_volume"] = (n[..., :N, :] * tau[..., :N, None]).sum(-2)
    ret.update(alpha=1.0 - p_i, p_i=p_i, visibility_weights=tau)
    return ret


def volume_render(rays_o, rays_d, sd, cfg, near=0.0, far=6.0, obj_bounding_radius=3.0, calc_normal=True,
                  white_bkgd=False, perturb=False, N_samples=128, N_importance=64, max_upsample_steps=5,
                  max_bisection_steps=10, epsilon=0.1, use_nerfplusplus=False, N_outside=32, dtype=torch.float32):
    """volsdf.py:334-551, one ray chunk, inference."""
    rays_o = rays_o.reshape(-1, 3).to(dtype)
    rays_d = F.normalize(rays_d.reshape(-1, 3).to(dtype), dim=-1)
    R = rays_o.shape[0]
    sdf_layers = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", cfg["D"] + 1, dtype=dtype)
    rad_layers = nets.layers_from_state_dict(sd, "radiance_net.layers", cfg["D_rad"] + 1, dtype=dtype)
    mr, skips = cfg["multires"], tuple(cfg["skips"])
    beta = torch.exp(sd["ln_beta"].to(dtype) * cfg["speed_factor"])
    alpha = 1.0 / beta
    use_sphere_bg = not use_nerfplusplus

    def surface(p):  # VolSDF.forward_surface, volsdf.py:310-315
        s = nets.sdf_forward(p, sdf_layers, mr, skips)
        return torch.min(s, obj_bounding_radius - p.norm(dim=-1)) if use_sphere_bg else s

    nears = near * torch.ones(R, 1, dtype=dtype)
    if use_nerfplusplus:
        _, fars, mask = sampling.get_sphere_intersection(rays_o, rays_d, r=obj_bounding_radius)
        assert mask.all()
    else:
        fars = far * torch.ones(R, 1, dtype=dtype)
    t = sampling.linspace01(N_samples, dtype)
    d_coarse = nears * (1 - t) + fars * t
    t4 = sampling.linspace01(N_samples * 4, dtype)
    d_init = nears * (1 - t4) + fars * t4
    d_fine, beta_map, iter_usage = fine_sample(surface, d_init, rays_o, rays_d, alpha, beta, fars, eps=epsilon,
                                               max_iter=max_upsample_steps, max_bisection=max_bisection_steps,
                                               final_N_importance=N_importance, N_up=N_samples * 4, perturb=perturb)
    d_all, _ = torch.sort(torch.cat([d_coarse, d_fine], dim=-1), dim=-1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * d_all[..., :, None]
    sdf, nablas, feat = nets.sdf_forward_with_nablas(pts, sdf_layers, mr, skips)
    if use_sphere_bg:  # forward_surface_with_nablas, volsdf.py:317-325
        d_bg = obj_bounding_radius - pts.norm(dim=-1)
        sdf = torch.where(d_bg < sdf, d_bg, sdf)
    views = rays_d.unsqueeze(-2).expand_as(pts)
    radiances = nets.radiance_forward(pts, views, nablas, feat, rad_layers, cfg["rad_multires"], cfg["multires_view"])
    sigma = sdf_to_sigma(sdf, alpha, beta)
    extra = {}
    if use_nerfplusplus:
        tt = sampling.linspace01(N_outside + 2, dtype)[1:-1]
        rs = (obj_bounding_radius / torch.flip(tt, dims=[-1])).expand(R, N_outside)
        d_out = sampling.get_dvals_from_radius(rays_o, rays_d, rs)
        pts_out = rays_o[..., None, :] + rays_d[..., None, :] * d_out[..., :, None]
        x_out = torch.cat([pts_out / rs[..., None], 1.0 / rs[..., None]], dim=-1)
        sigma_out, radiance_out = nets.nerf_forward(x_out, rays_d.unsqueeze(-2).expand_as(pts_out), sd)
        d_all = torch.cat([d_all, d_out], dim=-1)
        sigma = torch.cat([sigma, sigma_out], dim=-1)
        radiances = torch.cat([radiances, radiance_out], dim=-2)
        extra = dict(sigma_out=sigma_out, radiance_out=radiance_out)
    ret = composite(sigma, radiances, nablas, d_all, white_bkgd, calc_normal)
    ret.update(implicit_surface=sdf, implicit_nablas=nablas, radiance=radiances, d_vals=d_all, sigma=sigma,
               beta_map=beta_map, iter_usage=iter_usage, **extra)
    return ret["rgb"], ret["depth_volume"], ret
