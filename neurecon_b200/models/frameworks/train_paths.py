"""VolSDF / UNISURF ``volume_render`` under autograd (training).  Same split as ``neus_train.py``:
no-grad samplers on the inference kernels, the with-grad network query through ``models/autograd.py``
(hand-written fp32 forward/backward), compositing as differentiable tensor ops (gradients reach ``ln_beta``,
the sdf / logits and the radiances)."""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ... import _lib
from ..autograd import exclusive_cumprod


def volsdf_render_train(rays_o, rays_d, model, near, far, obj_bounding_radius, batched, calc_normal, rayschunk,
                        white_bkgd, use_nerfplusplus, detailed_output, perturb, N_samples, N_importance, N_outside,
                        max_upsample_steps, max_bisection_steps, epsilon):
    from . import volsdf
    from ...utils import rend_util
    lib = _lib.get_lib()
    B = rays_d.shape[0] if batched else 1
    prefix = [B, -1] if batched else [-1]
    dev = rays_o.device
    o_flat = _lib.f32c(rays_o.reshape(-1, 3))
    d_flat = _lib.f32c(rays_d.reshape(-1, 3))
    f = dict(dtype=torch.float32, device=dev)
    alpha, beta = model.forward_ab()
    N_init, M = N_samples * 4, N_samples + N_importance
    outs = []
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        step = int(rayschunk) * B
        for i0 in range(0, o_flat.shape[0], step):
            ro, rd = o_flat[i0:i0 + step], d_flat[i0:i0 + step]
            R = ro.shape[0]
            with torch.no_grad():
                dirs, fars = torch.empty(R, 3, **f), torch.empty(R, **f)
                d_init, pts_init = torch.empty(R, N_init, **f), torch.empty(R, N_init, 3, **f)
                miss = torch.zeros(1, dtype=torch.int32, device=dev)
                _lib.check(lib.nr_volsdf_ray_setup(_lib.ptr(ro), _lib.ptr(rd), R, float(near), float(far),
                                                   float(obj_bounding_radius) if use_nerfplusplus else -1.0, N_init,
                                                   _lib.ptr(dirs), _lib.ptr(fars), _lib.ptr(miss), _lib.ptr(d_init), N_init,
                                                   _lib.ptr(pts_init), st), "volsdf_ray_setup")
                d_fine, beta_map, iter_usage = volsdf.fine_sample(
                    lambda p: volsdf._surface_sdf(model, p), d_init, ro, dirs, alpha, beta, fars, eps=epsilon,
                    max_iter=max_upsample_steps, max_bisection=max_bisection_steps, final_N_importance=N_importance,
                    N_up=N_samples * 4, perturb=perturb)
                d_all, pts = torch.empty(R, M, **f), torch.empty(R, M, 3, **f)
                _lib.check(lib.nr_volsdf_merge(_lib.ptr(ro), _lib.ptr(dirs), _lib.ptr(fars), R, float(near), N_samples,
                                               _lib.ptr(d_fine), N_importance, _lib.ptr(d_all), _lib.ptr(pts), st), "merge")
            radiances, sdf, nablas = model.forward(pts, dirs.unsqueeze(-2).expand(R, M, 3))   # volsdf.py:450
            sigma = volsdf.sdf_to_sigma(sdf, alpha, beta)
            M_in = M
            if use_nerfplusplus:                                                           # volsdf.py:456-475
                t = torch.linspace(0, 1, N_outside + 2)[..., 1:-1].float().to(dev)
                rs = (obj_bounding_radius / torch.flip(t, dims=[-1])).expand([R, N_outside])
                if perturb:
                    mids = .5 * (rs[..., 1:] + rs[..., :-1])
                    upper = torch.cat([mids, rs[..., -1:]], -1)
                    lower = torch.cat([rs[..., :1], mids], -1)
                    rs = lower + (upper - lower) * torch.rand(upper.shape).float().to(dev)
                with torch.no_grad():
                    d_out = rend_util.get_dvals_from_radius(ro, dirs, rs)
                    pts_out = ro[..., None, :] + dirs[..., None, :] * d_out[..., :, None]
                    x_out = torch.cat([pts_out / rs[..., None], 1. / rs[..., None]], dim=-1)
                sigma_out, radiance_out = model.nerf_outside.forward(x_out, dirs.unsqueeze(-2).expand(R, N_outside, 3))
                d_all = torch.cat([d_all, d_out], dim=-1)
                sigma = torch.cat([sigma, sigma_out], dim=-1)
                radiances = torch.cat([radiances, radiance_out], dim=-2)
            delta_i = d_all[..., 1:] - d_all[..., :-1]
            p_i = torch.exp(-F.relu(sigma[..., :-1] * delta_i))
            tau_i = (1 - p_i + 1e-10) * exclusive_cumprod(p_i)
            rgb = torch.sum(tau_i[..., None] * radiances[..., :-1, :], dim=-2)
            depth = torch.sum(tau_i / (tau_i.sum(-1, keepdim=True) + 1e-10) * d_all[..., :-1], dim=-1)
            acc = torch.sum(tau_i, -1)
            if white_bkgd:
                rgb = rgb + (1.0 - acc[..., None])
            ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
            if calc_normal:
                nm = F.normalize(nablas, dim=-1)
                n_pts = min(tau_i.shape[-1], nm.shape[-2])                                  # volsdf.py:510-513
                ret_i['normals_volume'] = (nm[..., :n_pts, :] * tau_i[..., :n_pts, None]).sum(dim=-2)
            if detailed_output:
                ret_i.update(implicit_surface=sdf, implicit_nablas=nablas, radiance=radiances, alpha=1.0 - p_i, p_i=p_i,
                             visibility_weights=tau_i, d_vals=d_all, sigma=sigma, beta_map=beta_map, iter_usage=iter_usage)
                if use_nerfplusplus:
                    ret_i.update(sigma_out=sigma_out, radiance_out=radiance_out)
            outs.append(ret_i)
    ret = OrderedDict()
    for k in outs[0].keys():
        v = outs[0][k] if len(outs) == 1 else torch.cat([o[k] for o in outs], 0)
        ret[k] = v.reshape(*prefix, *v.shape[1:]) if batched else v
    return ret['rgb'], ret['depth_volume'], ret


def unisurf_render_train(rays_o, rays_d, model, batched, calc_normal, logit_tau, rayschunk, netchunk, white_bkgd,
                         near_bypass, far_bypass, detailed_output, radius_of_interest, perturb, interval,
                         too_close_threshold, N_query, N_freespace):
    from ..ray_casting import _root_find
    lib = _lib.get_lib()
    B = rays_d.shape[0] if batched else 1
    dev = rays_o.device
    o_b = _lib.f32c(rays_o.reshape(B, -1, 3))
    d_b = _lib.f32c(rays_d.reshape(B, -1, 3))
    f = dict(dtype=torch.float32, device=dev)
    M, N_steps, N_secant = N_query + N_freespace, 256, 8
    nan = float("nan")
    per_batch = []
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for b in range(B):
            outs = []
            for i0 in range(0, o_b.shape[1], int(rayschunk)):
                ro, rd = o_b[b, i0:i0 + rayschunk].contiguous(), d_b[b, i0:i0 + rayschunk].contiguous()
                R = ro.shape[0]
                with torch.no_grad():
                    dirs, near, far = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
                    pts_prop = torch.empty(R, N_steps, 3, **f)
                    _lib.check(lib.nr_unisurf_ray_setup(
                        _lib.ptr(ro), _lib.ptr(rd), R, float(radius_of_interest), nan if near_bypass is None else float(near_bypass),
                        nan if far_bypass is None else float(far_bypass), N_steps, _lib.ptr(dirs), _lib.ptr(near),
                        _lib.ptr(far), _lib.ptr(pts_prop), st), "unisurf_ray_setup")
                    state, mask, msc, m0 = _root_find(model.implicit_surface.forward, ro, dirs, near, far, pts_prop, N_steps,
                                                      logit_tau, N_secant)
                    u_int = torch.rand([R, N_query], device=dev) if perturb else None
                    u_free = torch.rand([R, N_freespace], device=dev) if perturb else None
                    depth_s, surf_pts = torch.empty(R, **f), torch.empty(R, 3, **f)
                    d_all, pts = torch.empty(R, M, **f), torch.empty(R, M, 3, **f)
                    _lib.check(lib.nr_unisurf_sample(
                        _lib.ptr(ro), _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), _lib.ptr(state), _lib.ptr(mask),
                        _lib.ptr(msc), _lib.ptr(m0), R, float(interval), float(too_close_threshold), N_query, N_freespace,
                        _lib.ptr(u_int), _lib.ptr(u_free), _lib.ptr(depth_s), _lib.ptr(surf_pts), _lib.ptr(d_all),
                        _lib.ptr(pts), st), "unisurf_sample")
                # UNISURF.forward on [1, chunk, 3] slices: same layout (and same chunk-wide F.normalize) as the
                # reference's batchify_query(dim_batchify=1) (unisurf.py:214, train_util.py:23-71)
                flat_pts = pts.reshape(1, -1, 3)
                flat_views = dirs.unsqueeze(-2).expand(R, M, 3).reshape(1, -1, 3)
                rad_l, sdf_l, nab_l = [], [], []
                for j0 in range(0, R * M, int(netchunk)):
                    r_, s_, n_ = model.forward(flat_pts[:, j0:j0 + netchunk], flat_views[:, j0:j0 + netchunk])
                    rad_l.append(r_); sdf_l.append(s_); nab_l.append(n_)
                radiances = torch.cat(rad_l, 1).reshape(R, M, 3)
                logits = torch.cat(sdf_l, 1).reshape(R, M)
                nablas = torch.cat(nab_l, 1).reshape(R, M, 3)
                alpha = model.get_opacity_from_surface(logits)
                w = alpha * exclusive_cumprod(1.0 - alpha + 1e-10)
                rgb = torch.sum(w[..., None] * radiances, -2)
                depth = torch.sum(w / (w.sum(-1, keepdim=True) + 1e-10) * d_all, -1)
                acc = torch.sum(w, -1)
                if white_bkgd:
                    rgb = rgb + (1.0 - acc[..., None])
                ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
                if calc_normal:
                    ret_i['normals_volume'] = (F.normalize(nablas, dim=-1) * w[..., None]).sum(dim=-2)
                if detailed_output:
                    ret_i.update(surface_points=surf_pts, mask_surface=mask.bool(), depth_surface=depth_s, radiance=radiances,
                                 implicit_surface=logits, implicit_nablas=nablas, alpha=alpha, visibility_weights=w)
                outs.append(ret_i)
            per_batch.append(OrderedDict((k, outs[0][k] if len(outs) == 1 else torch.cat([o_[k] for o_ in outs], 0))
                                         for k in outs[0].keys()))
    ret = OrderedDict()
    for k in per_batch[0].keys():
        ret[k] = torch.stack([pb[k] for pb in per_batch], 0) if batched else per_batch[0][k]
    return ret['rgb'], ret['depth_volume'], ret
