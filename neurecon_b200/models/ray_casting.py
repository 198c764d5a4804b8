"""Drop-in for the hot-path part of the reference's ``models/ray_casting.py``:
``root_finding_surface_points`` (:35-160) with ``run_secant_method`` (:11-30)."""
import numpy as np
import torch

from .. import _lib


def _root_find(surface_query_fn, o, dirs, near, far, pts_prop, N_steps, logit_tau, N_secant_steps):
    """Shared by the public function and unisurf.volume_render: flat [R,...] tensors in, returns
    (state [5,R], mask, mask_sign_change, mask_0_free) as device tensors."""
    lib = _lib.get_lib()
    R, dev = o.shape[0], o.device
    st = _lib.stream_ptr(dev)
    val = _lib.f32c(surface_query_fn(pts_prop).reshape(R, N_steps))
    state = torch.empty(5, R, dtype=torch.float32, device=dev)
    mask = torch.empty(R, dtype=torch.uint8, device=dev)
    msc, m0 = torch.empty_like(mask), torch.empty_like(mask)
    pts_pred = torch.empty(R, 3, dtype=torch.float32, device=dev)
    _lib.check(lib.nr_unisurf_first_crossing(_lib.ptr(val), _lib.ptr(o), _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), R,
                                             N_steps, float(logit_tau), _lib.ptr(state), _lib.ptr(mask), _lib.ptr(msc),
                                             _lib.ptr(m0), _lib.ptr(pts_pred), st), "unisurf_first_crossing")
    for _ in range(N_secant_steps):
        f_mid = _lib.f32c(surface_query_fn(pts_pred).reshape(R))
        _lib.check(lib.nr_unisurf_secant_step(_lib.ptr(f_mid), float(logit_tau), _lib.ptr(o), _lib.ptr(dirs),
                                              _lib.ptr(mask), R, _lib.ptr(state), _lib.ptr(pts_pred), st),
                   "unisurf_secant_step")
    return state, mask, msc, m0


def root_finding_surface_points(surface_query_fn, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={},
                                N_steps=256, logit_tau=0.0, method='secant', N_secant_steps=8, fill_inf=True):
    """ray_casting.py:35-160.  rays_o / rays_d: [(B), N_rays, 3] (rays_d normalised); near / far: float or
    [(B), N_rays].  Returns (d_pred_out, pt_pred, mask, mask_sign_change)."""
    if method != 'secant':
        raise NotImplementedError("only method='secant' (every shipped config) is built")
    _lib.require_cuda(rays_o, rays_d)
    with torch.no_grad():
        prefix = rays_o.shape[:-1]
        o = _lib.f32c(rays_o.detach().reshape(-1, 3))
        d = _lib.f32c(rays_d.detach().reshape(-1, 3))
        R, dev = o.shape[0], o.device
        f = dict(dtype=torch.float32, device=dev)
        nr = (near.detach().float().reshape(R) if torch.is_tensor(near) else torch.full((R,), float(near), **f)).contiguous()
        fr = (far.detach().float().reshape(R) if torch.is_tensor(far) else torch.full((R,), float(far), **f)).contiguous()
        t = torch.linspace(0., 1., N_steps, device=dev)[None, :]
        d_prop = nr[:, None] * (1 - t) + fr[:, None] * t
        pts_prop = o.unsqueeze(-2) + d_prop.unsqueeze(-1) * d.unsqueeze(-2)
        with torch.cuda.device(dev):
            state, mask, msc, m0 = _root_find(surface_query_fn, o, d, nr, fr, pts_prop, N_steps, logit_tau,
                                              N_secant_steps)
        mask, msc, m0 = mask.bool(), msc.bool(), m0.bool()
        d_pred = state[4]
        pt_pred = torch.where(mask[:, None], o + d_pred.unsqueeze(-1) * d, torch.ones(R, 3, **f))
        fill = torch.full_like(fr, float("inf")) if fill_inf else fr
        d_out = torch.where(mask, d_pred, fill)
        d_out = torch.where(m0, d_out, torch.zeros_like(d_out))
    return d_out.reshape(prefix), pt_pred.reshape(*prefix, 3), mask.reshape(prefix), msc.reshape(prefix)


def sphere_tracing_surface_points(implicit_surface, rays_o, rays_d, near=0.0, far=6.0, batched=True, batched_info={},
                                  N_iters=20):
    """ray_casting.py:163-184: N_iters steps of d += sdf(o + d*dir) (rays_d normalised).  Returns (d_preds, pts, mask)."""
    _lib.require_cuda(rays_o, rays_d)
    lib = _lib.get_lib()
    with torch.no_grad():
        prefix = rays_o.shape[:-1]
        o = _lib.f32c(rays_o.detach().reshape(-1, 3))
        d = _lib.f32c(rays_d.detach().reshape(-1, 3))
        R, dev = o.shape[0], o.device
        d_preds = torch.full((R,), float(near), dtype=torch.float32, device=dev)
        mask = torch.ones(R, dtype=torch.uint8, device=dev)
        pts = torch.empty(R, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            val = None
            for _ in range(N_iters + 1):
                _lib.check(lib.nr_sphere_trace_step(_lib.ptr(val), _lib.ptr(o), _lib.ptr(d), float(far), R, _lib.ptr(d_preds),
                                                    _lib.ptr(mask), _lib.ptr(pts), st), "sphere_trace_step")
                if _ == N_iters:
                    break
                val = _lib.f32c(getattr(implicit_surface, "forward_level_set", implicit_surface.forward)(pts).reshape(R))
    return d_preds.reshape(prefix), pts.reshape(*prefix, 3), mask.bool().reshape(prefix)


def surface_render(rays_o, rays_d, model, calc_normal=True, rayschunk=8192, netchunk=1048576, batched=True,
                   use_view_dirs=True, show_progress=False, ray_casting_algo='', ray_casting_cfgs={}, **not_used_kwargs):
    """ray_casting.py:187-263: surface rendering (the reference's "100x faster" mode): find the surface point per
    ray, evaluate ``model.forward`` there.  Returns (colors, depths, extras)."""
    from collections import OrderedDict
    import torch.nn.functional as F
    _lib.require_cuda(rays_o, rays_d)
    if bool(use_view_dirs) != bool(model.radiance_net.use_view_dirs):
        # the reference passes view_dirs=None for use_view_dirs=False, which only a RadianceNet built with
        # use_view_dirs=False accepts (base.py:379-384); here such a net ignores whatever views it is handed
        raise ValueError("use_view_dirs=%r needs a radiance net built with use_view_dirs=%r" % (use_view_dirs, use_view_dirs))
    with torch.no_grad():
        B = rays_d.shape[0] if batched else None
        shape = [B, -1, 3] if batched else [-1, 3]
        dim = 1 if batched else 0
        rays_o = torch.reshape(rays_o, shape).float()
        rays_d = F.normalize(torch.reshape(rays_d, shape).float(), dim=-1)
        colors, depths, nablas, masks = [], [], [], []
        for i in range(0, rays_o.shape[dim], rayschunk):
            ro = rays_o[:, i:i + rayschunk] if batched else rays_o[i:i + rayschunk]
            rd = rays_d[:, i:i + rayschunk] if batched else rays_d[i:i + rayschunk]
            if ray_casting_algo == 'root_finding':
                d_pred, pt_pred, mask, *_ = root_finding_surface_points(model.implicit_surface, ro, rd, batched=batched,
                                                                        **ray_casting_cfgs)
            elif ray_casting_algo == 'sphere_tracing':
                d_pred, pt_pred, mask = sphere_tracing_surface_points(model.implicit_surface, ro, rd, batched=batched,
                                                                      **ray_casting_cfgs)
            else:
                raise NotImplementedError
            color, _, nab = model.forward(pt_pred, rd)
            color = torch.where(mask[..., None], color, torch.zeros_like(color))
            colors.append(color); depths.append(d_pred); nablas.append(nab); masks.append(mask)
        colors, depths = torch.cat(colors, dim), torch.cat(depths, dim)
        nablas, masks = torch.cat(nablas, dim), torch.cat(masks, dim)
        extras = OrderedDict([('implicit_nablas', nablas), ('mask_surface', masks)])
        if calc_normal:
            normals = F.normalize(nablas, dim=-1)
            extras['normals_surface'] = torch.where(masks[..., None], normals, torch.zeros_like(normals))
        return colors, depths, extras
