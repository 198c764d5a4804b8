"""Drop-in for the reference's ``models/frameworks/volsdf.py`` hot path: ``sdf_to_sigma`` (:16-35),
``error_bound`` (:38-74), ``fine_sample`` (:77-272), the ``VolSDF`` module (:274-331),
``volume_render`` (:334-551) and ``SingleRenderer`` (:554-560)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ..base import ImplicitSurface, RadianceNet, query_radiance


class VolSDF(nn.Module):
    """volsdf.py:274-331 (same constructor, parameters and methods)."""

    def __init__(self, beta_init=0.1, speed_factor=1.0, input_ch=3, W_geo_feat=-1, obj_bounding_radius=3.0,
                 use_nerfplusplus=False, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        self.speed_factor = speed_factor
        self.ln_beta = nn.Parameter(data=torch.Tensor([np.log(beta_init) / self.speed_factor]), requires_grad=True)
        self.use_sphere_bg = not use_nerfplusplus
        self.obj_bounding_radius = obj_bounding_radius
        self.implicit_surface = ImplicitSurface(
            W_geo_feat=W_geo_feat, input_ch=input_ch, obj_bounding_size=obj_bounding_radius, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)
        if use_nerfplusplus:
            from ..base import NeRF
            self.nerf_outside = NeRF(input_ch=4, multires=10, multires_view=4, use_view_dirs=True)

    def forward_ab(self):
        beta = torch.exp(self.ln_beta * self.speed_factor)
        return 1. / beta, beta

    def forward_surface(self, x):
        sdf = self.implicit_surface.forward(x)
        if self.use_sphere_bg:
            return torch.min(sdf, self.obj_bounding_radius - x.norm(dim=-1))
        return sdf

    def forward_surface_with_nablas(self, x):
        sdf, nablas, h = self.implicit_surface.forward_with_nablas(x)
        if self.use_sphere_bg:
            d_bg = self.obj_bounding_radius - x.norm(dim=-1)
            sdf = torch.where(d_bg < sdf, d_bg, sdf)
        return sdf, nablas, h

    def forward(self, x, view_dirs):
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            radiances, sdf, nablas = query_radiance(self.implicit_surface, self.radiance_net, x, view_dirs)
            if self.use_sphere_bg:
                sdf = torch.min(sdf, self.obj_bounding_radius - x.norm(dim=-1))
            return radiances, sdf, nablas
        sdf, nablas, geometry_feature = self.forward_surface_with_nablas(x)
        radiances = self.radiance_net.forward(x, view_dirs, nablas, geometry_feature)
        return radiances, sdf, nablas


# --------------------------------------------------------------------------------------------
# tensor helpers of the reference API
# --------------------------------------------------------------------------------------------
def sdf_to_sigma(sdf, alpha, beta):
    """volsdf.py:16-35 (Laplace-CDF density).  Thin torch composition for API users; the renderer
    evaluates it inside nr_volsdf_composite / nr_volsdf_fine_iter."""
    e = 0.5 * torch.exp(-torch.abs(sdf) / beta)
    return alpha * torch.where(sdf >= 0, e, 1 - e)


def _as_ray_param(v, R, dev):
    """scalar / 0-d / [R,1] / [R] -> (contiguous fp32 tensor, stride)"""
    if not torch.is_tensor(v):
        v = torch.tensor(float(v), device=dev)
    v = v.detach().float().to(dev)
    if v.numel() == 1:
        return v.reshape(1).contiguous(), 0
    assert v.numel() == R, "per-ray parameter must have one entry per ray"
    return v.reshape(R).contiguous(), 1


def error_bound(d_vals, sdf, alpha, beta):
    """volsdf.py:38-74: [..., N_pts] -> [..., N_pts-1] error bounds (NaN -> inf)."""
    _lib.require_cuda(d_vals, sdf)
    lib = _lib.get_lib()
    prefix, M = d_vals.shape[:-1], d_vals.shape[-1]
    d = _lib.f32c(d_vals.detach().reshape(-1, M))
    s = _lib.f32c(sdf.detach().reshape(-1, M))
    R, dev = d.shape[0], d.device
    a, a_st = _as_ray_param(alpha, R, dev)
    b, b_st = _as_ray_param(beta, R, dev)
    out = torch.empty(R, M - 1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.nr_volsdf_error_bound(_lib.ptr(d), _lib.ptr(s), R, M, _lib.ptr(a), a_st, _lib.ptr(b), b_st,
                                             _lib.ptr(out), None, _lib.stream_ptr(dev)), "volsdf_error_bound")
    return out.reshape(*prefix, M - 1)


def fine_sample(implicit_surface_fn, init_dvals, rays_o, rays_d, alpha_net, beta_net, far, eps=0.1, max_iter: int = 5,
                max_bisection: int = 10, final_N_importance: int = 64, N_up: int = 128, perturb=True,
                early_exit=True):
    """volsdf.py:77-272 (error-bounded beta iteration).  Same arguments and returns
    (d_fine [..., N], beta [..., 1], iter_usage [...]); ``implicit_surface_fn(pts[..., 3]) -> sdf``.
    The per-ray state machine runs on the device; ``early_exit`` reads one flag per iteration to skip the
    remaining network queries once every ray is finished (set False for a sync-free, graph-capturable run)."""
    _lib.require_cuda(init_dvals, rays_o, rays_d)
    lib = _lib.get_lib()
    if torch.cuda.is_current_stream_capturing():
        early_exit = False                   # a CUDA-graph capture cannot read the flag: the sync-free form (every ray, every iteration)
    prefix, N0 = init_dvals.shape[:-1], init_dvals.shape[-1]
    o = _lib.f32c(rays_o.detach().reshape(-1, 3))
    dirs = _lib.f32c(rays_d.detach().reshape(-1, 3))
    R, dev = o.shape[0], o.device
    f = dict(dtype=torch.float32, device=dev)
    cap = N0 + max_iter * N_up
    d_buf, sdf_buf = torch.empty(R, cap, **f), torch.empty(R, cap, **f)
    d_buf[:, :N0] = init_dvals.detach().reshape(R, N0)
    fars, _ = _as_ray_param(far, R, dev)
    fars = fars.expand(R).contiguous()
    a_net = alpha_net.detach().float().reshape(1).contiguous() if torch.is_tensor(alpha_net) else torch.tensor([float(alpha_net)], **f)
    b_net = beta_net.detach().float().reshape(1).contiguous() if torch.is_tensor(beta_net) else torch.tensor([float(beta_net)], **f)
    beta, status = torch.zeros(R, **f), torch.zeros(R, dtype=torch.int32, device=dev)
    iter_usage, beta_map, d_fine = torch.zeros(R, **f), torch.zeros(R, **f), torch.zeros(R, final_N_importance, **f)
    d_new = [torch.zeros(R, N_up, **f), torch.zeros(R, N_up, **f)]
    pts_new = torch.empty(R, N_up, 3, **f)
    u_final = None if not perturb else torch.rand(R, final_N_importance, device=dev)
    with torch.cuda.device(dev), torch.no_grad():
        st = _lib.stream_ptr(dev)
        pts0 = o[:, None, :] + dirs[:, None, :] * d_buf[:, :N0, None]
        sdf_new = _lib.f32c(implicit_surface_fn(pts0).reshape(R, N0))
        for it in range(max_iter + 1):
            n_new = N0 if it == 0 else N_up
            m_cur = 0 if it == 0 else N0 + (it - 1) * N_up
            _lib.check(lib.nr_volsdf_fine_iter(
                _lib.ptr(o), _lib.ptr(dirs), _lib.ptr(fars), R, _lib.ptr(d_buf), _lib.ptr(sdf_buf), cap, m_cur,
                _lib.ptr(sdf_new), n_new, None if it == 0 else _lib.ptr(d_new[(it - 1) & 1]), _lib.ptr(a_net),
                _lib.ptr(b_net), float(eps), it, max_iter, max_bisection, N_up, final_N_importance, _lib.ptr(u_final),
                N0, _lib.ptr(beta), _lib.ptr(status), _lib.ptr(iter_usage), _lib.ptr(beta_map), _lib.ptr(d_fine),
                _lib.ptr(d_new[it & 1]), _lib.ptr(pts_new), st), "volsdf_fine_iter")
            if it == max_iter:
                break
            if early_exit:
                # one host read per iteration: WHICH rays are still refining.  Only their proposals go through the network
                # -- the reference's boolean-mask gathers (volsdf.py:151-264); finished rays ignore sdf_new in the kernel.
                # With every ray active (or early_exit=False: sync-free, graph-capturable) the whole batch is queried.
                active = torch.nonzero(status == 0).squeeze(1)
                if active.numel() == 0:
                    break
                if active.numel() < R:
                    sdf_act = _lib.f32c(implicit_surface_fn(pts_new.index_select(0, active)).reshape(-1, N_up))
                    sdf_new = torch.zeros(R, N_up, **f) if (sdf_new is None or sdf_new.shape != (R, N_up)) else sdf_new
                    sdf_new.index_copy_(0, active, sdf_act)
                    continue
            sdf_new = _lib.f32c(implicit_surface_fn(pts_new).reshape(R, N_up))
    return (d_fine.reshape(*prefix, final_N_importance), beta_map.reshape(*prefix, 1), iter_usage.reshape(*prefix))


def _surface_sdf(model, pts):
    """VolSDF.forward_surface (volsdf.py:310-315) with the bounding-sphere min fused in place."""
    sdf = model.implicit_surface.forward(pts)
    if model.use_sphere_bg:
        lib = _lib.get_lib()
        p = _lib.f32c(pts.reshape(-1, 3))
        _lib.check(lib.nr_sphere_min(_lib.ptr(p), _lib.ptr(sdf), p.shape[0], float(model.obj_bounding_radius),
                                     _lib.stream_ptr(p.device)), "sphere_min")
    return sdf


def volume_render(
        rays_o,
        rays_d,
        model: VolSDF,

        near=0.0,
        far=6.0,
        obj_bounding_radius=3.0,

        batched=False,
        batched_info={},

        # render algorithm config
        calc_normal=False,
        use_view_dirs=True,
        rayschunk=65536,
        netchunk=1048576,
        white_bkgd=False,
        use_nerfplusplus=False,

        # render function config
        detailed_output=True,
        show_progress=False,

        # sampling related
        perturb=False,
        N_samples=128,
        N_importance=64,
        N_outside=32,
        max_upsample_steps=5,
        max_bisection_steps=10,
        epsilon=0.1,

        # determinism hook for parity tests: {"d_all": [R, N_samples + N_importance] sorted depths, "beta_map": [R,1],
        # "iter_usage": [R]} replaces the error-bounded sampler's result
        samples_bypass=None,
        **dummy_kwargs):
    """volsdf.py:334-551.  rays_o / rays_d: [(B,) N_rays, 3].  Returns (rgb, depth_volume, ret)."""
    if bool(use_view_dirs) != bool(model.radiance_net.use_view_dirs):
        # the reference passes view_dirs=None for use_view_dirs=False, which only a RadianceNet built with
        # use_view_dirs=False accepts (base.py:379-384); here such a net ignores whatever views it is handed
        raise ValueError("use_view_dirs=%r needs a radiance net built with use_view_dirs=%r" % (use_view_dirs, use_view_dirs))
    _lib.require_cuda(rays_o, rays_d)
    # training (volsdf.py:578: Trainer.forward renders under autograd): the error-bounded sampler stays no_grad as in the
    # reference (volsdf.py:77 fine_sample is @torch.no_grad via its callers, the depths are detached), the network query
    # goes through VolSDF.forward (models/autograd.py), density + compositing through VolsdfComposite
    train = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    from ..composite import VolsdfComposite
    lib = _lib.get_lib()
    B = rays_d.shape[0] if batched else 1
    prefix = [B, -1] if batched else [-1]
    dev = rays_o.device
    o_flat = _lib.f32c(rays_o.reshape(-1, 3))
    d_flat = _lib.f32c(rays_d.reshape(-1, 3))
    n_total = o_flat.shape[0]
    f = dict(dtype=torch.float32, device=dev)
    alpha_g, beta_g = model.forward_ab()                      # with the graph to ln_beta when training
    alpha_t, beta_t = alpha_g.detach().float().contiguous(), beta_g.detach().float().contiguous()
    N_init = N_samples * 4
    M_in = N_samples + N_importance
    M_out = N_outside if use_nerfplusplus else 0
    miss = torch.zeros(1, dtype=torch.int32, device=dev)      # rays that miss the sphere / radii inside the ray's closest point

    outs = []
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        step = int(rayschunk) * B
        for i0 in range(0, n_total, step):
            ro, rd = o_flat[i0:i0 + step], d_flat[i0:i0 + step]
            R = ro.shape[0]
            with torch.no_grad():
                dirs, fars = torch.empty(R, 3, **f), torch.empty(R, **f)
                d_init, pts_init = torch.empty(R, N_init, **f), torch.empty(R, N_init, 3, **f)
                _lib.check(lib.nr_volsdf_ray_setup(
                    _lib.ptr(ro), _lib.ptr(rd), R, float(near), float(far),
                    float(obj_bounding_radius) if use_nerfplusplus else -1.0, N_init, _lib.ptr(dirs), _lib.ptr(fars),
                    _lib.ptr(miss), _lib.ptr(d_init), N_init, _lib.ptr(pts_init), st), "volsdf_ray_setup")
                if samples_bypass is None:
                    d_fine, beta_map, iter_usage = fine_sample(
                        lambda p: _surface_sdf(model, p), d_init, ro, dirs, alpha_t, beta_t, fars, eps=epsilon,
                        max_iter=max_upsample_steps, max_bisection=max_bisection_steps, final_N_importance=N_importance,
                        N_up=N_samples * 4, perturb=perturb)
                    d_in, pts = torch.empty(R, M_in, **f), torch.empty(R, M_in, 3, **f)
                    _lib.check(lib.nr_volsdf_merge(_lib.ptr(ro), _lib.ptr(dirs), _lib.ptr(fars), R, float(near), N_samples,
                                                   _lib.ptr(d_fine), N_importance, _lib.ptr(d_in), _lib.ptr(pts), st),
                               "volsdf_merge")
                else:
                    d_in = _lib.f32c(samples_bypass["d_all"][i0:i0 + step].to(dev))
                    pts = (ro[:, None, :] + dirs[:, None, :] * d_in[..., None]).contiguous()
                    beta_map = samples_bypass["beta_map"][i0:i0 + step].to(dev)
                    iter_usage = samples_bypass["iter_usage"][i0:i0 + step].to(dev)
            views = dirs.unsqueeze(-2).expand(R, M_in, 3)
            if train:
                radiances, sdf, nablas = model.forward(pts, views)                         # volsdf.py:450
            else:
                with torch.no_grad():
                    radiances, sdf, nablas = query_radiance(model.implicit_surface, model.radiance_net, pts, views)
                    if model.use_sphere_bg:
                        _lib.check(lib.nr_sphere_min(_lib.ptr(pts), _lib.ptr(sdf), R * M_in, float(model.obj_bounding_radius),
                                                     st), "sphere_min")
            sigma_out = radiance_out = d_out = None
            if use_nerfplusplus:
                # inverted-sphere samples (volsdf.py:456-467) in one launch; the jitter uniforms come from the CPU generator
                # like the reference's (device generator while a CUDA graph is being captured)
                with torch.no_grad():
                    u = None
                    if perturb:
                        u = (torch.rand([R, N_outside], device=dev) if torch.cuda.is_current_stream_capturing()
                             else torch.rand([R, N_outside]).float().to(dev))
                    d_out, x_out = torch.empty(R, N_outside, **f), torch.empty(R, N_outside, 4, **f)
                    _lib.check(lib.nr_volsdf_outside_points(_lib.ptr(ro), _lib.ptr(dirs), R, float(obj_bounding_radius),
                                                            N_outside, _lib.ptr(u), _lib.ptr(d_out), _lib.ptr(x_out),
                                                            _lib.ptr(miss), st), "volsdf_outside_points")
                sigma_out, radiance_out = model.nerf_outside.forward(x_out, dirs.unsqueeze(-2).expand(R, N_outside, 3))
            rgb, depth, acc, normals, sigma_all, p_i, tau = VolsdfComposite.apply(
                sdf, nablas if calc_normal else None, radiances, d_in, alpha_g if train else alpha_t,
                beta_g if train else beta_t, sigma_out, radiance_out, d_out, bool(white_bkgd), bool(calc_normal),
                bool(detailed_output))
            ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
            if calc_normal:
                ret_i['normals_volume'] = normals
            if detailed_output:
                ret_i['implicit_surface'] = sdf
                ret_i['implicit_nablas'] = nablas
                ret_i['radiance'] = radiances if not use_nerfplusplus else torch.cat([radiances, radiance_out], dim=-2)
                ret_i['alpha'] = 1.0 - p_i
                ret_i['p_i'] = p_i
                ret_i['visibility_weights'] = tau
                ret_i['d_vals'] = d_in if not use_nerfplusplus else torch.cat([d_in, d_out], dim=-1)
                ret_i['sigma'] = sigma_all
                ret_i['beta_map'] = beta_map
                ret_i['iter_usage'] = iter_usage
                if use_nerfplusplus:
                    ret_i['sigma_out'] = sigma_out
                    ret_i['radiance_out'] = radiance_out
            outs.append(ret_i)
    if use_nerfplusplus and not torch.cuda.is_current_stream_capturing():
        assert int(miss.item()) == 0, "every ray must intersect the bounding sphere (volsdf.py:404-405, rend_util.py:225)"

    ret = OrderedDict()
    for k in outs[0].keys():
        v = outs[0][k] if len(outs) == 1 else torch.cat([o_[k] for o_ in outs], 0)
        ret[k] = v.reshape(*prefix, *v.shape[1:]) if batched else v
    return ret['rgb'], ret['depth_volume'], ret


class SingleRenderer(nn.Module):
    """volsdf.py:554-560."""

    def __init__(self, model: VolSDF):
        super().__init__()
        self.model = model

    def forward(self, rays_o, rays_d, **kwargs):
        return volume_render(rays_o, rays_d, self.model, **kwargs)


def __getattr__(name):
    """``Trainer`` and ``get_model`` (volsdf.py of the reference) live in frameworks/trainers.py; resolved lazily because
    that module imports this one."""
    if name in ("Trainer", "get_model"):
        from . import trainers
        return {"Trainer": trainers.VolsdfTrainer, "get_model": trainers.get_model_volsdf}[name]
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
