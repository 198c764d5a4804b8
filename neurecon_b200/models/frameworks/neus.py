"""Drop-in for the reference's ``models/frameworks/neus.py`` hot path:
``cdf_Phi_s`` / ``sdf_to_alpha`` / ``sdf_to_w`` / ``alpha_to_w`` (:21-70), the ``NeuS`` module
(:72-115), ``volume_render`` (:118-397) and ``SingleRenderer`` (:399-405).

``volume_render`` keeps the reference's signature, defaults and return structure; per ray chunk
it issues a fixed sequence of library calls (no host syncs, CUDA-graph capturable):

    ray_setup -> [sdf MLP, upsample_step] x (N_upsample_iters + 1) -> sdf+nabla MLP (pts)
              -> sdf+nabla+feature+radiance MLP (mid points) -> composite
"""
import math
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ..base import ImplicitSurface, RadianceNet, query_radiance


# --------------------------------------------------------------------------------------------
# small tensor helpers of the reference API (neus.py:21-70); thin torch compositions, used by
# debug tools -- the renderer itself goes through nr_neus_composite.
# --------------------------------------------------------------------------------------------
def cdf_Phi_s(x, s):
    return torch.sigmoid(x * s)


def sdf_to_alpha(sdf, s):
    cdf = cdf_Phi_s(sdf, s)
    alpha = ((cdf[..., :-1] - cdf[..., 1:]) / (cdf[..., :-1] + 1e-10)).clamp_min(0)
    return cdf, alpha


def alpha_to_w(alpha):
    from ..autograd import exclusive_cumprod          # torch.cumprod's backward syncs with the host
    return alpha * exclusive_cumprod(1.0 - alpha + 1e-10)


def sdf_to_w(sdf, s):
    cdf, alpha = sdf_to_alpha(sdf, s)
    return cdf, alpha, alpha_to_w(alpha)


class NeuS(nn.Module):
    """neus.py:72-115 (same constructor, parameters and methods)."""

    def __init__(self, variance_init=0.05, speed_factor=1.0, input_ch=3, W_geo_feat=-1, use_outside_nerf=False,
                 obj_bounding_radius=1.0, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        self.ln_s = nn.Parameter(data=torch.Tensor([-np.log(variance_init) / speed_factor]), requires_grad=True)
        self.speed_factor = speed_factor
        self.implicit_surface = ImplicitSurface(
            W_geo_feat=W_geo_feat, input_ch=input_ch, obj_bounding_size=obj_bounding_radius, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)
        if use_outside_nerf:
            from ..base import NeRF
            self.nerf_outside = NeRF(input_ch=4, multires=10, multires_view=4, use_view_dirs=True)

    def _needs_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def forward_radiance(self, x, view_dirs):
        if not self._needs_grad():
            return query_radiance(self.implicit_surface, self.radiance_net, x, view_dirs)[0]
        _, nablas, geometry_feature = self.implicit_surface.forward_with_nablas(x)
        return self.radiance_net.forward(x, view_dirs, nablas, geometry_feature)

    def forward_s(self):
        return torch.exp(self.ln_s * self.speed_factor)

    def forward(self, x, view_dirs):
        if not self._needs_grad():
            return query_radiance(self.implicit_surface, self.radiance_net, x, view_dirs)
        sdf, nablas, geometry_feature = self.implicit_surface.forward_with_nablas(x)
        radiances = self.radiance_net.forward(x, view_dirs, nablas, geometry_feature)
        return radiances, sdf, nablas


def _composite(sdf, nablas, radiances, d_mid, s, white_bkgd, calc_normal, detailed):
    """nr_neus_composite on [R, M] per-sample tensors (differentiable: models/composite.py)."""
    from ..composite import NeusComposite
    return NeusComposite.apply(sdf, nablas if calc_normal else None, radiances, d_mid, s, None, None, None, None, 0.0, 0,
                               bool(white_bkgd), bool(calc_normal), bool(detailed))[:7]


def _outside_points(rays_o, dirs, far, d_mid, N_outside, perturb):
    """neus.py:303-318: d_vals [R, M-1+N_outside] = cat(d_mid, far / flip(linspace)) (stratified jitter if perturb) and
    the inverted-sphere inputs x_out [R, M-1+N_outside, 4] of NeRF.forward."""
    lib = _lib.get_lib()
    R, M1 = d_mid.shape
    dev = d_mid.device
    f = dict(dtype=torch.float32, device=dev)
    T = M1 + N_outside
    u = None
    if perturb:
        # the reference draws on the CPU generator (neus.py:310); a pageable host-to-device copy cannot be captured into
        # a CUDA graph, so a capturing stream draws on the device instead (different stream of random numbers)
        if torch.cuda.is_current_stream_capturing():
            u = torch.rand([R, N_outside], device=dev)
        else:
            u = torch.rand([R, N_outside]).float().to(dev)
    d_vals, x_out = torch.empty(R, T, **f), torch.empty(R, T, 4, **f)
    _lib.check(lib.nr_neus_outside_points(_lib.ptr(rays_o), _lib.ptr(dirs), _lib.ptr(far), _lib.ptr(d_mid), R, M1,
                                          N_outside, _lib.ptr(u), _lib.ptr(d_vals), _lib.ptr(x_out), _lib.stream_ptr(dev)),
               "neus_outside_points")
    return d_vals, x_out


def _forced_samples(rays_o, rays_d_raw, d_all, obj_bounding_radius, near_bypass, far_bypass):
    """Parity-test hook (``samples_bypass``): the tensors ``_upsample`` returns, for GIVEN sorted depths d_all [R, M]
    (neus.py:284-288: points, mid depths, mid points)."""
    lib = _lib.get_lib()
    R, dev = rays_o.shape[0], rays_o.device
    f = dict(dtype=torch.float32, device=dev)
    dirs, near, far = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
    d0, p0 = torch.empty(R, 2, **f), torch.empty(R, 2, 3, **f)
    nan = float("nan")
    _lib.check(lib.nr_neus_ray_setup(
        _lib.ptr(rays_o), _lib.ptr(rays_d_raw), R, float(obj_bounding_radius),
        nan if near_bypass is None else float(near_bypass), nan if far_bypass is None else float(far_bypass), 2,
        _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), _lib.ptr(d0), _lib.ptr(p0), _lib.stream_ptr(dev)), "neus_ray_setup")
    d_all = _lib.f32c(d_all.to(dev))
    pts = rays_o[:, None, :] + dirs[:, None, :] * d_all[..., None]
    d_mid = 0.5 * (d_all[..., 1:] + d_all[..., :-1])
    pts_mid = rays_o[:, None, :] + dirs[:, None, :] * d_mid[..., None]
    return dirs, d_all, pts.contiguous(), d_mid.contiguous(), pts_mid.contiguous(), far


def _upsample(model, rays_o, rays_d_raw, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance,
              N_upsample_iters, perturb, return_far=False, return_field=False, with_nablas=False):
    """neus.py:184-288 for one flat ray chunk [R,3]: returns dirs, d_all, pts, d_mid, pts_mid [, far]
    [, sdf_all, nablas_all].  ``return_field``: also the sdf at the sorted samples, which the up-sampler has evaluated
    anyway (the reference evaluates it again, neus.py:291); ``with_nablas``: the network queries of the up-sampler also
    produce the normals and the merge carries them along, so (sdf_all, nablas_all) replace that second evaluation."""
    lib = _lib.get_lib()
    R, dev = rays_o.shape[0], rays_o.device
    f = dict(dtype=torch.float32, device=dev)
    st = _lib.stream_ptr(dev)
    n_fine = N_importance // N_upsample_iters
    cap = N_samples + n_fine * N_upsample_iters
    dirs, near, far = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
    d_new, pts_new = torch.empty(R, N_samples, **f), torch.empty(R, N_samples, 3, **f)
    nan = float("nan")
    _lib.check(lib.nr_neus_ray_setup(
        _lib.ptr(rays_o), _lib.ptr(rays_d_raw), R, float(obj_bounding_radius),
        nan if near_bypass is None else float(near_bypass), nan if far_bypass is None else float(far_bypass),
        N_samples, _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), _lib.ptr(d_new), _lib.ptr(pts_new), st),
        "neus_ray_setup")
    d_buf, sdf_buf = torch.empty(R, cap, **f), torch.empty(R, cap, **f)
    nab_buf = torch.empty(R, cap, 3, **f) if with_nablas else None
    pts_all, d_mid, pts_mid = torch.empty(R, cap, 3, **f), torch.empty(R, cap - 1, **f), torch.empty(R, cap - 1, 3, **f)
    m_cur, n_new = 0, N_samples
    with torch.no_grad():
        for it in range(N_upsample_iters + 1):
            nab_new = None
            if with_nablas:
                sdf_new, nab_new, _ = model.implicit_surface._run(pts_new, want_nablas=True, want_feat=False)
            else:
                sdf_new = model.implicit_surface.forward(pts_new)
            last = it == N_upsample_iters
            n_next = 0 if last else n_fine
            u = None
            if not last and perturb:
                u = torch.rand([R, n_next], device=dev)  # same call as rend_util.py:271
            d_next = torch.empty(R, max(n_next, 1), **f)
            pts_next = torch.empty(R, max(n_next, 1), 3, **f)
            _lib.check(lib.nr_neus_upsample_step(
                _lib.ptr(rays_o), _lib.ptr(dirs), R, _lib.ptr(d_buf), _lib.ptr(sdf_buf), cap, m_cur,
                _lib.ptr(d_new), _lib.ptr(sdf_new), n_new, it, n_next, _lib.ptr(u), _lib.ptr(d_next),
                _lib.ptr(pts_next), _lib.ptr(pts_all), _lib.ptr(d_mid), _lib.ptr(pts_mid), _lib.ptr(nab_buf),
                _lib.ptr(nab_new), st), "neus_upsample_step")
            m_cur += n_new
            d_new, pts_new, n_new = d_next, pts_next, n_next
    out = (dirs, d_buf, pts_all, d_mid, pts_mid) + ((far,) if return_far else ())
    if return_field:
        out += (sdf_buf, nab_buf)
    return out


def _upsample_direct(model, rays_o, rays_d_raw, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance,
                     perturb, fixed_s_recp, algo, N_nograd_samples):
    """neus.py:216-243, the two NeRF-like up-samplers: the visibility weights of a FIXED slope 1 / fixed_s_recp on the
    coarse samples ('direct_use') or on N_nograd_samples extra samples ('direct_more'; no gradients there, hence the
    name), ONE inverse-CDF draw of N_importance depths, merged with the coarse ones.  Returns the sorted depths [R, M]."""
    lib = _lib.get_lib()
    R, dev = rays_o.shape[0], rays_o.device
    f = dict(dtype=torch.float32, device=dev)
    st = _lib.stream_ptr(dev)
    nan = float("nan")

    def setup(n):
        dirs, near, far = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
        d, pts = torch.empty(R, n, **f), torch.empty(R, n, 3, **f)
        _lib.check(lib.nr_neus_ray_setup(
            _lib.ptr(rays_o), _lib.ptr(rays_d_raw), R, float(obj_bounding_radius),
            nan if near_bypass is None else float(near_bypass), nan if far_bypass is None else float(far_bypass),
            n, _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), _lib.ptr(d), _lib.ptr(pts), st), "neus_ray_setup")
        return d, pts

    from ...utils import rend_util
    with torch.no_grad():
        d_coarse, pts_coarse = setup(N_samples)
        if algo == 'direct_use':
            d_w, pts_w = d_coarse, pts_coarse
        else:
            d_w, pts_w = setup(int(N_nograd_samples))
        sdf_w = _lib.f32c(model.implicit_surface.forward(pts_w))
        n_w = d_w.shape[1]
        w = torch.empty(R, n_w - 1, **f)
        _lib.check(lib.nr_neus_sdf_to_w(_lib.ptr(sdf_w), 1.0 / float(fixed_s_recp), R, n_w, _lib.ptr(w), st), "neus_sdf_to_w")
        d_fine = rend_util.sample_pdf(d_w, w, N_importance, det=not perturb)
        d_all, _ = torch.sort(torch.cat([d_coarse, d_fine], dim=-1), dim=-1)
    return d_all.contiguous()


def volume_render(
        rays_o,
        rays_d,
        model: NeuS,

        obj_bounding_radius=1.0,

        batched=False,
        batched_info={},

        # render algorithm config
        calc_normal=False,
        use_view_dirs=True,
        rayschunk=65536,
        netchunk=1048576,
        white_bkgd=False,
        near_bypass: Optional[float] = None,
        far_bypass: Optional[float] = None,

        # render function config
        detailed_output=True,
        show_progress=False,

        # sampling related
        perturb=False,
        fixed_s_recp=1 / 64.,
        N_samples=64,
        N_importance=64,
        N_outside=0,

        # upsample related
        upsample_algo='official_solution',
        N_nograd_samples=2048,
        N_upsample_iters=4,

        # determinism hook for parity tests: {"d_all": [R, M] sorted depths} replaces the up-sampler's result
        samples_bypass=None,

        **dummy_kwargs):
    """neus.py:118-397.  rays_o / rays_d: [(B,) N_rays, 3] (rays_d not normalised).
    Returns (rgb, depth_volume, ret) with the reference's ``ret`` keys and shapes."""
    if upsample_algo not in ('official_solution', 'direct_use', 'direct_more'):
        raise NotImplementedError("upsample_algo=%r" % (upsample_algo,))      # neus.py:281
    if N_outside > 0 and not hasattr(model, "nerf_outside"):
        raise ValueError("N_outside > 0 needs a model built with use_outside_nerf=True")
    if bool(use_view_dirs) != bool(model.radiance_net.use_view_dirs):
        # the reference passes view_dirs=None for use_view_dirs=False (neus.py:190-193), which only a RadianceNet built
        # with use_view_dirs=False accepts (base.py:379-384)
        raise ValueError("use_view_dirs=%r needs a radiance net built with use_view_dirs=%r" % (use_view_dirs, use_view_dirs))
    _lib.require_cuda(rays_o, rays_d)
    from ..composite import NeusComposite
    # training (neus.py:440: Trainer.forward renders under autograd): the up-sampler stays no_grad as in the reference
    # (neus.py:214), the two with-grad network queries (neus.py:294,298) go through models/autograd.py, the compositing
    # through NeusComposite -- same kernels forward, hand-written adjoints backward
    train = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    if batched:
        B = rays_d.shape[0]
        prefix = [B, -1]
    else:
        B = 1
        prefix = [-1]
    dev = rays_o.device
    o_flat = _lib.f32c(rays_o.reshape(-1, 3))
    d_flat = _lib.f32c(rays_d.reshape(-1, 3))
    n_total = o_flat.shape[0]
    direct = upsample_algo != 'official_solution'
    M = N_samples + (N_importance if direct else (N_importance // N_upsample_iters) * N_upsample_iters)
    s = model.forward_s() if train else model.forward_s().detach().float().contiguous()

    outs = []
    with torch.cuda.device(dev):
        # the reference chunks along the ray axis (neus.py:385-395); batches are flattened here
        # because rays are independent, and chunk boundaries do not change any value.
        step = int(rayschunk) * B
        for i0 in range(0, n_total, step):
            ro, rd = o_flat[i0:i0 + step], d_flat[i0:i0 + step]
            R = ro.shape[0]
            if samples_bypass is not None or direct:
                d_given = samples_bypass["d_all"][i0:i0 + step] if samples_bypass is not None else _upsample_direct(
                    model, ro, rd, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance, perturb,
                    fixed_s_recp, upsample_algo, N_nograd_samples)
                dirs, d_all, pts, d_mid, pts_mid, far = _forced_samples(
                    ro, rd, d_given, obj_bounding_radius, near_bypass, far_bypass)
            elif train:
                with torch.no_grad():
                    dirs, d_all, pts, d_mid, pts_mid, far = _upsample(
                        model, ro, rd, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance,
                        N_upsample_iters, perturb, return_far=True)
            if train or samples_bypass is not None or direct:
                # neus.py:294,298: forward_with_nablas at the samples and (inside forward_radiance) at the mid points -- ONE
                # network call on both point sets, so that the per-layer training GEMMs run once over 255 points per ray
                n1 = R * M
                sdf_a, nab_a, feat_a = model.implicit_surface.forward_with_nablas(
                    torch.cat([pts.reshape(-1, 3), pts_mid.reshape(-1, 3)], dim=0))
                sdf, nablas = sdf_a[:n1].reshape(R, M), nab_a[:n1].reshape(R, M, 3)
                radiances = model.radiance_net.forward(pts_mid, dirs.unsqueeze(-2).expand(R, M - 1, 3),
                                                       nab_a[n1:].reshape(R, M - 1, 3), feat_a[n1:].reshape(R, M - 1, -1))
            else:
                # sdf (and, when the caller wants normals or per-sample outputs, nablas) at the sorted samples come out of
                # the up-sampler's own network queries: same points, same arithmetic as the reference's second
                # evaluation there
                need_nablas = bool(calc_normal or detailed_output)
                dirs, d_all, pts, d_mid, pts_mid, far, sdf, nablas = _upsample(
                    model, ro, rd, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance,
                    N_upsample_iters, perturb, return_far=True, return_field=True, with_nablas=need_nablas)
                with torch.no_grad():
                    views = dirs.unsqueeze(-2).expand(R, M - 1, 3)
                    radiances, _, _ = query_radiance(model.implicit_surface, model.radiance_net, pts_mid, views)
            sigma_out = radiance_out = None
            d_final = d_mid
            if N_outside > 0:
                with torch.no_grad():
                    d_final, x_out = _outside_points(ro, dirs, far, d_mid, N_outside, perturb)
                sigma_out, radiance_out = model.nerf_outside.forward(
                    x_out, dirs.unsqueeze(-2).expand(R, d_final.shape[-1], 3))
            rgb, depth, acc, normals, cdf, alpha, w, blended = NeusComposite.apply(
                sdf, nablas if calc_normal else None, radiances, d_final, s, sigma_out, radiance_out, ro, dirs,
                float(obj_bounding_radius), int(N_outside), bool(white_bkgd), bool(calc_normal), bool(detailed_output))
            ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
            if calc_normal:
                ret_i['normals_volume'] = normals
            if detailed_output:
                ret_i['implicit_nablas'] = nablas
                ret_i['implicit_surface'] = sdf
                ret_i['radiance'] = blended if N_outside > 0 else radiances
                ret_i['alpha'] = alpha
                ret_i['cdf'] = cdf
                ret_i['visibility_weights'] = w
                ret_i['d_final'] = d_final
                if N_outside > 0:
                    ret_i['sigma_out'] = sigma_out
                    ret_i['radiance_out'] = radiance_out
            outs.append(ret_i)

    ret = OrderedDict()
    for k in outs[0].keys():
        v = outs[0][k] if len(outs) == 1 else torch.cat([o[k] for o in outs], 0)
        ret[k] = v.reshape(*prefix, *v.shape[1:]) if batched else v
    return ret['rgb'], ret['depth_volume'], ret


class SingleRenderer(nn.Module):
    """neus.py:399-405."""

    def __init__(self, model: NeuS):
        super().__init__()
        self.model = model

    def forward(self, rays_o, rays_d, **kwargs):
        return volume_render(rays_o, rays_d, self.model, **kwargs)


def __getattr__(name):
    """``Trainer`` and ``get_model`` (neus.py of the reference) live in frameworks/trainers.py; resolved lazily because
    that module imports this one."""
    if name in ("Trainer", "get_model"):
        from . import trainers
        return {"Trainer": trainers.NeusTrainer, "get_model": trainers.get_model_neus}[name]
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
