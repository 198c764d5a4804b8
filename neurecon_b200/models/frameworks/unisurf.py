"""Drop-in for the reference's ``models/frameworks/unisurf.py`` hot path: the ``UNISURF`` module
(:16-62), ``volume_render`` (:64-283) and ``SingleRenderer`` (:285-291)."""
from collections import OrderedDict
from typing import Optional, Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib
from ..base import ImplicitSurface, RadianceNet, query_radiance


class UNISURF(nn.Module):
    """unisurf.py:16-62 (same constructor, parameters and methods)."""

    def __init__(self, input_ch=3, W_geo_feat=-1, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        self.implicit_surface = ImplicitSurface(input_ch=input_ch, W_geo_feat=W_geo_feat, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)

    def forward(self, x, view_dirs):
        occ, nablas, geometry_feature = self.implicit_surface.forward_with_nablas(x)
        # NOTE: like the reference (unisurf.py:36) F.normalize has no dim => dim=1, the point axis of a
        # batched [B, points, 3] chunk.
        normals = F.normalize(nablas)
        radiances = self.radiance_net.forward(x, view_dirs, normals, geometry_feature)
        return radiances, occ, nablas

    @staticmethod
    def get_surface_from_opacity(opacity: Union[torch.Tensor, np.ndarray], eps=1e-4):
        if isinstance(opacity, torch.Tensor):
            opacity = torch.clamp(opacity, min=eps, max=1 - eps)
            return -1. * torch.log(opacity / (1 - opacity))
        opacity = np.clip(opacity, a_min=eps, a_max=1 - eps)
        return -1. * np.log(opacity / (1 - opacity))

    @staticmethod
    def get_opacity_from_surface(imp_surface: Union[torch.Tensor, np.ndarray]):
        if isinstance(imp_surface, torch.Tensor):
            odds = torch.exp(-1. * imp_surface)
        else:
            odds = np.exp(-1. * imp_surface)
        return odds / (1 + odds)


def volume_render(
        rays_o,
        rays_d,
        model: UNISURF,

        batched=False,
        batched_info={},

        # render algorithm config
        calc_normal=False,
        logit_tau=0.0,
        use_view_dirs=True,
        method='secant',
        rayschunk=65536,
        netchunk=1048576,
        white_bkgd=False,
        near_bypass: Optional[float] = None,
        far_bypass: Optional[float] = None,

        # render function config
        detailed_output=True,
        show_progress=False,

        # sampling related
        radius_of_interest=4.0,
        perturb=False,
        interval=1.0,
        too_close_threshold=0.1,
        N_query=64,
        N_freespace=32,

        # determinism hook for parity tests: {"d_all": [B, R, N_query + N_freespace] sorted depths} replaces the sampler's
        # depths (root finding still runs: its outputs are part of ``ret``)
        samples_bypass=None,
        **dummy_kwargs):
    """unisurf.py:64-283.  rays_o / rays_d: [(B,) N_rays, 3].  Returns (rgb, depth_volume, ret).

    Like the reference, the normals fed to the radiance net are normalised over the point axis of each
    net-chunk (``F.normalize`` without ``dim``, unisurf.py:36), so results depend on ``rayschunk`` /
    ``netchunk`` exactly as they do there; ``batched=False`` (which raises IndexError in the reference)
    is treated as one batch."""
    if method != 'secant':
        raise NotImplementedError("only method='secant' (every shipped config) is built")
    if bool(use_view_dirs) != bool(model.radiance_net.use_view_dirs):
        # the reference passes view_dirs=None for use_view_dirs=False, which only a RadianceNet built with
        # use_view_dirs=False accepts (base.py:379-384); here such a net ignores whatever views it is handed
        raise ValueError("use_view_dirs=%r needs a radiance net built with use_view_dirs=%r" % (use_view_dirs, use_view_dirs))
    _lib.require_cuda(rays_o, rays_d)
    # training (unisurf.py:307: Trainer.forward renders under autograd): root finding and sampling stay no_grad as in the
    # reference (ray_casting.py:35 is @torch.no_grad, the samples are detached), the network query goes through
    # UNISURF.forward (models/autograd.py), the compositing through UnisurfComposite
    train = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    from ..composite import UnisurfComposite
    from ..ray_casting import _root_find
    lib = _lib.get_lib()
    B = rays_d.shape[0] if batched else 1
    dev = rays_o.device
    o_b = _lib.f32c(rays_o.reshape(B, -1, 3))
    d_b = _lib.f32c(rays_d.reshape(B, -1, 3))
    n_rays = o_b.shape[1]
    f = dict(dtype=torch.float32, device=dev)
    M = N_query + N_freespace
    N_steps, N_secant = 256, 8  # root_finding_surface_points defaults (ray_casting.py:43-46)
    nan = float("nan")
    surface_fn = model.implicit_surface.forward

    def render_chunk(ro, rd, st, d_forced=None):
        R = ro.shape[0]
        with torch.no_grad():
            dirs, near, far = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
            pts_prop = torch.empty(R, N_steps, 3, **f)
            _lib.check(lib.nr_unisurf_ray_setup(
                _lib.ptr(ro), _lib.ptr(rd), R, float(radius_of_interest), nan if near_bypass is None else float(near_bypass),
                nan if far_bypass is None else float(far_bypass), N_steps, _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far),
                _lib.ptr(pts_prop), st), "unisurf_ray_setup")
            state, mask, msc, m0 = _root_find(surface_fn, ro, dirs, near, far, pts_prop, N_steps, logit_tau, N_secant)
            u_int = torch.rand([R, N_query], device=dev) if perturb else None     # unisurf.py:164
            u_free = torch.rand([R, N_freespace], device=dev) if perturb else None  # unisurf.py:193
            depth_s, surf_pts = torch.empty(R, **f), torch.empty(R, 3, **f)
            d_all, pts = torch.empty(R, M, **f), torch.empty(R, M, 3, **f)
            _lib.check(lib.nr_unisurf_sample(
                _lib.ptr(ro), _lib.ptr(dirs), _lib.ptr(near), _lib.ptr(far), _lib.ptr(state), _lib.ptr(mask),
                _lib.ptr(msc), _lib.ptr(m0), R, float(interval), float(too_close_threshold), N_query, N_freespace,
                _lib.ptr(u_int), _lib.ptr(u_free), _lib.ptr(depth_s), _lib.ptr(surf_pts), _lib.ptr(d_all),
                _lib.ptr(pts), st), "unisurf_sample")
            if d_forced is not None:
                d_all = _lib.f32c(d_forced.to(dev))
                pts = (ro[:, None, :] + dirs[:, None, :] * d_all[..., None]).contiguous()
                if "surface_points" in samples_bypass:
                    surf_pts = _lib.f32c(samples_bypass["surface_points"].reshape(-1, 3)[:R].to(dev))
        # network query, one net-chunk of the flattened points at a time (batchify_query semantics, incl. the chunk-wide
        # F.normalize of unisurf.py:36): the fused inference query, or UNISURF.forward under autograd on [1, chunk, 3]
        # slices -- the layout batchify_query(dim_batchify=1) feeds it (unisurf.py:214)
        flat_pts = pts.reshape(-1, 3)
        flat_views = dirs.unsqueeze(-2).expand(R, M, 3).reshape(-1, 3)
        rad_l, sdf_l, nab_l = [], [], []
        for j0 in range(0, R * M, int(netchunk)):
            if train:
                r_, s_, n_ = model.forward(flat_pts[None, j0:j0 + netchunk], flat_views[None, j0:j0 + netchunk])
                r_, s_, n_ = r_[0], s_[0], n_[0]
            else:
                with torch.no_grad():
                    r_, s_, n_ = query_radiance(model.implicit_surface, model.radiance_net, flat_pts[j0:j0 + netchunk],
                                                flat_views[j0:j0 + netchunk], chunk_normalize=True)
            rad_l.append(r_); sdf_l.append(s_); nab_l.append(n_)
        radiances = (rad_l[0] if len(rad_l) == 1 else torch.cat(rad_l)).reshape(R, M, 3)
        logits = (sdf_l[0] if len(sdf_l) == 1 else torch.cat(sdf_l)).reshape(R, M)
        nablas = (nab_l[0] if len(nab_l) == 1 else torch.cat(nab_l)).reshape(R, M, 3)
        rgb, depth, acc, normals, alpha, w = UnisurfComposite.apply(
            logits, nablas if calc_normal else None, radiances, d_all, bool(white_bkgd), bool(calc_normal),
            bool(detailed_output))
        ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
        if calc_normal:
            ret_i['normals_volume'] = normals
        if detailed_output:
            ret_i['surface_points'] = surf_pts
            ret_i['mask_surface'] = mask.bool()
            ret_i['depth_surface'] = depth_s
            ret_i['radiance'] = radiances
            ret_i['implicit_surface'] = logits
            ret_i['implicit_nablas'] = nablas
            ret_i['alpha'] = alpha
            ret_i['visibility_weights'] = w
        return ret_i

    per_batch = []
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for b in range(B):
            outs = [render_chunk(o_b[b, i0:i0 + rayschunk].contiguous(), d_b[b, i0:i0 + rayschunk].contiguous(), st,
                                 None if samples_bypass is None
                                 else samples_bypass["d_all"].reshape(B, n_rays, -1)[b, i0:i0 + rayschunk])
                    for i0 in range(0, n_rays, int(rayschunk))]
            per_batch.append(OrderedDict((k, outs[0][k] if len(outs) == 1 else torch.cat([o_[k] for o_ in outs], 0))
                                         for k in outs[0].keys()))
    ret = OrderedDict()
    for k in per_batch[0].keys():
        ret[k] = torch.stack([pb[k] for pb in per_batch], 0) if batched else per_batch[0][k]
    return ret['rgb'], ret['depth_volume'], ret


class SingleRenderer(nn.Module):
    """unisurf.py:285-291."""

    def __init__(self, model: UNISURF):
        super().__init__()
        self.model = model

    def forward(self, rays_o, rays_d, **kwargs):
        return volume_render(rays_o, rays_d, self.model, **kwargs)


def __getattr__(name):
    """``Trainer`` and ``get_model`` (unisurf.py of the reference) live in frameworks/trainers.py; resolved lazily because
    that module imports this one."""
    if name in ("Trainer", "get_model"):
        from . import trainers
        return {"Trainer": trainers.UnisurfTrainer, "get_model": trainers.get_model_unisurf}[name]
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
