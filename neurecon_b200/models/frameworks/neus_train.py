"""NeuS ``volume_render`` under autograd (the training path, neus.py:118-397 called from
``Trainer.forward`` neus.py:440 with ``detailed_output=True``).

* up-sampling is ``no_grad`` in the reference (neus.py:214) -> the inference kernels (any tier);
* the two with-grad network queries (neus.py:294,298) go through the hand-written fp32 forward/backward
  of ``models/autograd.py`` (incl. the second-order path of the eikonal loss);
* alpha / transmittance / compositing (neus.py:296,346-352) are a dozen [R,128] tensor ops whose
  gradients reach ``ln_s``, the sdf values and the radiances through PyTorch autograd -- < 0.1 % of the
  step's work (the step is ~1 TFLOP of MLP), so they are not worth a custom backward.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from ... import _lib


def volume_render_train(rays_o, rays_d, model, obj_bounding_radius=1.0, batched=False, calc_normal=False,
                        rayschunk=65536, white_bkgd=False, near_bypass=None, far_bypass=None, detailed_output=True,
                        perturb=False, N_samples=64, N_importance=64, N_upsample_iters=4, N_outside=0):
    from . import neus
    B = rays_d.shape[0] if batched else 1
    prefix = [B, -1] if batched else [-1]
    o_flat = _lib.f32c(rays_o.reshape(-1, 3))
    d_flat = _lib.f32c(rays_d.reshape(-1, 3))
    n_total = o_flat.shape[0]
    M = N_samples + (N_importance // N_upsample_iters) * N_upsample_iters
    outs = []
    with torch.cuda.device(o_flat.device):
        step = int(rayschunk) * B
        for i0 in range(0, n_total, step):
            ro, rd = o_flat[i0:i0 + step], d_flat[i0:i0 + step]
            R = ro.shape[0]
            with torch.no_grad():
                dirs, d_all, pts, d_mid, pts_mid, far = neus._upsample(
                    model, ro, rd, obj_bounding_radius, near_bypass, far_bypass, N_samples, N_importance,
                    N_upsample_iters, perturb, return_far=True)
            sdf, nablas, _ = model.implicit_surface.forward_with_nablas(pts)          # neus.py:294
            views = dirs.unsqueeze(-2).expand(R, M - 1, 3)
            radiances = model.forward_radiance(pts_mid, views)                         # neus.py:298
            cdf, alpha = neus.sdf_to_alpha(sdf, model.forward_s())                     # neus.py:296
            d_final = d_mid
            if N_outside > 0:                                                          # neus.py:303-343
                dev = ro.device
                with torch.no_grad():
                    t = torch.linspace(0, 1, N_outside + 2)[..., 1:-1].float().to(dev)
                    d_out = far[..., None] / torch.flip(t, dims=[-1])
                    if perturb:
                        mids = .5 * (d_out[..., 1:] + d_out[..., :-1])
                        upper = torch.cat([mids, d_out[..., -1:]], -1)
                        lower = torch.cat([d_out[..., :1], mids], -1)
                        d_out = lower + (upper - lower) * torch.rand(upper.shape).float().to(dev)
                    d_final = torch.cat([d_mid, d_out], dim=-1)
                    pts_out = ro[..., None, :] + dirs[..., None, :] * d_final[..., :, None]
                    r = pts_out.norm(dim=-1, keepdim=True)
                    x_out = torch.cat([pts_out / r, 1. / r], dim=-1)
                    inside = (pts_mid.norm(dim=-1) <= obj_bounding_radius).float()
                sigma_out, radiance_out = model.nerf_outside.forward(x_out, dirs.unsqueeze(-2).expand(R, d_final.shape[-1], 3))
                dists = d_final[..., 1:] - d_final[..., :-1]
                dists = torch.cat([dists, 1e10 * torch.ones_like(dists[..., :1])], dim=-1)
                alpha_out = 1 - torch.exp(-F.softplus(sigma_out) * dists)
                n1 = d_mid.shape[-1]
                alpha = torch.cat([alpha * inside + alpha_out[..., :n1] * (1 - inside), alpha_out[..., n1:]], dim=-1)
                radiances = torch.cat([radiances * inside[..., None] + radiance_out[..., :n1, :] * (1 - inside)[..., None],
                                       radiance_out[..., n1:, :]], dim=-2)
            w = neus.alpha_to_w(alpha)                                                 # neus.py:346
            rgb = torch.sum(w[..., None] * radiances, -2)
            depth = torch.sum(w / (w.sum(-1, keepdim=True) + 1e-10) * d_final, -1)
            acc = torch.sum(w, -1)
            if white_bkgd:
                rgb = rgb + (1.0 - acc[..., None])
            ret_i = OrderedDict([('rgb', rgb), ('depth_volume', depth), ('mask_volume', acc)])
            if calc_normal:
                nm = F.normalize(nablas, dim=-1)
                n_pts = min(w.shape[-1], nm.shape[-2])                                    # neus.py:365-367
                ret_i['normals_volume'] = (nm[..., :n_pts, :] * w[..., :n_pts, None]).sum(dim=-2)
            if detailed_output:
                ret_i['implicit_nablas'] = nablas
                ret_i['implicit_surface'] = sdf
                ret_i['radiance'] = radiances
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
