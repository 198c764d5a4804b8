"""Differentiable alpha / transmittance / compositing of the three frameworks.

Forward = the inference kernels (``nr_neus_composite[_bg]``, ``nr_volsdf_composite``, ``nr_unisurf_composite``) with
their per-sample outputs; backward = their hand-written adjoints in ``csrc/composite_bwd.cu`` (one launch per ray
batch).  These Functions are what ``volume_render`` calls in inference AND in training, so the training render has no
tensor-op restatement of neus.py:296-352 / volsdf.py:452-503 / unisurf.py:216-240 any more.

Gradient paths: rgb, depth_volume, mask_volume, normals_volume and visibility_weights back to the sdf / logits, the
radiances, the nablas (through the normals), ``s`` (NeuS), ``alpha, beta`` (VolSDF) and the NeRF++ ``sigma_out /
radiance_out``.  alpha, cdf, p_i, sigma and the blended radiance are returned without a gradient path (no loss of the
reference consumes them).
"""
import torch

from .. import _lib


def _f(t):
    return None if t is None else _lib.f32c(t)


def _g(t, like=None):
    """contiguous fp32 upstream gradient or None"""
    return None if t is None else _lib.f32c(t)


class NeusComposite(torch.autograd.Function):
    """(sdf [R,M], nablas [R,M,3] | None, radiances [R,M-1,3], d_vals [R,K], s [1],
        sigma_out [R,K] | None, radiance_out [R,K,3] | None, rays_o, dirs, radius, n_out, white_bkgd)
       -> rgb, depth, acc, normals | None, cdf, alpha, weights, blended radiance [R,K,3] (None without background)"""

    @staticmethod
    def forward(ctx, sdf, nablas, radiances, d_vals, s, sigma_out, radiance_out, rays_o, dirs, radius, n_out,
                white_bkgd, calc_normal, detailed=True):
        lib = _lib.get_lib()
        detailed = bool(detailed) or any(ctx.needs_input_grad)   # the backward reads the per-sample outputs
        sdf, radiances, d_vals, s = _f(sdf), _f(radiances), _f(d_vals), _f(s)
        nablas = _f(nablas)
        R, M = sdf.shape
        K = M - 1 + n_out
        dev = sdf.device
        f = dict(dtype=torch.float32, device=dev)
        rgb, depth, acc = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
        normals = torch.empty(R, 3, **f) if calc_normal else None
        cdf = torch.empty(R, M, **f) if detailed else None
        alpha = torch.empty(R, K, **f) if detailed else None
        w = torch.empty(R, K, **f) if detailed else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            if n_out > 0:
                sigma_out, radiance_out, rays_o, dirs = _f(sigma_out), _f(radiance_out), _f(rays_o), _f(dirs)
                used = torch.empty(R, K, 3, **f) if detailed else None
                _lib.check(lib.nr_neus_composite_bg(
                    _lib.ptr(sdf), _lib.ptr(nablas) if calc_normal else None, _lib.ptr(radiances), _lib.ptr(rays_o),
                    _lib.ptr(dirs), _lib.ptr(d_vals), _lib.ptr(sigma_out), _lib.ptr(radiance_out), _lib.ptr(s),
                    float(radius), R, M, n_out, int(bool(white_bkgd)), _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc),
                    _lib.ptr(normals), _lib.ptr(cdf), _lib.ptr(alpha), _lib.ptr(w), _lib.ptr(used), st), "neus_composite_bg")
            else:
                used = radiances
                _lib.check(lib.nr_neus_composite(
                    _lib.ptr(sdf), _lib.ptr(nablas) if calc_normal else None, _lib.ptr(radiances), _lib.ptr(d_vals),
                    _lib.ptr(s), R, M, int(bool(white_bkgd)), _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc),
                    _lib.ptr(normals), _lib.ptr(cdf), _lib.ptr(alpha), _lib.ptr(w), st), "neus_composite")
        ctx.save_for_backward(sdf, nablas if calc_normal else None, d_vals, s, sigma_out if n_out > 0 else None,
                              rays_o if n_out > 0 else None, dirs if n_out > 0 else None, cdf, alpha, w, used, acc, depth)
        ctx.cfg = (R, M, n_out, float(radius), int(bool(white_bkgd)))
        if detailed:
            ctx.mark_non_differentiable(cdf, alpha)
            if n_out > 0:
                ctx.mark_non_differentiable(used)
        return rgb, depth, acc, normals, cdf, alpha, w, (used if n_out > 0 else None)

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, g_normals, _g_cdf, _g_alpha, g_w, _g_used):
        lib = _lib.get_lib()
        sdf, nablas, d_vals, s, sigma_out, rays_o, dirs, cdf, alpha, w, used, acc, depth = ctx.saved_tensors
        R, M, n_out, radius, white = ctx.cfg
        K = M - 1 + n_out
        dev = sdf.device
        f = dict(dtype=torch.float32, device=dev)
        g_rgb, g_depth, g_acc, g_normals, g_w = _g(g_rgb), _g(g_depth), _g(g_acc), _g(g_normals), _g(g_w)
        if nablas is None:
            g_normals = None
        g_sdf, g_s_part = torch.empty(R, M, **f), torch.empty(R, **f)
        g_rad = torch.empty(R, M - 1, 3, **f)
        g_nab = torch.empty(R, M, 3, **f) if g_normals is not None else None
        g_sig = torch.empty(R, K, **f) if n_out > 0 else None
        g_rad_out = torch.empty(R, K, 3, **f) if n_out > 0 else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_neus_composite_bwd(
                _lib.ptr(sdf), _lib.ptr(cdf), _lib.ptr(alpha), _lib.ptr(w), _lib.ptr(used), _lib.ptr(d_vals),
                _lib.ptr(nablas), _lib.ptr(s), _lib.ptr(acc), _lib.ptr(depth), R, M, n_out, _lib.ptr(rays_o),
                _lib.ptr(dirs), _lib.ptr(sigma_out), radius, white, _lib.ptr(g_rgb), _lib.ptr(g_depth), _lib.ptr(g_acc),
                _lib.ptr(g_normals), _lib.ptr(g_w), _lib.ptr(g_sdf), _lib.ptr(g_s_part), _lib.ptr(g_rad), _lib.ptr(g_nab),
                _lib.ptr(g_sig), _lib.ptr(g_rad_out), _lib.stream_ptr(dev)), "neus_composite_bwd")
        g_s = g_s_part.sum().reshape(s.shape) if ctx.needs_input_grad[4] else None
        return (g_sdf, g_nab, g_rad, None, g_s, g_sig, g_rad_out, None, None, None, None, None, None, None)


class VolsdfComposite(torch.autograd.Function):
    """(sdf [R,M_in], nablas | None, radiances [R,M_in,3], d_in [R,M_in], alpha [1], beta [1],
        sigma_out [R,M_out] | None, radiance_out | None, d_out | None, white_bkgd, calc_normal)
       -> rgb, depth, acc, normals | None, sigma_all [R,M], p_i [R,M-1], tau [R,M-1]"""

    @staticmethod
    def forward(ctx, sdf, nablas, radiances, d_in, alpha, beta, sigma_out, radiance_out, d_out, white_bkgd, calc_normal,
                detailed=True):
        lib = _lib.get_lib()
        detailed = bool(detailed) or any(ctx.needs_input_grad)
        sdf, radiances, d_in, alpha, beta = _f(sdf), _f(radiances), _f(d_in), _f(alpha), _f(beta)
        nablas, sigma_out, radiance_out, d_out = _f(nablas), _f(sigma_out), _f(radiance_out), _f(d_out)
        R, M_in = sdf.shape
        M_out = 0 if sigma_out is None else sigma_out.shape[-1]
        M = M_in + M_out
        dev = sdf.device
        f = dict(dtype=torch.float32, device=dev)
        rgb, depth, acc = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
        normals = torch.empty(R, 3, **f) if calc_normal else None
        sigma_all = torch.empty(R, M, **f) if detailed else None
        p_i = torch.empty(R, M - 1, **f) if detailed else None
        tau = torch.empty(R, M - 1, **f) if detailed else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_volsdf_composite(
                _lib.ptr(sdf), _lib.ptr(nablas) if calc_normal else None, _lib.ptr(radiances), _lib.ptr(d_in),
                _lib.ptr(alpha), _lib.ptr(beta), R, M_in, _lib.ptr(sigma_out), _lib.ptr(radiance_out), _lib.ptr(d_out),
                M_out, int(bool(white_bkgd)), _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(normals),
                _lib.ptr(sigma_all), _lib.ptr(p_i), _lib.ptr(tau), _lib.stream_ptr(dev)), "volsdf_composite")
        ctx.save_for_backward(sdf, nablas if calc_normal else None, radiances, d_in, alpha, beta, radiance_out, d_out,
                              sigma_all, p_i, tau, acc, depth)
        ctx.cfg = (R, M_in, M_out, int(bool(white_bkgd)))
        if detailed:
            ctx.mark_non_differentiable(sigma_all, p_i)
        return rgb, depth, acc, normals, sigma_all, p_i, tau

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, g_normals, _g_sigma, _g_p, g_w):
        lib = _lib.get_lib()
        sdf, nablas, radiances, d_in, alpha, beta, radiance_out, d_out, sigma_all, p_i, tau, acc, depth = ctx.saved_tensors
        R, M_in, M_out, white = ctx.cfg
        dev = sdf.device
        f = dict(dtype=torch.float32, device=dev)
        g_rgb, g_depth, g_acc, g_normals, g_w = _g(g_rgb), _g(g_depth), _g(g_acc), _g(g_normals), _g(g_w)
        if nablas is None:
            g_normals = None
        g_sdf = torch.empty(R, M_in, **f)
        g_a_part, g_b_part = torch.empty(R, **f), torch.empty(R, **f)
        g_rad = torch.empty(R, M_in, 3, **f)
        g_nab = torch.empty(R, M_in, 3, **f) if g_normals is not None else None
        g_sig = torch.empty(R, M_out, **f) if M_out > 0 else None
        g_rad_out = torch.empty(R, M_out, 3, **f) if M_out > 0 else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_volsdf_composite_bwd(
                _lib.ptr(sdf), _lib.ptr(sigma_all), _lib.ptr(p_i), _lib.ptr(tau), _lib.ptr(radiances), _lib.ptr(d_in),
                _lib.ptr(nablas), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(acc), _lib.ptr(depth), R, M_in,
                _lib.ptr(radiance_out), _lib.ptr(d_out), M_out, white, _lib.ptr(g_rgb), _lib.ptr(g_depth), _lib.ptr(g_acc),
                _lib.ptr(g_normals), _lib.ptr(g_w), _lib.ptr(g_sdf), _lib.ptr(g_a_part), _lib.ptr(g_b_part),
                _lib.ptr(g_rad), _lib.ptr(g_nab), _lib.ptr(g_sig), _lib.ptr(g_rad_out), _lib.stream_ptr(dev)),
                "volsdf_composite_bwd")
        g_alpha = g_a_part.sum().reshape(alpha.shape) if ctx.needs_input_grad[4] else None
        g_beta = g_b_part.sum().reshape(beta.shape) if ctx.needs_input_grad[5] else None
        return (g_sdf, g_nab, g_rad, None, g_alpha, g_beta, g_sig, g_rad_out, None, None, None, None)


class UnisurfComposite(torch.autograd.Function):
    """(logits [R,M], nablas | None, radiances [R,M,3], d_all [R,M], white_bkgd, calc_normal)
       -> rgb, depth, acc, normals | None, alpha, weights"""

    @staticmethod
    def forward(ctx, logits, nablas, radiances, d_all, white_bkgd, calc_normal, detailed=True):
        lib = _lib.get_lib()
        detailed = bool(detailed) or any(ctx.needs_input_grad)
        logits, radiances, d_all, nablas = _f(logits), _f(radiances), _f(d_all), _f(nablas)
        R, M = logits.shape
        dev = logits.device
        f = dict(dtype=torch.float32, device=dev)
        rgb, depth, acc = torch.empty(R, 3, **f), torch.empty(R, **f), torch.empty(R, **f)
        normals = torch.empty(R, 3, **f) if calc_normal else None
        alpha = torch.empty(R, M, **f) if detailed else None
        w = torch.empty(R, M, **f) if detailed else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_unisurf_composite(
                _lib.ptr(logits), _lib.ptr(nablas) if calc_normal else None, _lib.ptr(radiances), _lib.ptr(d_all), R, M,
                int(bool(white_bkgd)), _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(normals), _lib.ptr(alpha),
                _lib.ptr(w), _lib.stream_ptr(dev)), "unisurf_composite")
        ctx.save_for_backward(logits, nablas if calc_normal else None, radiances, d_all, alpha, w, acc, depth)
        ctx.cfg = (R, M, int(bool(white_bkgd)))
        if detailed:
            ctx.mark_non_differentiable(alpha)
        return rgb, depth, acc, normals, alpha, w

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, g_normals, _g_alpha, g_w):
        lib = _lib.get_lib()
        logits, nablas, radiances, d_all, alpha, w, acc, depth = ctx.saved_tensors
        R, M, white = ctx.cfg
        dev = logits.device
        f = dict(dtype=torch.float32, device=dev)
        g_rgb, g_depth, g_acc, g_normals, g_w = _g(g_rgb), _g(g_depth), _g(g_acc), _g(g_normals), _g(g_w)
        if nablas is None:
            g_normals = None
        g_logits, g_rad = torch.empty(R, M, **f), torch.empty(R, M, 3, **f)
        g_nab = torch.empty(R, M, 3, **f) if g_normals is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_unisurf_composite_bwd(
                _lib.ptr(logits), _lib.ptr(alpha), _lib.ptr(w), _lib.ptr(radiances), _lib.ptr(d_all), _lib.ptr(nablas),
                _lib.ptr(acc), _lib.ptr(depth), R, M, white, _lib.ptr(g_rgb), _lib.ptr(g_depth), _lib.ptr(g_acc),
                _lib.ptr(g_normals), _lib.ptr(g_w), _lib.ptr(g_logits), _lib.ptr(g_rad), _lib.ptr(g_nab),
                _lib.stream_ptr(dev)), "unisurf_composite_bwd")
        return g_logits, g_nab, g_rad, None, None, None, None
