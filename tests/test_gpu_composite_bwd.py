"""Hand-written adjoints of the compositing passes (csrc/composite_bwd.cu, models/composite.py) against torch autograd
through the fp64 oracle compositions (oracle/{neus,volsdf,unisurf}.py), and the NeRF++ sphere helpers (csrc/sphere.cu)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4      # north_star: fp32 path <= 1e-4 relative (max|a-b| / max|b| per tensor)


def _rays(R, M, seed, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    d = torch.sort(torch.rand(R, M, generator=g) * 2.0 + 0.5, dim=-1).values
    # an sdf that crosses zero along the ray, with noise
    sdf = (1.6 - d) * 0.4 + 0.03 * torch.randn(R, M, generator=g)
    nab = torch.randn(R, M, 3, generator=g)
    return d, sdf, nab, g


def _weights(R, n, g):
    return torch.randn(R, n, generator=g)


def _loss(ret, cw, keys):
    return sum((ret[k].double() * cw[k].double()).sum() for k in keys)


def test_neus_composite_adjoint():
    from oracle import neus as oneus
    from neurecon_b200.models.composite import NeusComposite
    R, M = 37, 128
    d, sdf, nab, g = _rays(R, M, 1)
    rad = torch.rand(R, M - 1, 3, generator=g)
    s = torch.tensor([23.0])
    cw = dict(rgb=_weights(R, 3, g), depth_volume=_weights(R, 1, g)[:, 0], mask_volume=_weights(R, 1, g)[:, 0],
              normals_volume=_weights(R, 3, g), visibility_weights=0.1 * _weights(R, M - 1, g))
    keys = list(cw)
    for white in (False, True):
        # oracle, fp64 autograd
        a = [t.double().requires_grad_() for t in (sdf, nab, rad, s)]
        ret = oneus.composite(a[0], a[1], a[2], d.double(), a[3], white_bkgd=white)
        want = torch.autograd.grad(_loss(ret, cw, keys), a)
        # kernels
        b = [t.cuda().requires_grad_() for t in (sdf, nab, rad, s)]
        d_mid = (0.5 * (d[:, 1:] + d[:, :-1])).cuda()
        rgb, depth, acc, normals, cdf, alpha, w, _ = NeusComposite.apply(b[0], b[1], b[2], d_mid, b[3], None, None, None, None,
                                                                         1.0, 0, white, True, True)
        got_ret = dict(rgb=rgb, depth_volume=depth, mask_volume=acc, normals_volume=normals, visibility_weights=w)
        for k in keys:
            assert rel_err(got_ret[k], ret[k]) < TOL, (k, rel_err(got_ret[k], ret[k]))
        got = torch.autograd.grad(_loss(got_ret, {k: v.cuda() for k, v in cw.items()}, keys), b)
        for name, x, y in zip(("sdf", "nablas", "radiance", "s"), got, want):
            assert rel_err(x, y) < TOL, (white, name, rel_err(x, y))


def test_neus_composite_bg_adjoint():
    """neus.py:303-352 with the NeRF++ blend: gradients reach sigma_out / radiance_out outside the sphere and behind it,
    the sdf / inside radiance only where the mid point is inside."""
    from oracle import neus as oneus
    from neurecon_b200.models.composite import NeusComposite
    R, M, n_out, radius = 29, 128, 32, 1.0
    d, sdf, nab, g = _rays(R, M, 2)
    T = M - 1 + n_out
    o = F.normalize(torch.randn(R, 3, generator=g), dim=-1) * 1.8
    dirs = F.normalize(-o + 0.2 * torch.randn(R, 3, generator=g), dim=-1)
    d_mid = 0.5 * (d[:, 1:] + d[:, :-1])
    d_out = d[:, -1:] + 0.5 + torch.cumsum(torch.rand(R, n_out, generator=g) + 0.05, dim=-1)
    d_vals = torch.cat([d_mid, d_out], dim=-1)
    rad = torch.rand(R, M - 1, 3, generator=g)
    sig_o = 2.0 * torch.randn(R, T, generator=g)
    sig_o[0, 3] = 25.0       # softplus' linear branch
    rad_o = torch.rand(R, T, 3, generator=g)
    s = torch.tensor([31.0])
    cw = dict(rgb=_weights(R, 3, g), depth_volume=_weights(R, 1, g)[:, 0], mask_volume=_weights(R, 1, g)[:, 0],
              normals_volume=_weights(R, 3, g), visibility_weights=0.1 * _weights(R, T, g))
    keys = list(cw)

    def ref(sdf_, nab_, rad_, sig_, rado_, s_):      # restates neus.py:320-352,365-367 in fp64 (test infrastructure)
        dd = d_vals.double()
        pts_mid = o.double()[:, None] + dirs.double()[:, None] * dd[:, :M - 1, None]
        inside = (pts_mid.float().norm(dim=-1) <= radius).double()
        cdf, a_in = oneus.sdf_to_alpha(sdf_, s_)
        dists = torch.cat([dd[:, 1:] - dd[:, :-1], 1e10 * torch.ones_like(dd[:, :1])], dim=-1)
        a_out = 1 - torch.exp(-F.softplus(sig_) * dists)
        alpha = torch.cat([a_in * inside + a_out[:, :M - 1] * (1 - inside), a_out[:, M - 1:]], dim=-1)
        rr = torch.cat([rad_ * inside[..., None] + rado_[:, :M - 1] * (1 - inside)[..., None], rado_[:, M - 1:]], dim=-2)
        w = oneus.alpha_to_w(alpha)
        acc = w.sum(-1)
        n = F.normalize(nab_, dim=-1)
        N = min(T, M)
        return dict(rgb=(w[..., None] * rr).sum(-2), depth_volume=(w / (acc[:, None] + 1e-10) * dd).sum(-1), mask_volume=acc,
                    normals_volume=(n[:, :N] * w[:, :N, None]).sum(-2), visibility_weights=w), inside

    a = [t.double().requires_grad_() for t in (sdf, nab, rad, sig_o, rad_o, s)]
    ret, inside = ref(*a)
    assert 0.05 < inside.mean().item() < 0.95, "the test rays should cross the bounding sphere"
    want = torch.autograd.grad(_loss(ret, cw, keys), a)
    b = [t.cuda().requires_grad_() for t in (sdf, nab, rad, sig_o, rad_o, s)]
    rgb, depth, acc, normals, cdf, alpha, w, blend = NeusComposite.apply(
        b[0], b[1], b[2], d_vals.cuda(), b[5], b[3], b[4], o.cuda(), dirs.cuda(), radius, n_out, False, True, True)
    got_ret = dict(rgb=rgb, depth_volume=depth, mask_volume=acc, normals_volume=normals, visibility_weights=w)
    for k in keys:
        assert rel_err(got_ret[k], ret[k]) < TOL, (k, rel_err(got_ret[k], ret[k]))
    got = torch.autograd.grad(_loss(got_ret, {k: v.cuda() for k, v in cw.items()}, keys), b)
    for name, x, y in zip(("sdf", "nablas", "radiance", "sigma_out", "radiance_out", "s"), got, want):
        assert rel_err(x, y) < TOL, (name, rel_err(x, y))


@pytest.mark.parametrize("m_out", [0, 32])
def test_volsdf_composite_adjoint(m_out):
    """volsdf.py:452-503 incl. ln_beta's two paths (alpha = 1 / beta and beta) and rays whose product of p hits zero."""
    from oracle import volsdf as ovol
    from neurecon_b200.models.composite import VolsdfComposite
    R, M_in = 33, 192
    d, sdf, nab, g = _rays(R, M_in, 3 + m_out)
    rad = torch.rand(R, M_in, 3, generator=g)
    ln_beta = torch.tensor([math.log(0.02) / 10.0])
    sig_o = (3.0 * torch.rand(R, m_out, generator=g)) if m_out else None
    rad_o = torch.rand(R, m_out, 3, generator=g) if m_out else None
    # rays 0 and 1: one interval, resp. two intervals, opaque enough for exp(-x) to underflow to exactly 0
    d = d.clone()
    d[0, 100:] += 40.0
    sdf[0, 99] = -1.0      # sigma_99 * (d_100 - d_99) = 50 * 40
    d[1, 60:] += 40.0
    d[1, 90:] += 40.0
    sdf[1, 59] = -1.0
    sdf[1, 89] = -1.0
    d_o = d[:, -1:] + 0.3 + torch.cumsum(torch.rand(R, max(m_out, 1), generator=g) + 0.05, dim=-1) if m_out else None
    M = M_in + m_out
    cw = dict(rgb=_weights(R, 3, g), depth_volume=_weights(R, 1, g)[:, 0], mask_volume=_weights(R, 1, g)[:, 0],
              normals_volume=_weights(R, 3, g), visibility_weights=0.1 * _weights(R, M - 1, g))
    keys = list(cw)

    def ref(sdf_, nab_, rad_, lnb, sig_, rado_):
        beta = torch.exp(lnb * 10.0)
        alpha = 1.0 / beta
        sigma = ovol.sdf_to_sigma(sdf_, alpha, beta)
        dd = d.double()
        if m_out:
            sigma = torch.cat([sigma, sig_], dim=-1)
            rad_ = torch.cat([rad_, rado_], dim=-2)
            dd = torch.cat([dd, d_o.double()], dim=-1)
        return ovol.composite(sigma, rad_, nab_, dd)

    ins = [sdf, nab, rad, ln_beta] + ([sig_o, rad_o] if m_out else [])
    a = [t.double().requires_grad_() for t in ins]
    ret = ref(*(a + ([None, None] if not m_out else [])))
    p_ref = ret["p_i"]
    assert (p_ref[0] == 0).sum() == 1 and (p_ref[1] == 0).sum() == 2, "the zero-product rays must underflow in fp64 too"
    want = torch.autograd.grad(_loss(ret, cw, keys), a)
    b = [t.cuda().requires_grad_() for t in ins]
    beta_c = torch.exp(b[3] * 10.0)
    rgb, depth, acc, normals, sigma_all, p_i, tau = VolsdfComposite.apply(
        b[0], b[1], b[2], d.cuda(), 1.0 / beta_c, beta_c, b[4] if m_out else None, b[5] if m_out else None,
        d_o.cuda() if m_out else None, False, True, True)
    got_ret = dict(rgb=rgb, depth_volume=depth, mask_volume=acc, normals_volume=normals, visibility_weights=tau)
    for k in keys:
        assert rel_err(got_ret[k], ret[k]) < TOL, (k, rel_err(got_ret[k], ret[k]))
    got = torch.autograd.grad(_loss(got_ret, {k: v.cuda() for k, v in cw.items()}, keys), b)
    names = ("sdf", "nablas", "radiance", "ln_beta") + (("sigma_out", "radiance_out") if m_out else ())
    for name, x, y in zip(names, got, want):
        assert rel_err(x, y) < TOL, (name, rel_err(x, y))
    # the zero-product rays individually (their gradients are tiny next to the others' maximum)
    for r in (0, 1):
        assert rel_err(got[0][r], want[0][r]) < TOL, ("sdf of zero-product ray", r, rel_err(got[0][r], want[0][r]))


def test_unisurf_composite_adjoint():
    from oracle import unisurf as ouni
    from neurecon_b200.models.composite import UnisurfComposite
    R, M = 41, 96
    d, sdf, nab, g = _rays(R, M, 7)
    logits = sdf * 40.0
    rad = torch.rand(R, M, 3, generator=g)
    cw = dict(rgb=_weights(R, 3, g), depth_volume=_weights(R, 1, g)[:, 0], mask_volume=_weights(R, 1, g)[:, 0],
              normals_volume=_weights(R, 3, g), visibility_weights=0.1 * _weights(R, M, g))
    keys = list(cw)
    for white in (False, True):
        a = [t.double().requires_grad_() for t in (logits, rad, nab)]
        ret = ouni.composite(a[0], a[1], a[2], d.double(), white_bkgd=white)
        want = torch.autograd.grad(_loss(ret, cw, keys), a)
        b = [t.cuda().requires_grad_() for t in (logits, rad, nab)]
        rgb, depth, acc, normals, alpha, w = UnisurfComposite.apply(b[0], b[2], b[1], d.cuda(), white, True, True)
        got_ret = dict(rgb=rgb, depth_volume=depth, mask_volume=acc, normals_volume=normals, visibility_weights=w)
        for k in keys:
            assert rel_err(got_ret[k], ret[k]) < TOL, (k, rel_err(got_ret[k], ret[k]))
        got = torch.autograd.grad(_loss(got_ret, {k: v.cuda() for k, v in cw.items()}, keys), b)
        for name, x, y in zip(("logits", "radiance", "nablas"), got, want):
            assert rel_err(x, y) < TOL, (white, name, rel_err(x, y))


def test_composite_without_upstream_gradients():
    """Only rgb has a consumer (the usual case): the other outputs' gradients arrive as None."""
    from neurecon_b200.models.composite import NeusComposite
    R, M = 8, 128
    d, sdf, nab, g = _rays(R, M, 9)
    sdf_c = sdf.cuda().requires_grad_()
    rad = torch.rand(R, M - 1, 3, generator=g).cuda().requires_grad_()
    s = torch.tensor([20.0], device="cuda", requires_grad=True)
    out = NeusComposite.apply(sdf_c, None, rad, (0.5 * (d[:, 1:] + d[:, :-1])).cuda(), s, None, None, None, None, 1.0, 0,
                              False, False, False)
    out[0].sum().backward()
    assert sdf_c.grad is not None and torch.isfinite(sdf_c.grad).all() and sdf_c.grad.abs().max() > 0
    assert torch.isfinite(rad.grad).all() and torch.isfinite(s.grad).all()


# ---- sphere helpers (rend_util.py:188-234, volsdf.py:456-467) ---------------------------------------------------------
def test_sphere_intersection_and_dvals():
    """rend_util.py:188-234 at the SURVEY's bar for the sphere helpers (<= 1e-6 relative, identical hit mask)."""
    from oracle import sampling as osamp
    from neurecon_b200.utils import rend_util
    g = torch.Generator().manual_seed(11)
    R = 4099
    o = F.normalize(torch.randn(R, 3, generator=g), dim=-1) * (1.0 + 2.0 * torch.rand(R, 1, generator=g))
    dirs = F.normalize(-o + 0.8 * torch.randn(R, 3, generator=g), dim=-1)
    near_w, far_w, mask_w = osamp.get_sphere_intersection(o, dirs, r=1.3)
    near, far, mask = rend_util.get_sphere_intersection(o.cuda(), dirs.cuda(), r=1.3)
    assert 0.2 < mask_w.float().mean() < 0.98
    assert torch.equal(mask.cpu(), mask_w)
    assert rel_err(near, near_w) < 1e-6 and rel_err(far, far_w) < 1e-6, (rel_err(near, near_w), rel_err(far, far_w))
    # radii beyond every origin: the assertion of rend_util.py:225 holds
    rs = (o.norm(dim=-1, keepdim=True) + 0.1) * (1.0 + torch.rand(R, 32, generator=g))
    for far_end in (True, False):
        want = osamp.get_dvals_from_radius(o, dirs, rs, far_end=far_end)
        got = rend_util.get_dvals_from_radius(o.cuda(), dirs.cuda(), rs.cuda(), far_end=far_end, strict=True)
        assert rel_err(got, want) < 1e-6, (far_end, rel_err(got, want))
    # a radius inside the closest approach: NaN there, and strict=True raises like the reference's assert
    rs_bad = rs.clone()
    rs_bad[5, 7] = 1e-3
    got = rend_util.get_dvals_from_radius(o.cuda(), dirs.cuda(), rs_bad.cuda())
    assert torch.isnan(got[5, 7]) and torch.isfinite(got[6]).all()
    with pytest.raises(AssertionError):
        rend_util.get_dvals_from_radius(o.cuda(), dirs.cuda(), rs_bad.cuda(), strict=True)


def test_volsdf_outside_points_matches_reference_ops():
    """nr_volsdf_outside_points against the tensor ops of volsdf.py:456-467 (torch fp32 on the CPU)."""
    from neurecon_b200 import _lib
    from oracle import sampling as osamp
    lib = _lib.get_lib()
    g = torch.Generator().manual_seed(12)
    R, n_out, radius = 515, 32, 3.0
    o = F.normalize(torch.randn(R, 3, generator=g), dim=-1) * 2.7
    dirs = F.normalize(-o + 0.2 * torch.randn(R, 3, generator=g), dim=-1)
    for perturb in (False, True):
        t = torch.linspace(0, 1, n_out + 2)[..., 1:-1].float()
        rs = (radius / torch.flip(t, dims=[-1])).expand(R, n_out)
        u = torch.rand(R, n_out, generator=g) if perturb else None
        if perturb:
            mids = .5 * (rs[..., 1:] + rs[..., :-1])
            upper = torch.cat([mids, rs[..., -1:]], -1)
            lower = torch.cat([rs[..., :1], mids], -1)
            rs = lower + (upper - lower) * u
        d_w = osamp.get_dvals_from_radius(o, dirs, rs)
        pts = o[:, None] + dirs[:, None] * d_w[..., None]
        x_w = torch.cat([pts / rs[..., None], 1. / rs[..., None]], dim=-1)
        oc, dc = o.cuda(), dirs.cuda()
        d_out = torch.empty(R, n_out, device="cuda")
        x_out = torch.empty(R, n_out, 4, device="cuda")
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.nr_volsdf_outside_points(_lib.ptr(oc), _lib.ptr(dc), R, radius, n_out, _lib.ptr(u.cuda() if perturb else None),
                                                _lib.ptr(d_out), _lib.ptr(x_out), _lib.ptr(bad), _lib.stream_ptr(oc.device)), "outside")
        assert int(bad.item()) == 0
        assert rel_err(d_out, d_w) < 1e-6 and rel_err(x_out, x_w) < 1e-6, (perturb, rel_err(d_out, d_w), rel_err(x_out, x_w))
