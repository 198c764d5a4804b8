"""Reference behaviours outside the shipped configs (missing item 8): the NeRF-like up-samplers
``upsample_algo in {'direct_use', 'direct_more'}`` (neus.py:216-243) and radiance nets built with ``use_view_dirs=False``
(base.py:335-336,383-384; ray_casting.py:217-230), against tests/golden/neus_variants_r32.npz written by the UNMODIFIED
reference (tests/golden/make_golden.py variants)."""
import numpy as np
import pytest
import torch

import neurecon_b200
from neurecon_b200 import _lib
from neurecon_b200.models import ray_casting
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import synthetic
from oracle import neus as oneus
from conftest import build_neus, frac_close, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _restore_tier():
    yield
    neurecon_b200.set_precision("fp16")


def _noview_model(device=DEV):
    kw = dict(synthetic.NEUS_MODEL_KWARGS, radiance_cfg=dict(synthetic.NEUS_MODEL_KWARGS["radiance_cfg"], use_view_dirs=False))
    torch.manual_seed(0)
    m = neus.NeuS(**kw)
    synthetic.reseed_parameters(m, seed=1)
    return m.to(device)


@pytest.mark.parametrize("M", [2, 33, 64, 257, 2048])
def test_sdf_to_w_kernel_vs_oracle(M):
    rs = np.random.RandomState(M)
    R = 19
    sdf = torch.from_numpy((np.abs(np.linspace(-1, 1, M))[None] * 0.4 - 0.15 + 0.01 * rs.normal(size=(R, M))).astype(np.float32))
    s = 64.0
    _, _, want = (lambda c_a: (c_a[0], c_a[1], oneus.alpha_to_w(c_a[1])))(oneus.sdf_to_alpha(sdf, s))
    w = torch.empty(R, M - 1, device=DEV)
    _lib.check(_lib.get_lib().nr_neus_sdf_to_w(_lib.ptr(sdf.to(DEV)), s, R, M, _lib.ptr(w), _lib.stream_ptr(torch.device(DEV))))
    assert rel_err(w, want) < 2e-6, rel_err(w, want)


@pytest.mark.parametrize("algo", ["direct_use", "direct_more"])
@pytest.mark.parametrize("tier,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_direct_upsamplers_vs_reference(algo, tier, tol):
    g = load_golden("neus_variants_r32.npz")
    neurecon_b200.set_precision(tier)
    m = build_neus(seed=1, device=DEV)
    o, d = synthetic.make_rays(int(g["n_rays"]), shell_radius=2.5, jitter=0.1, seed=int(g["seed"]))
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=True, perturb=False,
                                             upsample_algo=algo, N_nograd_samples=256)
    assert ret["d_final"].shape == (32, 127) and ret["implicit_surface"].shape == (32, 128)
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k], g["%s__%s" % (algo, k)]) < tol, (k, rel_err(ret[k], g["%s__%s" % (algo, k)]))
    if tier == "fp32":   # a sample may hop an inverse-CDF bin where the CDF is flat to an ulp: per-sample tensors by fraction
        for k in ("implicit_surface", "radiance", "visibility_weights", "d_final"):
            assert frac_close(ret[k], g["%s__%s" % (algo, k)], 1e-3) > 0.97, (k, frac_close(ret[k], g["%s__%s" % (algo, k)], 1e-3))
    # stochastic draw: reproducible from torch's CUDA generator, different from the deterministic one
    with torch.no_grad():
        torch.manual_seed(11)
        a = neus.volume_render(o.to(DEV), d.to(DEV), m, detailed_output=False, perturb=True, upsample_algo=algo, N_nograd_samples=256)[0]
        torch.manual_seed(11)
        b = neus.volume_render(o.to(DEV), d.to(DEV), m, detailed_output=False, perturb=True, upsample_algo=algo, N_nograd_samples=256)[0]
    assert torch.equal(a, b) and torch.isfinite(a).all() and not torch.equal(a, rgb)
    with pytest.raises(NotImplementedError):
        neus.volume_render(o.to(DEV), d.to(DEV), m, upsample_algo="nope")


@pytest.mark.parametrize("tier,tol", [("fp32", 1e-5), ("fp16", 1e-2)])
def test_radiance_net_without_view_dirs(tier, tol):
    g = load_golden("neus_variants_r32.npz")
    neurecon_b200.set_precision(tier)
    m = _noview_model()
    assert m.radiance_net.layers[0].weight_v.shape == (256, 3 + 256)        # base.py:336: cat([x, feature]) only
    x = synthetic.make_points(128, extent=1.0, seed=2).to(DEV)
    with torch.no_grad():
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x)
        rad = m.radiance_net.forward(x, None, nab, feat)                   # the reference's call (base.py:372-391)
        rad_q = m.forward_radiance(x, None)                                # fused query (tensor tier: one kernel pair)
    assert rel_err(rad, g["noview__net_radiance"]) < max(tol, 2e-5 if tier == "fp32" else tol)
    assert rel_err(rad_q, g["noview__net_radiance"]) < max(tol, 2e-5 if tier == "fp32" else tol)
    # surface_render(use_view_dirs=False): ray_casting.py:217-230
    o, d = synthetic.make_rays(int(g["n_rays"]), shell_radius=2.5, jitter=0.1, seed=int(g["seed"]))
    col, dep, ex = ray_casting.surface_render(o.to(DEV)[None], d.to(DEV)[None], m, calc_normal=True, batched=True, use_view_dirs=False,
                                              ray_casting_algo="sphere_tracing", ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
    assert torch.equal(ex["mask_surface"][0].cpu(), g["noview__st_mask"])
    assert rel_err(col[0], g["noview__st_color"]) < max(tol, 1e-4)
    # a mismatch between the flag and the network is an error (the reference fails on it too, base.py:379)
    with pytest.raises(ValueError):
        ray_casting.surface_render(o.to(DEV)[None], d.to(DEV)[None], m, use_view_dirs=True, ray_casting_algo="sphere_tracing",
                                   ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
    # volume_render(use_view_dirs=False) (crashes inside batchify_query in the reference): equals the render of a
    # with-view-dirs model whose layer 0 ignores views and normals
    mv = build_neus(seed=1, device=DEV)
    with torch.no_grad():
        l0, r0 = mv.radiance_net.layers[0], m.radiance_net.layers[0]
        v = torch.zeros_like(l0.weight_v)
        v[:, :3] = r0.weight_v[:, :3]
        v[:, -256:] = r0.weight_v[:, 3:]
        l0.weight_v.copy_(v)
        l0.weight_g.copy_(r0.weight_g)
        l0.bias.copy_(r0.bias)
        for a, b in zip(list(mv.radiance_net.layers)[1:], list(m.radiance_net.layers)[1:]):
            a.load_state_dict(b.state_dict())
        mv.implicit_surface.load_state_dict(m.implicit_surface.state_dict())
        mv.ln_s.copy_(m.ln_s)
        want = neus.volume_render(o.to(DEV), d.to(DEV), mv, calc_normal=True, detailed_output=False)[0]
        got = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=False, use_view_dirs=False)[0]
    assert rel_err(got, want) < 1e-6


def test_radiance_net_without_view_dirs_trains():
    """gradients of RadianceNet(use_view_dirs=False).forward under autograd against plain torch on the same parameters"""
    neurecon_b200.set_precision("fp32")
    m = _noview_model()
    rad = m.radiance_net
    rs = np.random.RandomState(0)
    n = 300
    x = torch.from_numpy(rs.uniform(-1, 1, size=(n, 3)).astype(np.float32)).to(DEV)
    feat = torch.from_numpy(rs.normal(size=(n, 256)).astype(np.float32)).to(DEV).requires_grad_(True)
    nab = torch.from_numpy(rs.normal(size=(n, 3)).astype(np.float32)).to(DEV).requires_grad_(True)
    tgt = torch.from_numpy(rs.uniform(size=(n, 3)).astype(np.float32)).to(DEV)

    def torch_forward():
        h = torch.cat([x, feat], dim=-1).double()
        for i, l in enumerate(rad.layers):
            W = (l.weight_v * (l.weight_g / l.weight_v.norm(dim=1, keepdim=True))).double()
            h = h @ W.t() + l.bias.double()
            h = torch.sigmoid(h) if i == len(rad.layers) - 1 else torch.relu(h)
        return h

    params = [p for p in rad.parameters()]
    loss_ref = ((torch_forward() - tgt.double()) ** 2).mean()
    g_ref = torch.autograd.grad(loss_ref, params + [feat])
    out = rad.forward(x, None, nab, feat)
    loss = ((out - tgt) ** 2).mean()
    g = torch.autograd.grad(loss, params + [feat], allow_unused=True)
    assert rel_err(out, torch_forward()) < 1e-5
    for a, b, p in zip(g, g_ref, params + [feat]):
        assert a is not None and a.shape == p.shape
        assert rel_err(a, b) < 1e-4, rel_err(a, b)
