"""Split-precision tensor tier (precision 'fp16x2', csrc/mlp_rev_split.cu): the SDF net on (hi, lo) fp16 operand pairs, held to
the fp32 tier's bars (north_star: <= 1e-4 end to end) against the oracle and the reference's golden vectors."""
import numpy as np
import pytest
import torch

import neurecon_b200
from neurecon_b200 import _lib
from neurecon_b200.models.base import query_radiance
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import synthetic
from oracle import nets, neus as oneus
from conftest import NEUS_CFG, build_neus, cpu_state_dict, frac_close, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _split_tier():
    neurecon_b200.set_precision("fp16x2")
    yield
    neurecon_b200.set_precision("fp16")


def _oracle_layers(sd):
    return (nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", 9),
            nets.layers_from_state_dict(sd, "radiance_net.layers", 5))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 129, 256, 1000, 20001])
def test_split_mlp_vs_oracle_and_golden(n):
    """every entry point of the SDF net: sdf only, sdf + feature, sdf + normals + feature, fused with the radiance pass"""
    m = build_neus(seed=1, device=DEV)
    L, Lr = _oracle_layers(cpu_state_dict(m))
    x = synthetic.make_points(n, extent=1.0, seed=2)
    v = torch.nn.functional.normalize(synthetic.make_points(n, extent=1.0, seed=3), dim=-1)
    with torch.no_grad():
        sdf0 = m.implicit_surface.forward(x.to(DEV))
        sdf1, feat1 = m.implicit_surface.forward(x.to(DEV), return_h=True)
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x.to(DEV))
        rad, sdf2, nab2 = query_radiance(m.implicit_surface, m.radiance_net, x.to(DEV), v.to(DEV))
    osdf, onab, ofeat = nets.sdf_forward_with_nablas(x, L)
    orad = nets.radiance_forward(x, v, onab, ofeat, Lr, -1, 4)
    assert sdf.shape == (n,) and nab.shape == (n, 3) and feat.shape == (n, 256) and rad.shape == (n, 3)
    # sdf / feature: the fp32 tier's 1e-5; normals: its 1e-4; radiance: the radiance net itself runs on plain fp16 operands
    for a, b, tol in ((sdf0, osdf, 1e-5), (sdf1, osdf, 1e-5), (sdf, osdf, 1e-5), (sdf2, osdf, 1e-5), (feat1, ofeat, 1e-5),
                      (feat, ofeat, 1e-5), (nab, onab, 1e-4), (nab2, onab, 1e-4), (rad, orad, 1e-4)):
        assert rel_err(a, b) < tol, rel_err(a, b)
    assert torch.equal(sdf0, sdf) and torch.equal(sdf1, sdf) and torch.equal(sdf2, sdf)   # one forward sweep, whatever follows it
    if n == 256:
        g = load_golden("neus_nets_n256.npz")
        assert rel_err(sdf, g["sdf"]) < 1e-5 and rel_err(nab, g["nabla"]) < 1e-4
        assert rel_err(feat, g["feat"]) < 1e-5 and rel_err(rad, g["radiance"]) < 1e-4


def test_split_kernel_is_reproducible_and_position_independent():
    m = build_neus(seed=1, device=DEV)
    x = (torch.rand(3000, 3, device=DEV) - 0.5) * 1.6
    with torch.no_grad():
        a = m.implicit_surface.forward_with_nablas(x)
        b = m.implicit_surface.forward_with_nablas(x)
        c = m.implicit_surface.forward_with_nablas(x[37:1900])
    for u, w, z in zip(a, b, c):
        assert torch.equal(u, w)
        assert rel_err(u[37:1900], z) < 1e-6   # a point's result does not depend on its tile slot or column


def test_split_neus_render_vs_golden_and_oracle():
    m = build_neus(seed=1, device=DEV)
    g = load_golden("neus_render_r48.npz")
    o, d = synthetic.make_rays(48, shell_radius=2.5, jitter=0.1, seed=1)
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=True, perturb=False)
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k], g[k]) < 1e-4, (k, rel_err(ret[k], g[k]))
    for k in ("implicit_surface", "radiance", "alpha", "visibility_weights", "d_final", "cdf"):
        assert frac_close(ret[k], g[k], 1e-3) > 0.97, (k, frac_close(ret[k], g[k], 1e-3))  # bins may flip
    _, _, want = oneus.volume_render(o, d, cpu_state_dict(m), NEUS_CFG, calc_normal=True)
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k], want[k]) < 1e-4, k


def test_split_neus_nerfpp_background_vs_golden():
    from conftest import build_neus_bg
    m = build_neus_bg(device=DEV)
    g = load_golden("neus_render_nerfpp_r24.npz")
    o, d = synthetic.make_rays(24, shell_radius=2.5, jitter=0.15, seed=5)
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=True, perturb=False,
                                             N_outside=32)
    # rays whose up-sampling hopped an inverse-CDF bin have moved samples (the fp32 tier's test has the same clause,
    # test_gpu_neus.py): tight on the others, loose on all.  The NeRF++ background net runs on plain fp16 operands.
    same = (ret["d_final"].cpu() - g["d_final"]).abs().amax(-1) < 1e-4
    assert same.float().mean() > 0.5
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k].cpu()[same], g[k][same]) < 2e-4, (k, rel_err(ret[k].cpu()[same], g[k][same]))
        assert rel_err(ret[k], g[k]) < 5e-3, (k, rel_err(ret[k], g[k]))


def test_split_training_gradients_vs_fp32_tier():
    """'fp16x2' under autograd: the reverse-mode training GEMMs with split-precision FORWARD sweeps (nr_gemm16_split).  Weight
    gradients of sdf + eikonal + feature losses within 1e-2 of the fp32 tier's (plain fp16 operands: ~1e-1)."""
    from neurecon_b200.models import autograd, autograd_rev
    assert _lib.tensor_tier() and _lib.split_tier() and autograd._tc() and autograd_rev.split_forward()
    m = build_neus(seed=1, device=DEV)
    x = (torch.rand(4000, 3, device=DEV) - 0.5) * 1.6

    def grads():
        m.zero_grad()
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x)
        loss = (sdf ** 2).mean() + ((nab.norm(dim=-1) - 1) ** 2).mean() + 1e-2 * (feat ** 2).mean()
        loss.backward()
        return float(loss), [p.grad.clone() for p in m.implicit_surface.parameters()]

    l2, g2 = grads()
    neurecon_b200.set_precision("fp32")
    l1, g1 = grads()
    assert abs(l1 - l2) < 1e-4 * abs(l1)
    worst = max(rel_err(a, b) for a, b in zip(g2, g1))
    assert worst < 1e-2, worst


@pytest.mark.parametrize("N,K", [(256, 256), (217, 256), (256, 39), (257, 256), (256, 289), (3, 256), (64, 64)])
def test_gemm16_split_vs_fp64(N, K):
    """nr_gemm16_split (the training path's split-precision forward GEMM) against a float64 product: linear output, softplus
    output with its derivative, and the (hi, lo) pair it writes for the next layer"""
    from neurecon_b200.models import autograd_rev as ar
    rs = np.random.RandomState(N * 1000 + K)
    n = 1000
    kp = (K + 63) // 64 * 64
    A = torch.from_numpy(rs.normal(size=(n, K)).astype(np.float32)).to(DEV)
    W = torch.from_numpy((rs.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)).to(DEV)
    b = torch.from_numpy(rs.normal(size=(N,)).astype(np.float32) * 0.1).to(DEV)
    rows = torch.zeros(n, 2 * kp, dtype=torch.float16, device=DEV)
    hi = A.half()
    rows[:, :K] = hi
    rows[:, kp:kp + K] = (A - hi.float()).half()
    Wp = torch.zeros(N, (K + 3) & ~3, device=DEV)
    Wp[:, :K] = W
    img = ar._PackedSplit(Wp.contiguous(), N, K)
    want = A.double() @ W.double().t() + b.double()
    n16 = (N + 15) & ~15
    y = torch.empty(n, n16, device=DEV)
    ar._gemm16_split(rows, img, b, n, N, K, y, 0, 0, ar.G_LINEAR)
    assert rel_err(y[:, :N], want) < 3e-6, rel_err(y[:, :N], want)
    lo_off = (N + 63) // 64 * 64
    out = torch.empty(n, 2 * lo_off, dtype=torch.float16, device=DEV)
    S = torch.empty(n, lo_off, dtype=torch.float16, device=DEV)
    ar._gemm16_split(rows, img, b, n, N, K, out, 1, lo_off, ar.G_SOFTPLUS, out2=S)
    sp = torch.nn.functional.softplus(want, beta=100)
    got = out[:, :N].double() + out[:, lo_off:lo_off + N].double()
    assert rel_err(got, sp) < 3e-6, rel_err(got, sp)
    assert rel_err(out[:, :N], sp) < 1e-3                                # the hi part alone is an fp16 number
    assert rel_err(S[:, :N], torch.sigmoid(100 * want)) < 1e-3           # softplus' is kept in fp16


@pytest.mark.parametrize("rows,N,K", [(1000, 256, 256), (4096, 217, 256), (70000, 256, 39), (513, 256, 295), (64, 3, 256)])
def test_gemm16_tn_and_tn2_vs_fp64(rows, N, K):
    """nr_gemm16_tn (dW += scale G^T X on fp16 rows) and nr_gemm16_tn2 (two such products through one accumulator and one
    pass of atomics: the SDF layers' dW = zb^T h + p^T gb) against float64 products of the same fp16 operands"""
    lib = _lib.get_lib()
    rs = np.random.RandomState(rows + 7 * N + K)
    ldg, ldx = (N + 63) // 64 * 64, (K + 63) // 64 * 64

    def operand(cols, ld):
        t = torch.zeros(rows, ld, dtype=torch.float16, device=DEV)
        t[:, :cols] = torch.from_numpy(rs.normal(size=(rows, cols)).astype(np.float32)).to(DEV).half()
        return t

    G, X, G2, X2 = operand(N, ldg), operand(K, ldx), operand(N, ldg), operand(K, ldx)
    ldw = (K + 3) & ~3
    st = torch.cuda.current_stream().cuda_stream
    base = torch.from_numpy(rs.normal(size=(N, ldw)).astype(np.float32)).to(DEV)
    one = G[:, :N].double().t() @ X[:, :K].double()
    two = one + G2[:, :N].double().t() @ X2[:, :K].double()
    dW = base.clone()
    _lib.check(lib.nr_gemm16_tn(_lib.ptr(G), ldg, _lib.ptr(X), ldx, rows, N, K, _lib.ptr(dW), ldw, 0.5, st), "gemm16_tn")
    assert rel_err(dW[:, :K], base[:, :K].double() + 0.5 * one) < 2e-6
    dW2 = base.clone()
    _lib.check(lib.nr_gemm16_tn2(_lib.ptr(G), ldg, _lib.ptr(X), ldx, _lib.ptr(G2), ldg, _lib.ptr(X2), ldx, rows, N, K,
                                 _lib.ptr(dW2), ldw, 0.5, st), "gemm16_tn2")
    assert rel_err(dW2[:, :K], base[:, :K].double() + 0.5 * two) < 2e-6
    assert torch.equal(dW2[:, K:], base[:, K:])                          # pad columns of dW untouched
