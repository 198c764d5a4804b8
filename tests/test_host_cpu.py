"""CPU: the C-ABI library builds, loads and exports every symbol include/neurecon_b200.h declares;
the host-side mirror of the reference interface behaves (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, build_neus


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "neurecon_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from neurecon_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    header = _header_symbols()
    assert header, "no symbols parsed from the header"
    for name in header:
        assert hasattr(lib, name), "library lacks %s" % name
    assert set(header) == set(_lib.declared_symbols()), (
        set(header) ^ set(_lib.declared_symbols()))
    assert lib.nr_version() >= 100


def test_error_convention():
    from neurecon_b200 import _lib
    lib = _lib.get_lib()
    # invalid arguments are reported through the return code + nr_last_error, never by throwing
    rc = lib.nr_sample_pdf(None, None, None, 4, 8, 4, 0, 1e-5, None, None, None, None, None)  # R=4, null data
    assert rc == -1
    assert "null" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sample_pdf")


def test_cpu_tensors_raise_no_fallback():
    from neurecon_b200.utils import rend_util
    m = build_neus()
    x = torch.zeros(4, 3)
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.implicit_surface.forward(x)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rend_util.near_far_from_sphere(x, x)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            rend_util.sample_pdf(torch.zeros(2, 8), torch.zeros(2, 7), 4, det=True)


def test_unknown_upsample_algo_raises_like_reference():
    from neurecon_b200.models.frameworks import neus
    m = build_neus()
    with pytest.raises(NotImplementedError):
        neus.volume_render(torch.zeros(2, 3), torch.ones(2, 3), m, upsample_algo="nope")


def test_batchify_query_shapes():
    from neurecon_b200.utils.train_util import batchify_query
    x = torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 5, 7, 3)

    def fn(p):
        return p.sum(-1), {"twice": p * 2}

    s, d = batchify_query(fn, x, chunk=11, dim_batchify=1)
    assert torch.equal(s, x.sum(-1)) and torch.equal(d["twice"], x * 2)
    s0 = batchify_query(lambda p: p[..., 0], x[0], chunk=4, dim_batchify=0)
    assert torch.equal(s0, x[0, ..., 0])


def test_reverse_mode_program_layout_on_cpu():
    """Host side of csrc/mlp_rev.cu (no GPU): the weight image holds W_l^T chunks for the backward sweep, the program is
    hidden x L, [feature], sdf row, backward x (L-1), Jacobian; every chunk range lies inside the image, the transposed
    chunks unswizzle back to the weights, rev_ok() refuses what the kernel's stash cannot hold."""
    from neurecon_b200 import _lib, umma_pack
    m = build_neus()
    net = m.implicit_surface._umma_net(None)
    P = net.program("rev", want_feat=True)
    epi = [P.steps[i].epi for i in range(P.n_steps)]
    H, F, S, B, N = umma_pack.EPI_HIDDEN, umma_pack.EPI_FEAT, umma_pack.EPI_SDF_OUT, umma_pack.EPI_BWD, umma_pack.EPI_NABLA
    assert P.reverse == 1 and P.tangents == 0
    assert epi == [H] * 8 + [F, S] + [B] * 7 + [N]
    assert [P.steps[i].sig_slot for i in range(8)] == list(range(8))
    assert P.steps[9].sig_slot == 7 and [P.steps[i].sig_slot for i in range(10, 17)] == [6, 5, 4, 3, 2, 1, 0]
    assert [P.steps[i].pe_fill for i in range(10, 17)] == [0, 0, 0, 1, 0, 0, 0]       # backward of the skip layer (l = 4)
    assert P.steps[13].out_rows == 217 and P.steps[17].out_rows == 39 and P.steps[17].pe_fill == 1
    n_chunks = net.image.shape[0]
    for i in range(P.n_steps):
        st = P.steps[i]
        assert 0 <= st.chunk_begin and st.chunk_begin + st.n_mt * (st.k_steps // 4) <= n_chunks
        assert st.n_cols == 128
    # w_sdf table: the sdf row's weights, fp32, zero padded to 256
    aux = net.bias[P.steps[9].aux_off:P.steps[9].aux_off + 256]
    from neurecon_b200.models.base import _effective_weight
    w_last = _effective_weight(m.implicit_surface.surface_fc_layers[8]).detach().float()
    assert torch.allclose(aux, w_last[0], atol=0, rtol=0)
    # chunk (k-chunk 0, M-tile 0) of the first backward step = rows 0..127 x k 0..63 of W_7^T, K-major 128-byte swizzle
    W7t = _effective_weight(m.implicit_surface.surface_fc_layers[7]).detach().float().t().contiguous()
    chunk = net.image[P.steps[10].chunk_begin].view(torch.float16)
    idx = umma_pack._a_tile_index().reshape(128, 64)
    got = chunk[idx.reshape(-1)].reshape(128, 64).float()
    assert torch.equal(got, W7t[:128, :64].half().float())
    assert net.rev_ok(True) and net.rev_ok(False)
    # 'rev_img': no feature step, the last hidden step also writes the radiance operand image
    Pi = net.program("rev_img")
    assert Pi.n_steps == 17 and Pi.steps[7].to_rad == 1 and Pi.steps[8].epi == S
    assert _lib.NR_UMMA_MAX_STEPS >= P.n_steps


def test_exclusive_cumprod_backward_matches_torch_incl_zeros():
    """models/autograd.py: the transmittance product with a host-sync-free backward (torch.cumprod's reads a flag on the
    host, which a CUDA-graph capture of the training step cannot do); zeros as torch handles them."""
    import torch
    from neurecon_b200.models.autograd import exclusive_cumprod
    g = torch.Generator().manual_seed(0)
    for with_zeros in (False, True):
        p = torch.rand(7, 33, dtype=torch.float64, generator=g)
        if with_zeros:
            p[1, 5] = 0; p[2, 0] = 0; p[3, 32] = 0; p[4, 3] = 0; p[4, 9] = 0; p[5, 31] = 0
        w = torch.randn(7, 33, dtype=torch.float64, generator=g)
        a, b = p.clone().requires_grad_(), p.clone().requires_grad_()
        Ta = torch.cumprod(torch.cat([torch.ones_like(a[..., :1]), a], -1), -1)[..., :-1]
        Tb = exclusive_cumprod(b)
        assert torch.equal(Ta, Tb)
        (Ta * w).sum().backward()
        (Tb * w).sum().backward()
        assert (a.grad - b.grad).abs().max() < 1e-12


def test_captured_step_refuses_cpu_tensors():
    """train_util.CapturedStep is a CUDA-graph capture: on CPU tensors it raises instead of running the step eagerly."""
    import pytest
    import torch
    from neurecon_b200.utils import train_util
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        train_util.CapturedStep(lambda x: x * 2, (torch.zeros(4),))


def test_training_weights_match_the_oracle_incl_gradients():
    """neus.sdf_to_alpha / alpha_to_w (the differentiable tensor ops of the training render, neus.py:21-70) against the
    oracle's restatement on the CPU: same weights bit for bit, same gradient w.r.t. the sdf values and s (the oracle
    differentiates torch.cumprod, the product path its own host-sync-free backward)."""
    import torch
    from neurecon_b200.models.frameworks import neus
    from oracle import neus as oneus
    g = torch.Generator().manual_seed(3)
    sdf = (torch.rand(9, 128, generator=g, dtype=torch.float64) - 0.4) * 0.5
    sdf[0] = torch.linspace(0.5, -0.5, 128, dtype=torch.float64)          # a clean surface crossing: alpha reaches ~1
    wgt = torch.randn(9, 127, generator=g, dtype=torch.float64)
    outs = []
    for mod in (oneus, neus):
        x = sdf.clone().requires_grad_()
        s = torch.tensor([64.0], dtype=torch.float64, requires_grad=True)
        _, alpha = mod.sdf_to_alpha(x, s)
        w = mod.alpha_to_w(alpha)
        (w * wgt).sum().backward()
        outs.append((w.detach(), x.grad, s.grad))
    assert torch.equal(outs[0][0], outs[1][0])
    assert (outs[0][1] - outs[1][1]).abs().max() < 1e-10 * outs[0][1].abs().max()
    assert (outs[0][2] - outs[1][2]).abs().max() < 1e-10 * outs[0][2].abs().max()
