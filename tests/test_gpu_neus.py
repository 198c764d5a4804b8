"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C-ABI
library; the checker is the CPU oracle and the committed golden vectors of the reference.

Tolerances (north_star): sample indices bit-exact given the same CDF and uniforms; composited
outputs <= 1e-4 relative (max|a-b|/max|b|) for the fp32 tier."""
import numpy as np
import pytest
import torch

from conftest import NEUS_CFG, build_neus, cpu_state_dict, frac_close, load_golden, rel_err
from oracle import nets, neus as oneus, sampling
from neurecon_b200.utils import rend_util, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32_tier():
    import neurecon_b200
    neurecon_b200.set_precision("fp32")
    yield
    neurecon_b200.set_precision("fp16")


def test_native_library_is_loaded():
    from neurecon_b200 import _lib
    lib = _lib.get_lib()
    assert lib.nr_version() >= 100
    import ctypes
    sm, smem, maj, mnr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.nr_device_info(ctypes.byref(sm), ctypes.byref(smem), ctypes.byref(maj), ctypes.byref(mnr)) == 0
    assert maj.value == 10, "these kernels are built for sm_100a only"


def test_near_far_bit_exact():
    g = load_golden("sampling.npz")
    o, d = synthetic.make_rays(64, seed=4)
    d = torch.nn.functional.normalize(d, dim=-1)
    near, far = rend_util.near_far_from_sphere(o.to(DEV), d.to(DEV), r=1.0)
    assert near.shape == (64, 1)
    assert torch.equal(near.cpu(), g["near"]) and torch.equal(far.cpu(), g["far"])
    n2, f2 = rend_util.near_far_from_sphere(o.to(DEV), d.to(DEV), r=1.0, keepdim=False)
    assert n2.shape == (64,)


@pytest.mark.parametrize("M,N", [(64, 16), (96, 16), (513, 514), (3584, 64), (2, 1)])
def test_sample_cdf_indices_bit_exact(M, N):
    """Same CDF + same uniforms -> identical (below, above) and identical samples."""
    rs = np.random.RandomState(M * 1000 + N)
    R = 37  # ragged: not a multiple of the rays-per-block
    bins = torch.from_numpy(np.sort(rs.uniform(0, 6, size=(R, M)).astype(np.float32), axis=1))
    w = (rs.uniform(size=(R, M - 1)) ** 6).astype(np.float32)
    w[:, : (M - 1) // 2] = 0
    cdf = torch.from_numpy(np.cumsum(w / np.maximum(w.sum(1, keepdims=True), 1e-8), axis=1).astype(np.float32)) * 0.95
    u = torch.from_numpy(rs.uniform(size=(R, N)).astype(np.float32))
    u[:, 0] = 0.0
    u[:, -1] = 1.0
    for uu in (u, None):
        want, wb, wa = sampling.sample_cdf(bins, cdf, N, det=uu is None, u=uu, return_inds=True)
        got, gb, ga, gcdf = rend_util.sample_cdf(bins.to(DEV), cdf.to(DEV), N, det=uu is None,
                                                 u=None if uu is None else uu.to(DEV), return_details=True)
        assert torch.equal(gb.cpu().long(), wb) and torch.equal(ga.cpu().long(), wa)
        assert torch.equal(got.cpu(), want)
        assert torch.equal(gcdf.cpu()[:, 1:], cdf)


def test_sample_pdf_golden_and_same_cdf_contract():
    g = load_golden("sampling.npz")
    bins, w, u = g["bins"].to(DEV), g["weights"].to(DEV), g["u"].to(DEV)
    N = u.shape[-1]
    for uu, key in ((None, "det"), (u, "sto")):
        got, gb, ga, gcdf = rend_util.sample_pdf(bins, w, N, det=uu is None, u=uu, return_details=True)
        # the kernel's CDF differs from torch.cumsum only by scan order
        assert rel_err(gcdf, sampling.pdf_to_cdf(g["weights"])) < 1e-6
        # bit-exact contract: the oracle's search on the kernel's CDF reproduces indices and samples
        uo = sampling.linspace01(N).expand(bins.shape[0], N).contiguous() if uu is None else g["u"]
        want, wb, wa = sampling.invert_cdf(g["bins"], gcdf.cpu(), uo, return_inds=True)
        assert torch.equal(gb.cpu().long(), wb) and torch.equal(ga.cpu().long(), wa)
        assert torch.equal(got.cpu(), want)
        # vs the reference's own samples: identical except where a uniform falls in a flat CDF
        # run (weights == 0), where the found bin is decided by the last bit of the cumsum
        assert frac_close(got, g[key], 1e-5) > 0.9
    got = rend_util.sample_cdf(bins, g["cdf_in"].to(DEV), N, u=u)
    assert torch.equal(got.cpu(), g["sto_cdf"])


@pytest.mark.parametrize("M,N", [(64, 16), (80, 16), (96, 16), (112, 16), (128, 32), (32, 8), (8, 5), (2, 1), (64, 128)])
@pytest.mark.parametrize("R", [37, 64, 1000])
def test_sample_pdf_thread_per_ray_kernel_returns_the_warp_kernels_bits(M, N, R):
    """Short even rows go to the thread-per-ray kernel (sample_pdf_rows_kernel), which reproduces the warp kernel's
    association order: identical samples, deterministic and with given uniforms, for weights and for a given CDF."""
    rs = np.random.RandomState(M * 131 + N * 7 + R)
    bins = torch.from_numpy(np.sort(rs.uniform(0, 6, size=(R, M)).astype(np.float32), axis=1)).to(DEV)
    w = (rs.uniform(size=(R, M - 1)) ** 6).astype(np.float32)
    w[: R // 2, : (M - 1) // 2] = 0
    w = torch.from_numpy(w).to(DEV)
    u = torch.from_numpy(rs.uniform(size=(R, N)).astype(np.float32))
    u[:, 0], u[:, -1] = 0.0, 1.0
    u = u.to(DEV)
    for uu in (u, None):
        fast = rend_util.sample_pdf(bins, w, N, det=uu is None, u=uu)
        slow, _, _, cdf = rend_util.sample_pdf(bins, w, N, det=uu is None, u=uu, return_details=True)
        assert torch.equal(fast, slow)
        assert torch.equal(rend_util.sample_cdf(bins, cdf[:, 1:].contiguous(), N, det=uu is None, u=uu),
                           rend_util.sample_cdf(bins, cdf[:, 1:].contiguous(), N, det=uu is None, u=uu, return_details=True)[0])
    # unaligned rows (a slice that starts 4 bytes into an allocation) take the scalar staging loops
    flat_b, flat_w = torch.zeros(R * M + 1, device=DEV), torch.zeros(R * (M - 1) + 1, device=DEV)
    flat_b[1:] = bins.reshape(-1); flat_w[1:] = w.reshape(-1)
    fast = rend_util.sample_pdf(flat_b[1:].view(R, M), flat_w[1:].view(R, M - 1), N, det=True)
    assert torch.equal(fast, rend_util.sample_pdf(bins, w, N, det=True, return_details=True)[0])


def test_sample_pdf_empty_and_batched_prefix():
    bins = torch.linspace(0, 1, 8, device=DEV).expand(0, 8)
    out = rend_util.sample_pdf(bins, torch.zeros(0, 7, device=DEV), 4, det=True)
    assert out.shape == (0, 4)
    b = torch.sort(torch.rand(2, 3, 16, device=DEV), dim=-1).values
    w = torch.rand(2, 3, 15, device=DEV)
    out = rend_util.sample_pdf(b, w, 5, det=True)
    want = sampling.sample_pdf(b.cpu(), w.cpu(), 5, det=True)
    assert out.shape == (2, 3, 5) and rel_err(out, want) < 1e-5


def _oracle_layers(sd):
    return (nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", 9),
            nets.layers_from_state_dict(sd, "radiance_net.layers", 5))


@pytest.mark.parametrize("n", [1, 100, 256, 1000])
def test_mlp_fp32_vs_oracle_and_golden(n):
    m = build_neus(seed=1, device=DEV)
    L, Lr = _oracle_layers(cpu_state_dict(m))
    x = synthetic.make_points(n, extent=1.0, seed=2)
    v = torch.nn.functional.normalize(synthetic.make_points(n, extent=1.0, seed=3), dim=-1)
    with torch.no_grad():
        sdf0 = m.implicit_surface.forward(x.to(DEV))
        sdf1, feat1 = m.implicit_surface.forward(x.to(DEV), return_h=True)
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x.to(DEV))
        rad = m.radiance_net.forward(x.to(DEV), v.to(DEV), nab, feat)
    osdf, onab, ofeat = nets.sdf_forward_with_nablas(x, L)
    orad = nets.radiance_forward(x, v, onab, ofeat, Lr, -1, 4)
    assert sdf.shape == (n,) and nab.shape == (n, 3) and feat.shape == (n, 256) and rad.shape == (n, 3)
    for a, b, tol in ((sdf0, osdf, 1e-5), (sdf1, osdf, 1e-5), (sdf, osdf, 1e-5), (feat1, ofeat, 1e-5),
                      (feat, ofeat, 1e-5), (nab, onab, 1e-4), (rad, orad, 1e-5)):
        assert rel_err(a, b) < tol, rel_err(a, b)
    if n == 256:
        g = load_golden("neus_nets_n256.npz")
        assert rel_err(sdf, g["sdf"]) < 1e-5 and rel_err(nab, g["nabla"]) < 1e-4
        assert rel_err(feat, g["feat"]) < 1e-5 and rel_err(rad, g["radiance"]) < 1e-5


def test_mlp_accepts_prefix_shapes_and_empty():
    m = build_neus(seed=1, device=DEV)
    with torch.no_grad():
        x = torch.rand(2, 5, 7, 3, device=DEV) - 0.5
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x)
        assert sdf.shape == (2, 5, 7) and nab.shape == (2, 5, 7, 3) and feat.shape == (2, 5, 7, 256)
        flat = m.implicit_surface.forward(x.reshape(-1, 3))
        assert rel_err(flat.reshape(2, 5, 7), sdf) < 1e-6  # rows are independent of batch position
        again = m.implicit_surface.forward(x)
        assert torch.equal(again.reshape(-1), flat)       # and bit-reproducible
        e = m.implicit_surface.forward(torch.zeros(0, 3, device=DEV))
        assert e.shape == (0,)


def test_neus_upsample_matches_oracle():
    """The 4-iteration up-sampler given the oracle's own SDF callback semantic: same d_all."""
    from neurecon_b200.models.frameworks import neus
    m = build_neus(seed=1, device=DEV)
    sd = cpu_state_dict(m)
    L, _ = _oracle_layers(sd)
    o, d = synthetic.make_rays(33, seed=11)
    dn = torch.nn.functional.normalize(d, dim=-1)
    near, far = sampling.near_far_from_sphere(o, dn, 1.0)
    t = sampling.linspace01(64)
    d_coarse = near * (1 - t) + far * t
    want, _ = oneus.upsample(lambda p: nets.sdf_forward(p, L), o, dn, d_coarse)
    dirs, d_all, pts, d_mid, pts_mid = neus._upsample(m, o.to(DEV), d.to(DEV), 1.0, None, None, 64, 64, 4, False)
    assert rel_err(dirs, dn) < 1e-6
    assert frac_close(d_all, want, 1e-4) > 0.97
    assert torch.equal(d_mid.cpu(), 0.5 * (d_all.cpu()[:, 1:] + d_all.cpu()[:, :-1]))
    assert (d_all[:, 1:] >= d_all[:, :-1]).all()
    assert rel_err(pts, o[:, None, :] + dn[:, None, :] * d_all.cpu()[:, :, None]) < 1e-6


@pytest.mark.parametrize("perturb", [False, True])
def test_neus_upsampler_field_equals_evaluation_at_sorted_samples(perturb):
    """The (sdf, nablas) the up-sampler's merge carries along are forward_with_nablas at the sorted samples
    (neus.py:291, which the render no longer repeats): same points, same arithmetic, hence the same numbers; ties and
    the stable old-before-new order must keep every normal with its depth."""
    from neurecon_b200.models.frameworks import neus
    m = build_neus(seed=1, device=DEV)
    o, d = synthetic.make_rays(257, seed=21)
    torch.manual_seed(5)
    out = neus._upsample(m, o.to(DEV), d.to(DEV), 1.0, None, None, 64, 64, 4, perturb, return_field=True,
                         with_nablas=True)
    dirs, d_all, pts, d_mid, pts_mid, sdf_all, nab_all = out
    with torch.no_grad():
        sdf, nab, _ = m.implicit_surface.forward_with_nablas(pts)
    assert sdf_all.shape == (257, 128) and nab_all.shape == (257, 128, 3)
    assert torch.isfinite(nab_all).all()
    assert rel_err(sdf_all, sdf) < 1e-5 and rel_err(nab_all, nab) < 1e-5   # same points, other batch shape
    # without normals the up-sampler queries the sdf-only program: its field is the evaluation at ITS samples
    torch.manual_seed(5)
    out2 = neus._upsample(m, o.to(DEV), d.to(DEV), 1.0, None, None, 64, 64, 4, perturb, return_field=True)
    with torch.no_grad():
        sdf2 = m.implicit_surface.forward(out2[2])
    assert out2[6] is None and out2[5].shape == (257, 128) and rel_err(out2[5], sdf2) < 1e-5
    assert frac_close(out2[1], d_all, 1e-3) > 0.97


@pytest.mark.parametrize("white_bkgd", [False, True])
def test_neus_composite_vs_oracle(white_bkgd):
    from neurecon_b200.models.frameworks import neus
    rs = np.random.RandomState(3)
    R, M = 45, 128
    d_all = torch.from_numpy(np.sort(rs.uniform(1.5, 3.5, size=(R, M)).astype(np.float32), axis=1))
    sdf = torch.from_numpy((np.abs(np.linspace(-1, 1, M))[None] * 0.5 - 0.2 + 0.02 * rs.normal(size=(R, M))).astype(np.float32))
    nab = torch.from_numpy(rs.normal(size=(R, M, 3)).astype(np.float32))
    nab[0, 0] = 0  # zero vector stays zero under F.normalize
    rad = torch.from_numpy(rs.uniform(size=(R, M - 1, 3)).astype(np.float32))
    s = torch.tensor([37.5])
    want = oneus.composite(sdf, nab, rad, d_all, s, white_bkgd, True)
    d_mid = 0.5 * (d_all[:, 1:] + d_all[:, :-1])
    rgb, depth, acc, normals, cdf, alpha, w = neus._composite(
        sdf.to(DEV), nab.to(DEV), rad.to(DEV), d_mid.to(DEV), s.to(DEV), white_bkgd, True, True)
    for a, b in ((rgb, want["rgb"]), (depth, want["depth_volume"]), (acc, want["mask_volume"]),
                 (normals, want["normals_volume"]), (cdf, want["cdf"]), (alpha, want["alpha"]),
                 (w, want["visibility_weights"])):
        assert rel_err(a, b) < 1e-5, rel_err(a, b)


def test_neus_volume_render_vs_oracle_and_golden():
    from neurecon_b200.models.frameworks import neus
    m = build_neus(seed=1, device=DEV)
    g = load_golden("neus_render_r48.npz")
    o, d = synthetic.make_rays(48, shell_radius=2.5, jitter=0.1, seed=1)
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=True,
                                             perturb=False)
    assert list(ret.keys()) == ["rgb", "depth_volume", "mask_volume", "normals_volume", "implicit_nablas",
                                "implicit_surface", "radiance", "alpha", "cdf", "visibility_weights", "d_final"]
    assert ret["implicit_nablas"].shape == (48, 128, 3) and ret["radiance"].shape == (48, 127, 3)
    assert ret["cdf"].shape == (48, 128) and ret["d_final"].shape == (48, 127)
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k], g[k]) < 1e-4, (k, rel_err(ret[k], g[k]))
    for k in ("implicit_surface", "radiance", "alpha", "visibility_weights", "d_final", "cdf"):
        assert frac_close(ret[k], g[k], 1e-3) > 0.97, (k, frac_close(ret[k], g[k], 1e-3))  # bins may flip
    _, _, want = oneus.volume_render(o, d, cpu_state_dict(m), NEUS_CFG, calc_normal=True)
    for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
        assert rel_err(ret[k], want[k]) < 1e-4, k


def test_neus_volume_render_api_variants():
    from neurecon_b200.models.frameworks import neus
    m = build_neus(seed=1, device=DEV)
    o, d = synthetic.make_rays(40, seed=5)
    o, d = o.to(DEV), d.to(DEV)
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o, d, m, calc_normal=True, detailed_output=False)
        assert list(ret.keys()) == ["rgb", "depth_volume", "mask_volume", "normals_volume"]
        # batched prefix + ray chunking do not change values (rays are independent)
        rgb_b, depth_b, ret_b = neus.volume_render(o.reshape(2, 20, 3), d.reshape(2, 20, 3), m, batched=True,
                                                   calc_normal=True, detailed_output=False, rayschunk=7)
        assert rgb_b.shape == (2, 20, 3) and depth_b.shape == (2, 20)
        assert torch.equal(rgb_b.reshape(40, 3), rgb) and torch.equal(depth_b.reshape(40), depth)
        # near/far bypass, white background, unknown kwargs swallowed like **dummy_kwargs
        rgb_w, _, ret_w = neus.volume_render(o, d, m, white_bkgd=True, near_bypass=1.0, far_bypass=4.0,
                                             detailed_output=False, some_future_flag=3)
        assert torch.isfinite(rgb_w).all() and "normals_volume" not in ret_w
        # stochastic sampling draws from torch's CUDA generator, reproducibly
        torch.manual_seed(3)
        a = neus.volume_render(o, d, m, perturb=True, detailed_output=False)[0]
        torch.manual_seed(3)
        b = neus.volume_render(o, d, m, perturb=True, detailed_output=False)[0]
        assert torch.equal(a, b) and not torch.equal(a, rgb)
        # render through SingleRenderer
        r2 = neus.SingleRenderer(m)(o, d, calc_normal=True, detailed_output=False)[0]
        assert torch.equal(r2, rgb)


@pytest.mark.parametrize("tier,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_neus_nerfpp_background_vs_golden(tier, tol):
    """NeuS without mask: NeRF++ background net + blended compositing (neus.py:303-343)."""
    import neurecon_b200
    from test_oracle_golden import build_neus_bg
    from neurecon_b200.models.frameworks import neus
    neurecon_b200.set_precision(tier)
    m = build_neus_bg(device=DEV)
    g = load_golden("neus_render_nerfpp_r24.npz")
    o, d = synthetic.make_rays(24, shell_radius=2.5, jitter=0.15, seed=5)
    with torch.no_grad():
        rgb, depth, ret = neus.volume_render(o.to(DEV), d.to(DEV), m, calc_normal=True, detailed_output=True,
                                             perturb=False, N_outside=32)
        assert list(ret.keys())[-2:] == ["sigma_out", "radiance_out"]
        assert ret["alpha"].shape == (24, 159) and ret["radiance"].shape == (24, 159, 3) and ret["d_final"].shape == (24, 159)
        # rays whose up-sampling hopped an inverse-CDF bin have moved samples: tight on the others, loose on all
        same = ((ret["d_final"].cpu() - g["d_final"]).abs().amax(-1) < 1e-4) if tier == "fp32" else torch.ones(24, dtype=torch.bool)
        assert same.float().mean() > 0.5
        for k in ("rgb", "depth_volume", "mask_volume", "normals_volume"):
            assert rel_err(ret[k].cpu()[same], g[k][same]) < tol, (k, rel_err(ret[k].cpu()[same], g[k][same]))
            assert rel_err(ret[k], g[k]) < max(tol, 5e-3), (k, rel_err(ret[k], g[k]))
        if tier == "fp32":
            # the NeRF++ net is queried at all 159 samples (d_mid and the outside ones) through a 2^9 embedding, which
            # amplifies 1e-4 sample shifts to 1e-2: compare where the samples agree to rounding
            tight = (ret["d_final"].cpu() - g["d_final"]).abs().amax(-1) < 3e-6
            assert tight.sum() >= 4
            assert rel_err(ret["sigma_out"].cpu()[tight], g["sigma_out"][tight]) < 2e-4
            assert rel_err(ret["radiance_out"].cpu()[tight], g["radiance_out"][tight]) < 2e-4
        pj = neus.volume_render(o.to(DEV), d.to(DEV), m, detailed_output=False, perturb=True, N_outside=32)[0]
        assert torch.isfinite(pj).all()


def test_get_rays_vs_golden():
    from test_oracle_golden import _cams
    g = load_golden("get_rays.npz")
    c2w, K, H, W = _cams()
    ro, rd, sel = rend_util.get_rays(c2w.to(DEV), K.to(DEV), H, W, N_rays=-1)
    assert ro.shape == (2, H * W, 3) and sel.shape == (2, H * W)
    assert torch.equal(ro.cpu(), g["rays_o_all"]) and rel_err(rd, g["rays_d_all"]) < 1e-6
    torch.manual_seed(123)  # same CPU generator draws as the reference (rend_util.py:137-138)
    ro, rd, sel = rend_util.get_rays(c2w.to(DEV), K.to(DEV), H, W, N_rays=50)
    assert torch.equal(sel.cpu(), g["select_inds"])
    assert torch.equal(ro.cpu(), g["rays_o_sel"]) and rel_err(rd, g["rays_d_sel"]) < 1e-6
    # single camera without batch prefix, 3x3 intrinsics, quaternion pose
    ro1, rd1, _ = rend_util.get_rays(c2w[0].to(DEV), K[0, :3, :3].to(DEV), H, W)
    assert ro1.shape == (H * W, 3) and torch.equal(rd1, rend_util.get_rays(c2w[:1].to(DEV), K[:1].to(DEV), H, W)[1][0])
    q = torch.tensor([[0.9, 0.1, -0.3, 0.2, 1.0, 2.0, 3.0]], device=DEV)
    roq, rdq, _ = rend_util.get_rays(q, K[:1].to(DEV), H, W)
    assert torch.allclose(roq[0, 0].cpu(), torch.tensor([1.0, 2.0, 3.0])) and torch.isfinite(rdq).all()


def test_neus_volume_render_is_cuda_graph_capturable():
    """no host synchronisation inside the path.  A whole volume_render (ray setup, 5 up-sampling
    rounds with their network queries, radiance pass, compositing) is captured into one CUDA graph and replayed on new
    rays: same numbers as the eager call."""
    import neurecon_b200
    from neurecon_b200.models.frameworks import neus
    neurecon_b200.set_precision("fp16")       # the fixture restores the module's tier afterwards
    try:
        m = build_neus(seed=1, device=DEV)
        o, d = synthetic.make_rays(1000, seed=3)
        o, d = o.to(DEV), d.to(DEV)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(2):     # warm-up: weight images, workspaces and the allocator's pools exist before the capture
                neus.volume_render(o, d, m, calc_normal=True, detailed_output=False)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                rgb, depth, ret = neus.volume_render(o, d, m, calc_normal=True, detailed_output=False)
            o2, d2 = synthetic.make_rays(1000, seed=4)
            o.copy_(o2.to(DEV))
            d.copy_(d2.to(DEV))
            graph.replay()
            torch.cuda.synchronize()
            rgb_e, depth_e, ret_e = neus.volume_render(o, d, m, calc_normal=True, detailed_output=False)
            torch.cuda.synchronize()
        assert torch.equal(rgb, rgb_e) and torch.equal(depth, depth_e)
        assert torch.equal(ret["normals_volume"], ret_e["normals_volume"])
    finally:
        neurecon_b200.set_precision("fp32")
