"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle.ref_loader) on deterministic synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each .npz stores the inputs that are not reproducible from a seed (none: rays and weights are
re-derived from numpy seeds by neurecon_b200.utils.synthetic) and the reference outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def npz(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def golden_neus(ref):
    torch.manual_seed(0)
    m = ref.neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=1)
    R = 48
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.1, seed=1)
    with torch.no_grad():
        rgb, depth, ret = ref.neus.volume_render(o, d, m, calc_normal=True, detailed_output=True, perturb=False)
    keep = ["rgb", "depth_volume", "mask_volume", "normals_volume", "implicit_surface", "implicit_nablas",
            "radiance", "alpha", "cdf", "visibility_weights", "d_final"]
    npz("neus_render_r48.npz", seed=1, n_rays=R, **{k: ret[k] for k in keep})

    # NeuS without mask: NeRF++ background (configs/neus_nomask.yaml: N_outside = 32)
    torch.manual_seed(0)
    mb = ref.neus.NeuS(**dict(synthetic.NEUS_MODEL_KWARGS, use_outside_nerf=True))
    synthetic.reseed_parameters(mb, seed=5)
    ob, db = synthetic.make_rays(24, shell_radius=2.5, jitter=0.15, seed=5)
    with torch.no_grad():
        _, _, retb = ref.neus.volume_render(ob, db, mb, calc_normal=True, detailed_output=True, perturb=False, N_outside=32)
    npz("neus_render_nerfpp_r24.npz", seed=5, **{k: retb[k] for k in keep + ["sigma_out", "radiance_out"]})

    # networks on fixed points
    x = synthetic.make_points(256, extent=1.0, seed=2)
    v = torch.nn.functional.normalize(synthetic.make_points(256, extent=1.0, seed=3), dim=-1)
    with torch.no_grad():
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x)
        rad = m.radiance_net.forward(x, v, nab, feat)
    npz("neus_nets_n256.npz", seed=1, sdf=sdf, nabla=nab, feat=feat, radiance=rad)


def golden_sampling(ref):
    rs = np.random.RandomState(5)
    R, M, N = 32, 96, 16
    bins = np.sort(rs.uniform(0.5, 4.0, size=(R, M)).astype(np.float32), axis=1)
    w = (rs.uniform(size=(R, M - 1)) ** 4).astype(np.float32)
    w[:, : M // 3] = 0.0  # long empty run: exercises the denom < eps branch
    u = rs.uniform(size=(R, N)).astype(np.float32)
    tb, tw = torch.from_numpy(bins), torch.from_numpy(w)
    det = ref.rend_util.sample_pdf(tb, tw, N, det=True)
    # replay the stochastic branch with known u by monkeypatching torch.rand
    real_rand = torch.rand
    try:
        torch.rand = lambda *a, **k: torch.from_numpy(u)
        sto = ref.rend_util.sample_pdf(tb, tw, N, det=False)
        cdf_in = torch.cumsum(tw / tw.sum(-1, keepdim=True), -1) * 0.9  # tops out below 1
        sto_cdf = ref.rend_util.sample_cdf(tb, cdf_in, N, det=False)
    finally:
        torch.rand = real_rand
    det_cdf = ref.rend_util.sample_cdf(tb, cdf_in, N, det=True)
    o, d = synthetic.make_rays(64, seed=4)
    d = torch.nn.functional.normalize(d, dim=-1)
    near, far = ref.rend_util.near_far_from_sphere(o, d, r=1.0)
    # get_rays: two cameras, skewed intrinsics, random pixel selection from a seeded CPU generator
    H, W = 24, 32
    c2w = torch.stack([synthetic.look_at_pose([2.0, 1.0, 1.5]), synthetic.look_at_pose([-1.5, 2.2, 0.4])])
    K = synthetic.pinhole_intrinsics(H, W, skew=0.7)[None].expand(2, 4, 4).contiguous()
    ro_all, rd_all, _ = ref.rend_util.get_rays(c2w, K, H, W, N_rays=-1)
    torch.manual_seed(123)
    ro_sel, rd_sel, sel = ref.rend_util.get_rays(c2w, K, H, W, N_rays=50)
    npz("get_rays.npz", rays_o_all=ro_all, rays_d_all=rd_all, rays_o_sel=ro_sel.contiguous(), rays_d_sel=rd_sel, select_inds=sel.contiguous())
    npz("sampling.npz", bins=bins, weights=w, u=u, det=det, sto=sto, cdf_in=cdf_in, sto_cdf=sto_cdf,
        det_cdf=det_cdf, near=near, far=far)


def golden_volsdf(ref):
    for tag, beta_init, npp in (("b0p1", 0.1, False), ("b0p01", 0.01, False), ("b0p003", 0.003, False),
                                ("b0p01_nerfpp", 0.01, True)):
        kw = dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=beta_init, use_nerfplusplus=npp)
        torch.manual_seed(0)
        m = ref.volsdf.VolSDF(**kw)
        synthetic.reseed_parameters(m, seed=3)
        R = 24
        o, d = synthetic.make_rays(R, shell_radius=3.0 / 1.1, jitter=0.1, seed=3)
        with torch.no_grad():
            _, _, ret = ref.volsdf.volume_render(
                o, d, m, calc_normal=True, detailed_output=True, perturb=False, near=0.0, far=6.0,
                obj_bounding_radius=3.0, max_upsample_steps=5 if npp else 6, use_nerfplusplus=npp, N_outside=32)
        keep = ["rgb", "depth_volume", "mask_volume", "normals_volume", "beta_map", "iter_usage", "d_vals", "sigma",
                "visibility_weights"]
        npz("volsdf_render_%s_r24.npz" % tag, seed=3, beta_init=beta_init, nerfpp=int(npp), **{k: ret[k] for k in keep})

    # error_bound / sdf_to_sigma on synthetic 1-D data (incl. overflow -> inf entries)
    rs = np.random.RandomState(11)
    R, M = 16, 200
    dv = np.sort(rs.uniform(0, 6, size=(R, M)).astype(np.float32), axis=1)
    sdf = (np.abs(dv - 3.0) - 1.0 + 0.05 * rs.normal(size=(R, M))).astype(np.float32)
    outs = {}
    for i, beta in enumerate((0.5, 0.05, 0.002)):
        b = torch.tensor(beta)
        outs["bound_%d" % i] = ref.volsdf.error_bound(torch.from_numpy(dv), torch.from_numpy(sdf), 1.0 / b, b)
        outs["sigma_%d" % i] = ref.volsdf.sdf_to_sigma(torch.from_numpy(sdf), 1.0 / b, b)
    npz("volsdf_error_bound.npz", d_vals=dv, sdf=sdf, betas=np.array([0.5, 0.05, 0.002], dtype=np.float32), **outs)


def golden_unisurf(ref):
    torch.manual_seed(0)
    m = ref.unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=4)
    R = 40
    o, d = synthetic.make_rays(R, shell_radius=3.0, jitter=0.25, seed=4)
    with torch.no_grad():
        _, _, ret = ref.unisurf.volume_render(o[None], d[None], m, batched=True, calc_normal=True, detailed_output=True,
                                              perturb=False, logit_tau=0.0, radius_of_interest=4.0, interval=1.0)
    keep = ["rgb", "depth_volume", "mask_volume", "normals_volume", "surface_points", "mask_surface", "depth_surface",
            "implicit_surface", "alpha", "visibility_weights"]
    npz("unisurf_render_r40.npz", seed=4, **{k: ret[k][0] for k in keep})


def golden_surface_render(ref):
    """ray_casting.py:163-263: sphere tracing and root finding + surface rendering through the reference's own code"""
    torch.manual_seed(0)
    m = ref.neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=1)
    R = 48
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.12, seed=9)
    dn = torch.nn.functional.normalize(d, dim=-1)
    out = {}
    with torch.no_grad():
        dp, pts, mask = ref.ray_casting.sphere_tracing_surface_points(m.implicit_surface, o[None], dn[None], near=0.0, far=5.0, N_iters=20)
        out.update(st_d=dp[0], st_pts=pts[0], st_mask=mask[0])
        col, dep, ex = ref.ray_casting.surface_render(o[None], d[None], m, calc_normal=True, batched=True,
                                                      ray_casting_algo="sphere_tracing", ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
        out.update(st_color=col[0], st_depth=dep[0], st_normals=ex["normals_surface"][0], st_render_mask=ex["mask_surface"][0])
        near, far = ref.rend_util.near_far_from_sphere(o[None], dn[None], r=1.0, keepdim=False)
        col, dep, ex = ref.ray_casting.surface_render(o[None], d[None], m, calc_normal=True, batched=True, ray_casting_algo="root_finding",
                                                      ray_casting_cfgs=dict(near=near, far=far, logit_tau=0.0, N_steps=256, N_secant_steps=8))
        out.update(rf_color=col[0], rf_depth=dep[0], rf_normals=ex["normals_surface"][0], rf_mask=ex["mask_surface"][0], near=near[0], far=far[0])
    npz("surface_render_r48.npz", seed=9, n_rays=R, **out)


def golden_neus_variants(ref):
    """The reference behaviours outside the shipped configs: the two NeRF-like up-samplers (neus.py:216-243) and a
    radiance net without view directions (base.py:335-336,383-384; volume_render(use_view_dirs=False), neus.py:190-193)."""
    R = 32
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.1, seed=7)
    keep = ["rgb", "depth_volume", "mask_volume", "normals_volume", "implicit_surface", "radiance", "visibility_weights", "d_final"]
    out = {}
    torch.manual_seed(0)
    m = ref.neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=1)
    for algo in ("direct_use", "direct_more"):
        with torch.no_grad():
            _, _, ret = ref.neus.volume_render(o, d, m, calc_normal=True, detailed_output=True, perturb=False,
                                               upsample_algo=algo, N_nograd_samples=256)
        out.update({"%s__%s" % (algo, k): ret[k] for k in keep})
    kw = dict(synthetic.NEUS_MODEL_KWARGS, radiance_cfg=dict(synthetic.NEUS_MODEL_KWARGS["radiance_cfg"], use_view_dirs=False))
    torch.manual_seed(0)
    mv = ref.neus.NeuS(**kw)
    synthetic.reseed_parameters(mv, seed=1)
    # (volume_render(use_view_dirs=False) itself crashes in the reference: batchify_query flattens the None it passes as
    # view_dirs, train_util.py:27.  What works there: the network's forward and surface_render, ray_casting.py:217-230.)
    x = synthetic.make_points(128, extent=1.0, seed=2)
    with torch.no_grad():
        sdf, nab, feat = mv.implicit_surface.forward_with_nablas(x)
        out["noview__net_radiance"] = mv.radiance_net.forward(x, None, nab, feat)
        col, dep, ex = ref.ray_casting.surface_render(o[None], d[None], mv, calc_normal=True, batched=True, use_view_dirs=False,
                                                      ray_casting_algo="sphere_tracing", ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
        out.update(noview__st_color=col[0], noview__st_depth=dep[0], noview__st_mask=ex["mask_surface"][0])
    npz("neus_variants_r32.npz", seed=7, n_rays=R, **out)


if __name__ == "__main__":
    ref = ref_loader.load()
    if "variants" in sys.argv[1:]:
        golden_neus_variants(ref)
        raise SystemExit(0)
    if "surface" in sys.argv[1:]:
        golden_surface_render(ref)
        raise SystemExit(0)
    golden_neus(ref)
    golden_sampling(ref)
    golden_volsdf(ref)
    golden_unisurf(ref)
    golden_surface_render(ref)
    golden_neus_variants(ref)
