"""Surface rendering (the reference's "100x faster" mode, ray_casting.py:163-263): sphere tracing, root finding and
``surface_render`` against tests/golden/surface_render_r48.npz, which tests/golden/make_golden.py wrote by running the
UNMODIFIED reference's own functions."""
import pytest
import torch

import neurecon_b200
from neurecon_b200.models import ray_casting
from neurecon_b200.utils import rend_util, synthetic
from conftest import build_neus, load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tier,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_surface_render_matches_reference(tier, tol):
    z = load_golden("surface_render_r48.npz")
    neurecon_b200.set_precision(tier)
    try:
        m = build_neus(seed=1, device="cuda")
        o, d = synthetic.make_rays(int(z["n_rays"]), shell_radius=2.5, jitter=0.12, seed=int(z["seed"]))
        o, d = o.cuda(), d.cuda()
        dn = torch.nn.functional.normalize(d, dim=-1)
        # ---- sphere tracing (ray_casting.py:163-184)
        dp, pts, mask = ray_casting.sphere_tracing_surface_points(m.implicit_surface, o[None], dn[None], near=0.0, far=5.0, N_iters=20)
        assert torch.equal(mask[0].cpu(), z["st_mask"])
        assert 0.3 < z["st_mask"].float().mean() < 1.0, "the test rays should mostly hit"
        hit = z["st_mask"]
        assert rel_err(dp[0].cpu()[hit], z["st_d"][hit]) < tol and rel_err(pts[0].cpu()[hit], z["st_pts"][hit]) < tol
        # ---- surface_render with sphere tracing (ray_casting.py:187-263)
        col, dep, ex = ray_casting.surface_render(o[None], d[None], m, calc_normal=True, batched=True, ray_casting_algo="sphere_tracing",
                                                  ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
        assert torch.equal(ex["mask_surface"][0].cpu(), z["st_render_mask"])
        assert rel_err(col[0], z["st_color"]) < tol, rel_err(col[0], z["st_color"])
        assert rel_err(ex["normals_surface"][0], z["st_normals"]) < tol
        assert rel_err(dep[0].cpu()[hit], z["st_depth"][hit]) < tol
        # ---- surface_render with root finding (ray_casting.py:35-160)
        near, far = rend_util.near_far_from_sphere(o[None], dn[None], r=1.0, keepdim=False)
        col, dep, ex = ray_casting.surface_render(o[None], d[None], m, calc_normal=True, batched=True, ray_casting_algo="root_finding",
                                                  ray_casting_cfgs=dict(near=near, far=far, logit_tau=0.0, N_steps=256, N_secant_steps=8))
        assert torch.equal(ex["mask_surface"][0].cpu(), z["rf_mask"])
        assert rel_err(col[0], z["rf_color"]) < tol, rel_err(col[0], z["rf_color"])
        assert rel_err(ex["normals_surface"][0], z["rf_normals"]) < tol
        rf_hit = z["rf_mask"]                                            # rays without a root carry inf / far (ray_casting.py:153-157)
        assert torch.equal(torch.isinf(dep[0].cpu()), torch.isinf(z["rf_depth"]))
        assert rel_err(dep[0].cpu()[rf_hit], z["rf_depth"][rf_hit]) < tol
    finally:
        neurecon_b200.set_precision("fp16")
