import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, neurecon_b200
from neurecon_b200.models import base, ray_casting
from neurecon_b200.utils import synthetic
from conftest import build_neus, load_golden, rel_err
z = load_golden("surface_render_r48.npz")
def run(tier, via, R=None, seed=None):
    neurecon_b200.set_precision(tier); base._SDF_VIA_REV = via
    m = build_neus(seed=1, device="cuda")
    o, d = synthetic.make_rays(R or int(z["n_rays"]), shell_radius=2.5, jitter=0.12, seed=seed or int(z["seed"]))
    o, d = o.cuda(), d.cuda()
    col, dep, ex = ray_casting.surface_render(o[None], d[None], m, calc_normal=True, batched=True, ray_casting_algo="sphere_tracing",
                                              ray_casting_cfgs=dict(near=0.0, far=5.0, N_iters=20))
    return col[0], ex["normals_surface"][0], ex["mask_surface"][0]
for via in (False, True):
    c, nrm, msk = run("fp16", via)
    print("golden 48 rays, via_rev=%s: colour %.3e normals %.3e" % (via, rel_err(c, z["st_color"]), rel_err(nrm, z["st_normals"])))
c32, n32, m32 = run("fp32", False, 4096, 7)
for via in (False, True):
    c, nrm, msk = run("fp16", via, 4096, 7)
    same = (msk == m32)
    e = ((nrm - n32).abs().amax(-1))[same & m32]
    print("4096 rays vs fp32 tier, via_rev=%s: mask agreement %.4f, normals max err %.3e, 99.9%% %.3e, rms %.3e" % (
        via, same.float().mean().item(), e.max().item(), e.quantile(0.999).item(), e.pow(2).mean().sqrt().item()))
