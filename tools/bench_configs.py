"""Secondary measurements for the other BASELINE.json configs (the headline stays bench.py):
VolSDF 1024 rays (beta=0.1 and 0.01), UNISURF 2048 rays, VolSDF+NeRF++ 65536 rays, NeuS+NeRF++ 16384 rays,
dense SDF grid 256^3 (sdf only / with normals).  Prints one JSON line."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from conftest import build_neus, build_neus_bg, build_unisurf, build_volsdf
from neurecon_b200.models.frameworks import neus, unisurf, volsdf
from neurecon_b200.utils import mesh_util, synthetic

dev = torch.device("cuda:0")
out = {"precision": neurecon_b200.get_precision()}


def timeit(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    for beta in (0.1, 0.01):
        m = build_volsdf(beta, False, device=dev)
        o, d = synthetic.make_rays(1024, shell_radius=3.0 / 1.1, jitter=0.1, seed=3)
        o, d = o.to(dev), d.to(dev)
        ms = timeit(lambda: volsdf.volume_render(o, d, m, detailed_output=False, calc_normal=True, max_upsample_steps=6))
        it = volsdf.volume_render(o, d, m, detailed_output=True, max_upsample_steps=6)[2]["iter_usage"]
        out["volsdf_1024rays_beta%g" % beta] = {"ms": ms, "rays_per_s": 1024 / ms * 1e3, "mean_upsample_iters": it.clamp_min(0).mean().item()}
    m = build_unisurf(device=dev)
    o, d = synthetic.make_rays(2048, shell_radius=3.0, jitter=0.25, seed=4)
    o, d = o[None].to(dev), d[None].to(dev)
    ms = timeit(lambda: unisurf.volume_render(o, d, m, batched=True, detailed_output=False, calc_normal=True))
    out["unisurf_2048rays"] = {"ms": ms, "rays_per_s": 2048 / ms * 1e3}
    m = build_volsdf(0.01, True, device=dev)
    o, d = synthetic.make_rays(65536, shell_radius=3.0 / 1.1, jitter=0.1, seed=3)
    o, d = o.to(dev), d.to(dev)
    ms = timeit(lambda: volsdf.volume_render(o, d, m, detailed_output=False, calc_normal=True, max_upsample_steps=5,
                                             use_nerfplusplus=True, N_outside=32), reps=2, warm=1)
    out["volsdf_nerfpp_65536rays"] = {"ms": ms, "rays_per_s": 65536 / ms * 1e3}
    m = build_neus_bg(device=dev)
    o, d = synthetic.make_rays(16384, shell_radius=2.5, jitter=0.15, seed=5)
    o, d = o.to(dev), d.to(dev)
    ms = timeit(lambda: neus.volume_render(o, d, m, detailed_output=False, calc_normal=True, N_outside=32), reps=2, warm=1)
    out["neus_nerfpp_16384rays"] = {"ms": ms, "rays_per_s": 16384 / ms * 1e3}
    m = build_neus(seed=1, device=dev)
    N = 256
    ms = timeit(lambda: mesh_util.query_sdf_grid(m.implicit_surface, N=N, plane_range=(0, N)), reps=2, warm=1)
    out["grid_%d3_sdf" % N] = {"ms": ms, "queries_per_s": N ** 3 / ms * 1e3}
    ms = timeit(lambda: mesh_util.query_sdf_grid(m.implicit_surface, N=N, plane_range=(0, N), with_nablas=True), reps=2, warm=1)
    out["grid_%d3_sdf_nablas" % N] = {"ms": ms, "queries_per_s": N ** 3 / ms * 1e3}
print(json.dumps(out))
