"""Launch the round-2 kernels a few times each (for ncu): the split-precision SDF kernel, the NeuS up-sampler, VolSDF's beta
iteration, marching cubes.  Usage: python tools/prof_r2.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from conftest import build_neus, build_volsdf  # noqa: E402
from neurecon_b200.models.frameworks import neus, volsdf  # noqa: E402
from neurecon_b200.utils import mesh_util, synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
x = (torch.rand(1 << 19, 3, device=dev) - 0.5) * 1.5
with torch.no_grad():
    neurecon_b200.set_precision("fp16x2")
    for _ in range(2):
        m.implicit_surface._run(x, True, False)                      # mlp_rev_split_kernel (sdf + normals)
    neurecon_b200.set_precision("fp16")
    o, d = synthetic.make_rays(65536, shell_radius=2.5, jitter=0.1, seed=1)
    neus.volume_render(o.to(dev), d.to(dev), m, calc_normal=True, detailed_output=False)      # neus_upsample_kernel x5
    mv = build_volsdf(0.01, False, device=dev)
    o, d = synthetic.make_rays(16384, shell_radius=3.0 / 1.1, jitter=0.1, seed=3)
    volsdf.volume_render(o.to(dev), d.to(dev), mv, detailed_output=False, near=0.0, far=6.0, obj_bounding_radius=3.0,
                         max_upsample_steps=6)                       # volsdf_fine_iter_kernel
    g = torch.linspace(-1, 1, 512, device=dev)
    vol = torch.sqrt(g[:, None, None] ** 2 + g[None, :, None] ** 2 + g[None, None, :] ** 2) - 0.6
    v, f = mesh_util.marching_cubes(vol, 0.0, (2.0 / 512,) * 3)      # mc_count_kernel, mc_generate_kernel
torch.cuda.synchronize()
print("done", v.shape, f.shape)
