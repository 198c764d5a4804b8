"""Kernel-time breakdown of a NeuS training step (torch.profiler / CUPTI).  Usage: prof_train.py [rays]"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from conftest import build_neus
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import synthetic
if os.environ.get("NEURECON_B200_PRECISION"):
    neurecon_b200.set_precision(os.environ["NEURECON_B200_PRECISION"])
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
opt = torch.optim.Adam(m.parameters(), lr=5e-4)
o, d = synthetic.make_rays(R, seed=3)
o, d = o.to(dev), d.to(dev)
target = torch.rand(R, 3, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    rgb, _, ret = neus.volume_render(o, d, m, detailed_output=True, perturb=True)
    nn_ = ret["implicit_nablas"].norm(dim=-1)
    loss = F.l1_loss(rgb, target) + 0.1 * F.mse_loss(nn_, torch.ones_like(nn_)) \
        + F.binary_cross_entropy(ret["mask_volume"].clamp(1e-3, 1 - 1e-3), torch.ones(R, device=dev))
    loss.backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
# kernels only (the autograd-function and aten entries of key_averages() nest the kernels they launch)
rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages()
        if e.device_time_total > 0 and ("kernel" in e.key.lower() or e.key.startswith("void ") or "Memcpy" in e.key or "Memset" in e.key)]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("[%s] device time per step %.2f ms" % (neurecon_b200.get_precision(), tot))
for k, ms, c in rows[:22]:
    print("%8.3f ms %5.1f%% x%-4d %s" % (ms, 100 * ms / tot, c, k[:90]))
