"""aten-level view of a NeuS training step's glue (which torch ops launch the small kernels).  Usage: prof_train_ops.py [rays]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from conftest import build_neus
from neurecon_b200.models.frameworks import neus
from neurecon_b200.utils import synthetic, train_util
if os.environ.get("NEURECON_B200_PRECISION"):
    neurecon_b200.set_precision(os.environ["NEURECON_B200_PRECISION"])
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
opt = train_util.FusedAdam(m.parameters(), lr=5e-4, capturable=True)
o, d = synthetic.make_rays(R, seed=3)
o, d = o.to(dev), d.to(dev)
target = torch.rand(R, 3, device=dev)
mask = torch.rand(R, device=dev) < 0.7
def step():
    opt.zero_grad(set_to_none=False)
    rgb, _, ret = neus.volume_render(o, d, m, detailed_output=True, perturb=True)
    losses = train_util.neus_losses(rgb, target, ret["implicit_nablas"], mask_volume=ret["mask_volume"], target_mask=mask, w_eikonal=0.1, w_mask=1.0)
    losses["total"].backward()
    opt.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
evs = prof.events()
kern = [e for e in evs if e.device_type == torch.autograd.DeviceType.CUDA]
print("kernels: %d, %.3f ms" % (len(kern), sum(e.device_time_total for e in kern) / 1e3))
# group small torch kernels by the python source line that launched them
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    if e.device_type != torch.autograd.DeviceType.CPU or not e.name.startswith("aten::") or e.device_time_total <= 0:
        continue
    if e.cpu_parent is not None and e.cpu_parent.name.startswith("aten::"):
        continue                                     # top-level aten ops only
    st = [s for s in (e.stack or []) if "neurecon_b200" in s]
    key = (e.name, st[0].split("neurecon_b200/")[-1][:60] if st else "?")
    agg[key][0] += 1
    agg[key][1] += e.device_time_total / 1e3
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
print("top-level aten ops with device time: %d calls, %.3f ms" % (sum(v[0] for _, v in rows), sum(v[1] for _, v in rows)))
for (name, where), (c, ms) in rows[:45]:
    print("%8.3f ms x%-3d %-28s %s" % (ms, c, name, where))
