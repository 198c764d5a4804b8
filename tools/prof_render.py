"""Kernel-time breakdown of a render (torch.profiler / CUPTI).  Usage: prof_render.py neus|volsdf|volsdf_nerfpp|unisurf [rays] [precision]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from conftest import build_neus, build_unisurf, build_volsdf  # noqa: E402
from neurecon_b200.models.frameworks import neus, unisurf, volsdf  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "neus"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
if len(sys.argv) > 3:
    neurecon_b200.set_precision(sys.argv[3])
dev = torch.device("cuda:0")
if which == "neus":
    m = build_neus(seed=1, device=dev)
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.1, seed=1)
    fn = lambda: neus.volume_render(o, d, m, calc_normal=True, detailed_output=False, perturb=False)
elif which.startswith("volsdf"):
    npp = which.endswith("nerfpp")
    m = build_volsdf(0.01, npp, device=dev)
    o, d = synthetic.make_rays(R, shell_radius=3.0 / 1.1, jitter=0.1, seed=3)
    fn = lambda: volsdf.volume_render(o, d, m, calc_normal=True, detailed_output=False, perturb=False, near=0.0, far=6.0,
                                      obj_bounding_radius=3.0, max_upsample_steps=5 if npp else 6, use_nerfplusplus=npp, N_outside=32)
else:
    m = build_unisurf(device=dev)
    o, d = synthetic.make_rays(R, shell_radius=3.0, jitter=0.25, seed=4)
    fn = lambda: unisurf.volume_render(o[None], d[None], m, batched=True, calc_normal=True, detailed_output=False, perturb=False)
o, d = o.to(dev), d.to(dev)
with torch.no_grad():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages()
        if e.device_time_total > 0 and ("kernel" in e.key.lower() or e.key.startswith("void ") or "Memcpy" in e.key or "Memset" in e.key)]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print("[%s, %s, %d rays] %.2f ms per render (CUDA events), kernels %.2f ms, %d launches" % (
    which, neurecon_b200.get_precision(), R, wall, tot, sum(r[2] for r in rows)))
for k, ms, c in rows[:14]:
    print("%9.3f ms %5.1f%% x%-4d %s" % (ms, 100 * ms / tot, c, k[:100]))

if os.environ.get("NR_PROF_EACH"):
    evs = [e for e in prof.events() if e.device_time_total > 0 and ("mlp_" in e.name or "fine_iter" in e.name)]
    evs.sort(key=lambda e: e.time_range.start)
    for e in evs:
        print("   %9.3f ms  %s" % (e.device_time_total / 1e3, e.name[:70]))

if os.environ.get("NR_PROF_TIMELINE"):
    # device-side timeline: start, duration and the idle gap before every kernel / copy of the render
    from torch.autograd import DeviceType
    evs = [e for e in prof.events() if e.device_type == DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, prev_end = evs[0].time_range.start, None
    for e in evs:
        st, en = e.time_range.start, e.time_range.end
        gap = (st - prev_end) if prev_end is not None else 0.0
        print("   +%9.3f ms  dur %8.3f ms  gap %7.3f ms  %s" % ((st - t0) / 1e3, (en - st) / 1e3, gap / 1e3, e.name[:60]))
        prev_end = en if prev_end is None else max(prev_end, en)
