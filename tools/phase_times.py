"""Where a NeuS render chunk spends its GPU time: CUDA events around every fused-MLP launch (by mode and size)
against the whole chunk.  Usage: python tools/phase_times.py [rays]"""
import os
import sys
import collections

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from neurecon_b200.models import base  # noqa: E402
from neurecon_b200.models.frameworks import neus  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402
from conftest import build_neus  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = torch.device("cuda:0")
    m = build_neus(seed=1, device=dev)
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.15, seed=5)
    o, d = o.to(dev), d.to(dev)
    events = []
    orig = base.ImplicitSurface._run_umma

    def wrapped(self, x, mode, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(self, x, mode, *a, **k)
        e1.record()
        events.append((mode, x.reshape(-1, 3).shape[0], e0, e1))
        return r
    base.ImplicitSurface._run_umma = wrapped
    with torch.no_grad():
        for it in range(4):
            events.clear()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            neus.volume_render(o, d, m, calc_normal=True, detailed_output=False, rayschunk=65536)
            t1.record()
            torch.cuda.synchronize()
    total = t0.elapsed_time(t1)
    agg = collections.OrderedDict()
    for mode, n, e0, e1 in events:
        k = (mode, n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + e0.elapsed_time(e1))
    s = 0.0
    for (mode, n), (c, t) in agg.items():
        print("%-7s n=%9d x%2d  %8.3f ms  (%6.1f Mpts/s)" % (mode, n, c, t, n * c / t / 1e3))
        s += t
    print("whole render of %d rays: %.3f ms (%.0f rays/s); inside MLP wrappers: %.3f ms (%.1f%%)" % (R, total, R / total * 1e3, s, 100 * s / total))


if __name__ == "__main__":
    main()
