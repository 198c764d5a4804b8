"""Split-precision tier ('fp16x2', csrc/mlp_rev_split.cu): accuracy against the fp64 oracle next to the fp32 and fp16 tiers,
and launch time.  Usage: python tools/check_split.py [n_points_for_timing]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from conftest import build_neus, cpu_state_dict, rel_err  # noqa: E402
from oracle import nets  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402

dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
sd = cpu_state_dict(m)
L = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", 9, dtype=torch.float64)
Lr = nets.layers_from_state_dict(sd, "radiance_net.layers", 5, dtype=torch.float64)
n = 4001
x = synthetic.make_points(n, extent=1.0, seed=2)
v = torch.nn.functional.normalize(synthetic.make_points(n, extent=1.0, seed=3), dim=-1)
osdf, onab, ofeat = nets.sdf_forward_with_nablas(x.double(), L)
orad = nets.radiance_forward(x.double(), v.double(), onab, ofeat, Lr, -1, 4)
for tier in ("fp32", "fp16", "fp16x2"):
    neurecon_b200.set_precision(tier)
    with torch.no_grad():
        sdf0 = m.implicit_surface.forward(x.to(dev))
        sdf1, feat1 = m.implicit_surface.forward(x.to(dev), return_h=True)
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x.to(dev))
        rad, sdf2, nab2 = neurecon_b200.models.base.query_radiance(m.implicit_surface, m.radiance_net, x.to(dev), v.to(dev))
    torch.cuda.synchronize()
    print("%-7s sdf %.2e / %.2e / %.2e / %.2e  feat %.2e / %.2e  nabla %.2e / %.2e  radiance %.2e" % (
        tier, rel_err(sdf0, osdf), rel_err(sdf1, osdf), rel_err(sdf, osdf), rel_err(sdf2, osdf), rel_err(feat1, ofeat),
        rel_err(feat, ofeat), rel_err(nab, onab), rel_err(nab2, onab), rel_err(rad, orad)), flush=True)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
xb = (torch.rand(N, 3, device=dev) - 0.5) * 1.5
vb = torch.nn.functional.normalize(torch.randn(N, 3, device=dev), dim=-1)
for tier in ("fp16", "fp16x2"):
    neurecon_b200.set_precision(tier)
    for name, fn in (("sdf", lambda: m.implicit_surface.forward(xb)),
                     ("sdf+nabla", lambda: m.implicit_surface._run(xb, True, False)),
                     ("sdf+nabla+radiance", lambda: neurecon_b200.models.base.query_radiance(m.implicit_surface, m.radiance_net, xb, vb))):
        with torch.no_grad():
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
        print("%-7s %-20s %8.3f ms per %d points" % (tier, name, e0.elapsed_time(e1) / 5, N), flush=True)
