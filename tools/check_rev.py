"""Reverse-mode normals (csrc/mlp_rev.cu) against the forward-mode tangent kernel and the fp64 oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from neurecon_b200.models import base  # noqa: E402
from conftest import build_neus, cpu_state_dict, rel_err  # noqa: E402
from oracle import nets  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402

dev = "cuda:0"
for tier in ("fp16", "bf16"):
    neurecon_b200.set_precision(tier)
    m = build_neus(seed=1, device=dev)
    for n in (1, 127, 128, 129, 1000, 40000):
        x = synthetic.make_points(n, extent=1.0, seed=2)
        v = torch.nn.functional.normalize(synthetic.make_points(n, extent=1.0, seed=3), dim=-1)
        sd = cpu_state_dict(m)
        L = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", 9)
        Lr = nets.layers_from_state_dict(sd, "radiance_net.layers", 5)
        osdf, onab, ofeat = nets.sdf_forward_with_nablas(x, L)
        orad = nets.radiance_forward(x, v, onab, ofeat, Lr, -1, 4)
        out = {}
        for which in ("reverse", "forward"):
            base._REVERSE_NABLAS = which == "reverse"
            with torch.no_grad():
                sdf, nab, feat = m.implicit_surface.forward_with_nablas(x.to(dev))
                rgb, sdf3, nab3 = base.query_radiance(m.implicit_surface, m.radiance_net, x.to(dev), v.to(dev))
            torch.cuda.synchronize()
            out[which] = dict(sdf=rel_err(sdf, osdf), nab=rel_err(nab, onab), feat=rel_err(feat, ofeat),
                              rgb=rel_err(rgb, orad), sdf3=rel_err(sdf3, osdf), nab3=rel_err(nab3, onab))
        print(tier, n, "reverse", {k: "%.2e" % v for k, v in out["reverse"].items()}, flush=True)
        print(tier, n, "forward", {k: "%.2e" % v for k, v in out["forward"].items()}, flush=True)
