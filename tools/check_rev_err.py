"""Errors of the reverse-mode kernel (csrc/mlp_rev.cu) and of the forward-mode tangent kernel against the oracle's fp32
autograd, and of one against the other (fp16 tier).  Usage: python tools/check_rev_err.py [n]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from neurecon_b200.models import base
from neurecon_b200.utils import synthetic
from oracle import nets
from conftest import build_neus, cpu_state_dict, rel_err

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
dev = "cuda:0"
neurecon_b200.set_precision("fp16")
m = build_neus(seed=1, device=dev)
x = synthetic.make_points(n, extent=1.0, seed=11)
sd = cpu_state_dict(m)
L = nets.layers_from_state_dict(sd, "implicit_surface.surface_fc_layers", 9)
osdf, onab, ofeat = nets.sdf_forward_with_nablas(x, L)
out = {}
for rev in (True, False):
    base._REVERSE_NABLAS = rev
    with torch.no_grad():
        out[rev] = m.implicit_surface.forward_with_nablas(x.to(dev))
for rev in (True, False):
    sdf, nab, feat = out[rev]
    rms = lambda a, b: ((a.double().cpu() - b.double()).pow(2).mean().sqrt() / b.double().abs().max()).item()
    print("%s: max rel err  sdf %.2e  nabla %.2e  feat %.2e   rms  sdf %.2e  nabla %.2e  feat %.2e" % (
        "reverse-mode" if rev else "tangent     ", rel_err(sdf, osdf), rel_err(nab, onab), rel_err(feat, ofeat),
        rms(sdf, osdf), rms(nab, onab), rms(feat, ofeat)))
print("reverse vs tangent: sdf %.2e  nabla %.2e  feat %.2e" % (rel_err(out[True][0], out[False][0]), rel_err(out[True][1], out[False][1]),
                                                             rel_err(out[True][2], out[False][2])))
