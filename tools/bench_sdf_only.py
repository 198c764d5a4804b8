"""sdf-only queries of n points through ImplicitSurface.forward (fp16 tier): the forward sweep of mlp_rev_kernel
(NEURECON_B200_SDF_VIA_REV=1, default) against mlp_umma_kernel's 'sdf' program (=0).  Usage: python tools/bench_sdf_only.py [n]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200
from neurecon_b200.models import base
from conftest import build_neus, rel_err
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
x = (torch.rand(n, 3, device=dev) - 0.5) * 1.5
out = {}
for via in (True, False, True, False):
    base._SDF_VIA_REV = via
    with torch.no_grad():
        for _ in range(3):
            sdf = m.implicit_surface.forward(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            sdf = m.implicit_surface.forward(x)
        e1.record()
        torch.cuda.synchronize()
    out[via] = sdf
    print("sdf only via %s: %.3f ms per %d points" % ("mlp_rev forward sweep" if via else "mlp_umma 'sdf'        ", e0.elapsed_time(e1) / 20, n), flush=True)
print("difference between the two: %.2e (max |a - b| / max |b|)" % rel_err(out[True], out[False]))
