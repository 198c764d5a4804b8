import sys, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import neurecon_b200
from neurecon_b200.models import autograd as ag
from neurecon_b200 import _lib
from conftest import build_neus, rel_err
from neurecon_b200.utils import synthetic
DEV = "cuda:0"
n = 300
x = synthetic.make_points(n, extent=0.9, seed=31).to(DEV)
v = F.normalize(synthetic.make_points(n, extent=1.0, seed=32), dim=-1).to(DEV)
g = torch.Generator().manual_seed(5)
c_sdf, c_feat, c_rgb = torch.randn(n, generator=g).to(DEV), (torch.randn(n, 256, generator=g) * 0.01).to(DEV), torch.randn(n, 3, generator=g).to(DEV)
def run(fwd_tc, bwd_tc, tn_tc):
    m = build_neus(seed=1, device=DEV)
    orig_gemm, orig_tn, orig_tc = ag._gemm, ag._gemm_tn, ag._tc
    def gemm(*a, grad=False, **k):
        ag._tc = (lambda: bwd_tc) if grad else (lambda: fwd_tc)
        try: return orig_gemm(*a, grad=grad, **k)
        finally: ag._tc = orig_tc
    def tn(*a, **k):
        ag._tc = lambda: tn_tc
        try: return orig_tn(*a, **k)
        finally: ag._tc = orig_tc
    ag._gemm, ag._gemm_tn = gemm, tn
    try:
        sdf, nab, feat = m.implicit_surface.forward_with_nablas(x)
        rgb = m.radiance_net.forward(x, v, nab, feat)
        eik = ((nab.norm(dim=-1) - 1.0) ** 2).mean()
        loss = (sdf * c_sdf).mean() + (feat * c_feat).sum() / n + (rgb * c_rgb).mean() + 0.1 * eik
        loss.backward()
    finally:
        ag._gemm, ag._gemm_tn = orig_gemm, orig_tn
    return {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
neurecon_b200.set_precision("fp16")
ref = run(False, False, False)
for name, cfg in (("fwd only", (True, False, False)), ("bwd only", (False, True, False)), ("tn only", (False, False, True)), ("all", (True, True, True))):
    got = run(*cfg)
    errs = {k: rel_err(got[k], ref[k]) for k in ref}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(name, [(k.replace("implicit_surface.surface_fc_layers", "sdf").replace("radiance_net.layers", "rad"), round(e, 4)) for k, e in worst])
