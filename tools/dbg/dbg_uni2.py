import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import neurecon_b200
from neurecon_b200.utils import synthetic
from neurecon_b200.models.frameworks import unisurf
from neurecon_b200.models.composite import UnisurfComposite
from oracle import unisurf as ouni
import test_gpu_train_golden as T
z = T._golden("train_unisurf_r48.npz")
neurecon_b200.set_precision("fp32")
torch.manual_seed(0)
m = unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS); synthetic.reseed_parameters(m, seed=4); m = m.cuda()
ro = z["rays_o"][0].cuda(); rd = torch.nn.functional.normalize(z["rays_d"][0].cuda(), dim=-1); d = z["d_all"].cuda()
pts = ro[:, None] + rd[:, None] * d[..., None]
with torch.no_grad():
    rad, logits, nab = m.forward(pts.reshape(1, -1, 3), rd[:, None].expand(-1, 96, -1).reshape(1, -1, 3))
rad = rad.reshape(-1, 96, 3); logits = logits.reshape(-1, 96)
g = torch.Generator().manual_seed(3)
target = torch.rand(48, 3, generator=g).cuda()
def torch_ref(dtype):
    x = logits.to(dtype).clone().requires_grad_(); c = rad.to(dtype)
    r = ouni.composite(x, c, None, d.to(dtype), calc_normal=False)
    (r["rgb"] - target.to(dtype)).abs().mean().backward()
    return x.grad
g64 = torch_ref(torch.float64); g32 = torch_ref(torch.float32)
x = logits.clone().requires_grad_()
out = UnisurfComposite.apply(x, None, rad, d, False, False, True)
(out[0] - target).abs().mean().backward()
gk = x.grad
mx = g64.abs().max()
print("elementwise max err / max|g|: torch fp32 %.3e, kernel %.3e" % (((g32.double() - g64).abs().max() / mx).item(), ((gk.double() - g64).abs().max() / mx).item()))
print("sum: fp64 %.6e torch fp32 %.6e kernel %.6e ; sum|g| %.3e" % (g64.sum().item(), g32.double().sum().item(), gk.double().sum().item(), g64.abs().sum().item()))
e = (gk.double() - g64).abs()
i = int(e.argmax()); print("worst element ray %d sample %d: fp64 %.6e kernel %.6e torch32 %.6e" % (i // 96, i % 96, g64.flatten()[i], gk.flatten()[i], g32.flatten()[i]))
print("per-ray sum errors kernel:", ((gk.double() - g64).sum(-1).abs().max() / mx).item(), " torch32:", ((g32.double() - g64).sum(-1).abs().max() / mx).item())
