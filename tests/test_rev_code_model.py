"""CPU model of what csrc/mlp_rev.cu does to the normal: the reverse sweep of ImplicitSurface.forward_with_nablas
(models/base.py:265-282) written out by hand in fp64, once with the exact softplus' and once with the kernel's 8-bit
codes 128 + round(254 (s - 1/2)).  Pins (i) that the hand-written sweep (skip connection, embedding Jacobian) IS the
autograd normal of the oracle and (ii) the accuracy cost of the codes that DESIGN.md section 4.1a quotes."""
import math

import torch

from conftest import build_neus, cpu_state_dict
from oracle import nets
from neurecon_b200.utils import synthetic


def _sweep(x, L, quantise):
    def pe(p):
        out = [p]
        for q in range(6):
            out += [torch.sin(p * 2 ** q), torch.cos(p * 2 ** q)]
        return torch.cat(out, -1)

    e = pe(x)
    h, sig = e, []
    for l, (W, b) in enumerate(L[:-1]):
        if l == 4:
            h = torch.cat([h, e], -1) / math.sqrt(2)
        z = h @ W.double().t() + b.double()
        s = torch.sigmoid(100 * z)
        if quantise:
            s = 0.5 + torch.round(254 * (s - 0.5)) / 254
        sig.append(s)
        h = torch.nn.functional.softplus(z, beta=100)
    g = L[-1][0].double()[0:1].expand(x.shape[0], -1) * sig[7]
    gpe = 0
    for l in range(7, 0, -1):
        gi = g @ L[l][0].double()
        if l == 4:
            gi = gi / math.sqrt(2)
            gpe, gi = gi[:, 217:], gi[:, :217]
        g = gi * sig[l - 1]
    gpe = gpe + g @ L[0][0].double()
    nab = gpe[:, 0:3].clone()
    for q in range(6):
        f = 2.0 ** q
        nab += gpe[:, 3 + 6 * q:6 + 6 * q] * f * torch.cos(x * f) - gpe[:, 6 + 6 * q:9 + 6 * q] * f * torch.sin(x * f)
    return nab


def test_reverse_sweep_is_the_autograd_normal_and_codes_cost_2e_3():
    m = build_neus(seed=1, device="cpu")
    L = nets.layers_from_state_dict(cpu_state_dict(m), "implicit_surface.surface_fc_layers", 9)
    x = synthetic.make_points(6000, extent=1.0, seed=2)
    _, want, _ = nets.sdf_forward_with_nablas(x, L)
    exact = _sweep(x.double(), L, False)
    assert ((exact - want.double()).abs().max() / want.abs().max()).item() < 1e-5
    coded = _sweep(x.double(), L, True)
    err = coded - exact
    assert (err.abs().max() / exact.abs().max()).item() < 3e-3
    assert (err.pow(2).mean().sqrt() / exact.pow(2).mean().sqrt()).item() < 1.5e-3
