"""Oracle restatement of the reference networks (TEST INFRASTRUCTURE ONLY).

Functional style: a network is a list of ``(W_eff, b)`` tensors where
``W_eff = g * v / ||v||_row`` is the weight-normalised matrix the reference
recomputes every forward (models/base.py:226-227, old-style weight_norm dim=0).
"""
import math
import torch
import torch.nn.functional as F


def embed(x, multires):
    """Embedder.forward (models/base.py:46-64): [x, sin(2^0 x), cos(2^0 x), ...]."""
    if multires < 0:
        return x
    out = [x]
    for k in range(multires):
        f = float(2.0 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def effective_weight(g, v):
    """weight_norm(dim=0): W[o,:] = g[o] * v[o,:] / ||v[o,:]||_2."""
    return v * (g / v.norm(dim=1, keepdim=True))


def layers_from_state_dict(sd, prefix, n_layers, weight_norm=True, dtype=None):
    """Collect [(W_eff, b)] for ``prefix.{i}`` from a reference-layout state_dict."""
    out = []
    for i in range(n_layers):
        if weight_norm:
            W = effective_weight(sd[f"{prefix}.{i}.weight_g"], sd[f"{prefix}.{i}.weight_v"])
        else:
            W = sd[f"{prefix}.{i}.weight"]
        b = sd[f"{prefix}.{i}.bias"]
        if dtype is not None:
            W, b = W.to(dtype), b.to(dtype)
        out.append((W, b))
    return out


def softplus100(x):
    """nn.Softplus(beta=100) (threshold=20) used at models/base.py:202."""
    return F.softplus(x, beta=100)


def sdf_forward(x, layers, multires=6, skips=(4,), return_h=False):
    """ImplicitSurface.forward (models/base.py:243-263), W_geo_feat > 0 layout."""
    pe = embed(x, multires)
    h = pe
    D = len(layers) - 1
    for i in range(D):
        if i in skips:
            h = torch.cat([h, pe], dim=-1) / math.sqrt(2)
        W, b = layers[i]
        h = softplus100(F.linear(h, W, b))
    W, b = layers[D]
    out = F.linear(h, W, b)
    sdf, feat = out[..., 0], out[..., 1:]
    return (sdf, feat) if return_h else sdf


def sdf_forward_with_nablas(x, layers, multires=6, skips=(4,)):
    """ImplicitSurface.forward_with_nablas (models/base.py:265-282), inference
    semantics (everything detached), nabla via autograd like the reference."""
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        sdf, feat = sdf_forward(xr, layers, multires, skips, return_h=True)
        nabla = torch.autograd.grad(sdf, xr, torch.ones_like(sdf))[0]
    return sdf.detach(), nabla.detach(), feat.detach()


def sdf_forward_with_nablas_analytic(x, layers, multires=6, skips=(4,)):
    """Same quantity by explicit forward-mode differentiation (the formulation the
    CUDA kernels use); cross-checks the autograd version in tests."""
    pe = embed(x, multires)
    # tangents of the embedding w.r.t. x_c, c = 0..2 : [..., 3, pe_dim]
    eye = torch.eye(3, dtype=x.dtype)
    t_parts = [eye.expand(*x.shape[:-1], 3, 3)]
    for k in range(max(multires, 0)):
        f = float(2.0 ** k)
        t_parts.append(eye * (f * torch.cos(x * f)).unsqueeze(-2))
        t_parts.append(eye * (-f * torch.sin(x * f)).unsqueeze(-2))
    tpe = torch.cat(t_parts, dim=-1) if multires >= 0 else eye.expand(*x.shape[:-1], 3, 3)
    h, t = pe, tpe
    D = len(layers) - 1
    for i in range(D):
        if i in skips:
            h = torch.cat([h, pe], dim=-1) / math.sqrt(2)
            t = torch.cat([t, tpe], dim=-1) / math.sqrt(2)
        W, b = layers[i]
        z = F.linear(h, W, b)
        h = softplus100(z)
        t = torch.sigmoid(100.0 * z).unsqueeze(-2) * F.linear(t, W)
    W, b = layers[D]
    out = F.linear(h, W, b)
    tout = F.linear(t, W)
    return out[..., 0], tout[..., 0], out[..., 1:]


def radiance_forward(x, view_dirs, normals, feat, layers, multires=-1, multires_view=4):
    """RadianceNet.forward (models/base.py:372-391), use_view_dirs=True, no skips."""
    inp = torch.cat([embed(x, multires), embed(view_dirs, multires_view), normals, feat], dim=-1)
    h = inp
    D = len(layers) - 1
    for i in range(D):
        W, b = layers[i]
        h = torch.relu(F.linear(h, W, b))
    W, b = layers[D]
    return torch.sigmoid(F.linear(h, W, b))


def nerf_forward(x4, views, sd, prefix="nerf_outside", multires=10, multires_view=4, skips=(4,), D=8):
    """NeRF.forward (NeRF++ background; models/base.py:426-453), use_view_dirs=True."""
    pts = embed(x4, multires)
    v = embed(views, multires_view)
    h = pts
    for i in range(D):
        h = torch.relu(F.linear(h, sd[f"{prefix}.pts_linears.{i}.weight"], sd[f"{prefix}.pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([pts, h], dim=-1)
    sigma = F.linear(h, sd[f"{prefix}.alpha_linear.weight"], sd[f"{prefix}.alpha_linear.bias"])
    feature = F.linear(h, sd[f"{prefix}.feature_linear.weight"], sd[f"{prefix}.feature_linear.bias"])
    h = torch.cat([feature, v], dim=-1)
    h = torch.relu(F.linear(h, sd[f"{prefix}.views_linears.0.weight"], sd[f"{prefix}.views_linears.0.bias"]))
    rgb = torch.sigmoid(F.linear(h, sd[f"{prefix}.rgb_linear.weight"], sd[f"{prefix}.rgb_linear.bias"]))
    return sigma.squeeze(-1), rgb
