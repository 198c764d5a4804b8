"""Deterministic synthetic workloads: random-init weights with the
reference's init distributions and inward-looking rays on a camera shell.

Everything is drawn from ``numpy.random.RandomState`` (a frozen, platform-independent stream)
rather than from torch's generator, so the build container, the GPU box and the committed golden
vectors all see bit-identical weights and rays for a given seed.
"""
import math

import numpy as np
import torch

NEUS_MODEL_KWARGS = dict(  # configs/neus.yaml:18-41 as get_model builds them (neus.py:494-523)
    variance_init=0.05, speed_factor=10.0, W_geo_feat=256, obj_bounding_radius=1.0,
    surface_cfg=dict(D=8, W=256, skips=[4], radius_init=0.5, embed_multires=6, geometric_init=True),
    radiance_cfg=dict(D=4, W=256, skips=[], embed_multires=-1, embed_multires_view=4, use_view_dirs=True),
)
VOLSDF_MODEL_KWARGS = dict(  # configs/volsdf.yaml as volsdf.get_model builds them
    beta_init=0.1, speed_factor=10.0, W_geo_feat=256, obj_bounding_radius=3.0, use_nerfplusplus=False,
    surface_cfg=dict(D=8, W=256, skips=[4], radius_init=1.0, embed_multires=6, geometric_init=True),
    radiance_cfg=dict(D=4, W=256, skips=[], embed_multires=-1, embed_multires_view=-1, use_view_dirs=True),
)
UNISURF_MODEL_KWARGS = dict(  # configs/unisurf.yaml as unisurf.get_model builds them
    W_geo_feat=256,
    surface_cfg=dict(D=8, W=256, skips=[4], radius_init=1.0, embed_multires=6, geometric_init=True),
    radiance_cfg=dict(D=4, W=256, skips=[], embed_multires=-1, embed_multires_view=-1, use_view_dirs=True),
)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def reseed_parameters(model, seed=0, perturb=0.003):
    """Overwrite every parameter of a NeuS / VolSDF / UNISURF style model (reference or
    neurecon_b200 -- the state_dict layouts are identical) from a numpy stream.

    The surface net gets the sphere initialisation of models/base.py:207-224 plus a small
    dense perturbation (so that the positional-encoding columns, the skip columns and the
    biases -- all zero at init -- take part in parity tests); the radiance / NeRF++ nets get
    nn.Linear's default U(-1/sqrt(in), 1/sqrt(in))."""
    rs = np.random.RandomState(seed)
    sd = model.state_dict()
    new = {}
    surf = model.implicit_surface
    D, skips, r_init = surf.D, list(surf.skips), surf.radius_init
    pe = sd["implicit_surface.surface_fc_layers.0.weight_v"].shape[1]  # embedding width
    has_pe = pe > 3
    for l in range(D + 1):
        pre = "implicit_surface.surface_fc_layers.%d." % l
        out_d, in_d = sd[pre + "weight_v"].shape
        if l == D:
            w = rs.normal(math.sqrt(math.pi) / math.sqrt(in_d), 0.0001, size=(out_d, in_d))
            b = np.full(out_d, -r_init)
        else:
            w = rs.normal(0.0, math.sqrt(2) / math.sqrt(out_d), size=(out_d, in_d))
            b = np.zeros(out_d)
            if has_pe and l == 0:
                w[:, 3:] = 0.0
            elif has_pe and l in skips:
                w[:, -(pe - 3):] = 0.0
        g = np.linalg.norm(w, axis=1, keepdims=True)
        w = w + perturb * rs.normal(size=w.shape)
        b = b + perturb * rs.normal(size=b.shape)
        new[pre + "weight_v"], new[pre + "weight_g"], new[pre + "bias"] = _t(w), _t(g), _t(b)
    for k in sd:
        if k in new or not (k.startswith("radiance_net.") or k.startswith("nerf_outside.")):
            continue
        if k.endswith("weight_v") or k.endswith(".weight"):
            out_d, in_d = sd[k].shape
            bound = 1.0 / math.sqrt(in_d)
            w = rs.uniform(-bound, bound, size=(out_d, in_d))
            new[k] = _t(w)
            base = k[: -len("weight_v")] if k.endswith("weight_v") else k[: -len("weight")]
            if k.endswith("weight_v"):
                new[base + "weight_g"] = _t(np.linalg.norm(w, axis=1, keepdims=True))
            new[base + "bias"] = _t(rs.uniform(-bound, bound, size=(out_d,)))
    merged = {k: (new[k].to(v.device) if k in new else v) for k, v in sd.items()}
    model.load_state_dict(merged)
    return model


def make_rays(n_rays, shell_radius=2.5, jitter=0.1, seed=0):
    """Origins uniform on a shell of radius ``shell_radius`` looking at the origin with a
    direction jitter; directions are NOT normalised (volume_render normalises, neus.py:172)."""
    rs = np.random.RandomState(seed + 7919)
    o = rs.normal(size=(n_rays, 3))
    o = o / np.linalg.norm(o, axis=1, keepdims=True) * shell_radius
    d = -o / shell_radius + jitter * rs.normal(size=(n_rays, 3))
    d = d * (0.5 + rs.uniform(size=(n_rays, 1)))  # un-normalised on purpose
    return _t(o), _t(d)


def make_points(n, extent=1.0, seed=0):
    rs = np.random.RandomState(seed + 104729)
    return _t(rs.uniform(-extent, extent, size=(n, 3)))


def look_at_pose(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """4x4 camera-to-world (OpenCV convention: +z forward, +y down) looking from `eye` at `target`."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    z = target - eye
    z /= np.linalg.norm(z)
    x = np.cross(z, up)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return _t(m)


def pinhole_intrinsics(H, W, fov_deg=50.0, skew=0.0):
    f = 0.5 * W / math.tan(0.5 * math.radians(fov_deg))
    K = np.eye(4)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1] = f, f * 1.01, 0.5 * W - 0.3, 0.5 * H + 0.2, skew
    return _t(K)


def make_view(seed, eye, H, W):
    """One synthetic training view (the layout the reference's trainers consume): ``model_input`` = {c2w [1,4,4],
    intrinsics [1,4,4], object_mask [1,H*W] bool}, ``ground_truth`` = {rgb [1,H*W,3]}."""
    rs = np.random.RandomState(seed)
    c2w = look_at_pose(eye)[None]
    intr = pinhole_intrinsics(H, W)[None]
    rgb = torch.from_numpy(rs.uniform(size=(1, H * W, 3)).astype(np.float32))
    mask = torch.from_numpy(rs.uniform(size=(1, H * W)) < 0.7)
    return dict(c2w=c2w, intrinsics=intr, object_mask=mask), dict(rgb=rgb)
