"""``Trainer`` and ``get_model`` of the three frameworks (reference: neus.py:408-546, volsdf.py:562-736,
unisurf.py:293-401) -- the callers of the hot path in training.  They stay PyTorch modules with the reference's
constructor / ``forward(args, indices, model_input, ground_truth, render_kwargs_train, it)`` signatures and return
structure (``OrderedDict(losses=..., extras=...)``), so ``train.py`` of the reference drives them unchanged; what they
call underneath is this package: ``rend_util.get_rays`` (device kernel), ``volume_render`` under autograd (fused MLP
Functions + compositing adjoints), the extra ``forward_with_nablas`` queries of SURVEY row a22, and for NeuS the device
loss kernel (``train_util.neus_losses``).

``trainer.rng_override`` is a determinism hook for parity tests: tensors stored under "eikonal_points" (VolSDF,
volsdf.py:610) or "surface_jitter" (UNISURF, unisurf.py:335) replace the device draws of those lines, so that a run can
be compared with gradients the reference produced from its own CPU generator.
"""
import copy
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ...utils import rend_util


class _RayTrainer(nn.Module):
    """What the three trainers share: the renderer (DataParallel over the ray axis for several devices, as
    neus.py:413-414), pixel selection and target gathering."""

    def __init__(self, model, renderer, device_ids=[0], batched=True):
        super().__init__()
        self.model = model
        self.renderer = renderer
        if len(device_ids) > 1:
            self.renderer = nn.DataParallel(self.renderer, device_ids=device_ids, dim=1 if batched else 0)
        self.device = device_ids[0]
        self.rng_override = {}

    def _select(self, args, model_input, ground_truth, render_kwargs_train, device):
        """get_rays for args.data.N_rays random pixels + the targets at those pixels."""
        intr = model_input["intrinsics"].to(device)
        c2w = model_input["c2w"].to(device)
        rays_o, rays_d, select_inds = rend_util.get_rays(c2w, intr, render_kwargs_train["H"], render_kwargs_train["W"],
                                                         N_rays=args.data.N_rays)
        target_rgb = torch.gather(ground_truth["rgb"].to(device), -2, select_inds.unsqueeze(-1).expand(*select_inds.shape, 3))
        ignore = None
        if "mask_ignore" in model_input:
            ignore = torch.gather(model_input["mask_ignore"].to(device), 1, select_inds)
        return rays_o, rays_d, select_inds, target_rgb, ignore

    @staticmethod
    def _image_loss(rgb, target_rgb, weight_mask=None):
        """mean |rgb - target|, or the masked mean sum(|.| m) / (sum(m) + 1e-10)."""
        l1 = (rgb - target_rgb).abs()
        if weight_mask is None:
            return l1.mean()
        return (l1 * weight_mask[..., None].float()).sum() / (weight_mask.sum() + 1e-10)

    @staticmethod
    def _total(losses):
        total = 0
        for v in losses.values():
            total = total + v
        losses["total"] = total
        return losses


class NeusTrainer(_RayTrainer):
    """neus.py:408-478."""

    def __init__(self, model, device_ids=[0], batched=True):
        from .neus import SingleRenderer
        super().__init__(model, SingleRenderer(model), device_ids, batched)

    def forward(self, args, indices, model_input, ground_truth, render_kwargs_train: dict, it: int, device="cuda"):
        from ...utils import train_util
        rays_o, rays_d, select_inds, target_rgb, ignore = self._select(args, model_input, ground_truth, render_kwargs_train, device)
        rgb, depth_v, extras = self.renderer(rays_o, rays_d, detailed_output=True, **render_kwargs_train)
        nablas = extras["implicit_nablas"]
        mask_volume = extras["mask_volume"]
        extras["mask_volume_clipped"] = torch.clamp(mask_volume, 1e-3, 1 - 1e-3)
        target_mask = None
        if args.training.with_mask:
            target_mask = torch.gather(model_input["object_mask"].to(device), 1, select_inds)
        # L1 + eikonal + mask BCE and their gradients in one device kernel (neus.py:453-478), no host read
        P = nablas.shape[-2]
        parts = train_util.neus_losses(
            rgb.reshape(-1, 3), target_rgb.reshape(-1, 3), nablas.reshape(-1, P, 3), mask_volume=mask_volume.reshape(-1),
            target_mask=None if target_mask is None else target_mask.reshape(-1),
            mask_ignore=None if ignore is None else ignore.reshape(-1),
            w_eikonal=args.training.w_eikonal, w_mask=args.training.w_mask if args.training.with_mask else 0.0)
        losses = OrderedDict(parts)
        extras["implicit_nablas_norm"] = torch.norm(nablas.detach(), dim=-1)
        extras["scalars"] = {"1/s": 1. / self.model.forward_s().data}
        extras["select_inds"] = select_inds
        return OrderedDict([("losses", losses), ("extras", extras)])


class VolsdfTrainer(_RayTrainer):
    """volsdf.py:562-634."""

    def __init__(self, model, device_ids=[0], batched=True):
        from .volsdf import SingleRenderer
        super().__init__(model, SingleRenderer(model), device_ids, batched)

    def forward(self, args, indices, model_input, ground_truth, render_kwargs_train: dict, it: int):
        device = self.device
        rays_o, rays_d, select_inds, target_rgb, ignore = self._select(args, model_input, ground_truth, render_kwargs_train, device)
        rgb, depth_v, extras = self.renderer(rays_o, rays_d, detailed_output=True, **render_kwargs_train)
        nablas = extras["implicit_nablas"]
        # eikonal term on ONE render sample per ray (the one of largest weight) and ONE uniform point of the bounding box
        # (volsdf.py:603-613; the extra forward_with_nablas is SURVEY row a22)
        best = extras["visibility_weights"][..., :nablas.shape[-2]].argmax(dim=-1)
        nablas_hit = torch.gather(nablas, -2, best[..., None, None].expand(*best.shape, 1, 3))
        box = args.model.obj_bounding_radius
        pts = self.rng_override.get("eikonal_points")
        if pts is None:
            pts = torch.empty_like(nablas_hit).uniform_(-box, box)
        _, nablas_box, _ = self.model.implicit_surface.forward_with_nablas(pts.to(device))
        nablas_norm = torch.norm(torch.cat([nablas_hit, nablas_box], dim=-2), dim=-1)
        losses = OrderedDict()
        losses["loss_img"] = self._image_loss(rgb, target_rgb, ignore)
        losses["loss_eikonal"] = args.training.w_eikonal * ((nablas_norm - 1.0) ** 2).mean()
        self._total(losses)
        extras["implicit_nablas_norm"] = nablas_norm
        alpha, beta = self.model.forward_ab()
        extras["scalars"] = {"beta": beta.data, "alpha": alpha.data}
        extras["select_inds"] = select_inds
        return OrderedDict([("losses", losses), ("extras", extras)])


class UnisurfTrainer(_RayTrainer):
    """unisurf.py:293-346."""

    def __init__(self, model, device_ids=[0], batched=True):
        from .unisurf import SingleRenderer
        super().__init__(model, SingleRenderer(model), device_ids, batched)

    def forward(self, args, indices, model_input, ground_truth, render_kwargs_train: dict, it: int, device="cuda"):
        rays_o, rays_d, select_inds, target_rgb, _ = self._select(args, model_input, ground_truth, render_kwargs_train, device)
        tr = args.training
        interval = max(tr.delta_max * np.exp(-it * tr.delta_beta), tr.delta_min)          # unisurf.py:322
        rgb, depth_v, extras = self.renderer(rays_o, rays_d, interval=interval, detailed_output=True, **render_kwargs_train)
        losses = OrderedDict()
        losses["loss_img"] = self._image_loss(rgb, target_rgb)
        losses["loss_reg"] = torch.zeros((), device=rgb.device)
        if tr.w_reg > 0:
            # surface-normal smoothness: normals at the surface points vs. at jittered neighbours (unisurf.py:331-341, a22)
            surf = extras["surface_points"]
            jitter = self.rng_override.get("surface_jitter")
            if jitter is None:
                jitter = torch.rand(surf.shape, device=surf.device)
            neighbours = surf + (jitter.to(surf.device) - 0.5) * 2. * tr.perturb_surface_pts
            _, n_surf, _ = self.model.implicit_surface.forward_with_nablas(surf)
            _, n_near, _ = self.model.implicit_surface.forward_with_nablas(neighbours)
            losses["loss_reg"] = tr.w_reg * ((F.normalize(n_near, dim=-1) - F.normalize(n_surf, dim=-1)) ** 2).mean()
        self._total(losses)
        extras["scalars"] = {"interval": torch.tensor([interval]).to(rgb.device)}
        return OrderedDict([("losses", losses), ("extras", extras)])


# ---- get_model: config -> (model, trainer, render_kwargs_train, render_kwargs_test, renderer) ---------------------------
def _net_cfgs(args):
    """surface_cfg / radiance_cfg from args.model.{surface,radiance} with the reference's defaults (written back into
    args with setdefault, as the reference does, so that the saved config is complete)."""
    m = args.model
    siren = m.setdefault("use_siren", False)
    surface = dict(use_siren=m.surface.setdefault("use_siren", siren), embed_multires=m.surface.setdefault("embed_multires", 6),
                   radius_init=m.surface.setdefault("radius_init", 1.0), geometric_init=m.surface.setdefault("geometric_init", True),
                   D=m.surface.setdefault("D", 8), W=m.surface.setdefault("W", 256), skips=m.surface.setdefault("skips", [4]))
    radiance = dict(use_siren=m.radiance.setdefault("use_siren", siren), embed_multires=m.radiance.setdefault("embed_multires", -1),
                    embed_multires_view=m.radiance.setdefault("embed_multires_view", -1),
                    use_view_dirs=m.radiance.setdefault("use_view_dirs", True), D=m.radiance.setdefault("D", 4),
                    W=m.radiance.setdefault("W", 256), skips=m.radiance.setdefault("skips", []))
    return surface, radiance


def _finish(args, model, trainer_cls, render_kwargs_train):
    render_kwargs_test = copy.deepcopy(render_kwargs_train)
    render_kwargs_test["rayschunk"] = args.data.val_rayschunk
    render_kwargs_test["perturb"] = False
    trainer = trainer_cls(model, device_ids=args.device_ids, batched=render_kwargs_train["batched"])
    return model, trainer, render_kwargs_train, render_kwargs_test, trainer.renderer


def get_model_neus(args):
    """neus.py:481-546."""
    from .neus import NeuS
    if not args.training.with_mask:
        assert "N_outside" in args.model.keys() and args.model.N_outside > 0, \
            "Please specify a positive model:N_outside for neus with nerf++"
    surface, radiance = _net_cfgs(args)
    model = NeuS(obj_bounding_radius=args.model.obj_bounding_radius, W_geo_feat=args.model.setdefault("W_geometry_feature", 256),
                 use_outside_nerf=not args.training.with_mask, speed_factor=args.training.setdefault("speed_factor", 1.0),
                 variance_init=args.model.setdefault("variance_init", 0.05), surface_cfg=surface, radiance_cfg=radiance)
    kw = dict(upsample_algo=args.model.setdefault("upsample_algo", "official_solution"),
              N_nograd_samples=args.model.setdefault("N_nograd_samples", 2048),
              N_upsample_iters=args.model.setdefault("N_upsample_iters", 4), N_outside=args.model.setdefault("N_outside", 0),
              obj_bounding_radius=args.data.setdefault("obj_bounding_radius", 1.0), batched=args.data.batch_size is not None,
              perturb=args.model.setdefault("perturb", True), white_bkgd=args.model.setdefault("white_bkgd", False))
    return _finish(args, model, NeusTrainer, kw)


def get_model_volsdf(args):
    """volsdf.py:672-736."""
    from .volsdf import VolSDF
    surface, radiance = _net_cfgs(args)
    model = VolSDF(use_nerfplusplus=args.model.setdefault("outside_scene", "builtin") == "nerf++",
                   obj_bounding_radius=args.model.obj_bounding_radius, W_geo_feat=args.model.setdefault("W_geometry_feature", 256),
                   speed_factor=args.training.setdefault("speed_factor", 1.0), beta_init=args.training.setdefault("beta_init", 0.1),
                   surface_cfg=surface, radiance_cfg=radiance)
    kw = dict(near=args.data.near, far=args.data.far, batched=True, perturb=args.model.setdefault("perturb", True),
              white_bkgd=args.model.setdefault("white_bkgd", False), max_upsample_steps=args.model.setdefault("max_upsample_iter", 5),
              use_nerfplusplus=args.model.outside_scene == "nerf++", obj_bounding_radius=args.model.obj_bounding_radius)
    return _finish(args, model, VolsdfTrainer, kw)


def get_model_unisurf(args):
    """unisurf.py:349-401."""
    from .unisurf import UNISURF
    surface, radiance = _net_cfgs(args)
    model = UNISURF(W_geo_feat=args.model.setdefault("W_geometry_feature", 256), surface_cfg=surface, radiance_cfg=radiance)
    kw = dict(batched=True, tau=args.model.tau, perturb=args.model.get("perturb", True),
              white_bkgd=args.model.get("white_bkgd", False), logit_tau=model.get_surface_from_opacity(args.model.tau),
              radius_of_interest=args.model.obj_bounding_radius)
    return _finish(args, model, UnisurfTrainer, kw)
