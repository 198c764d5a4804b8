"""Drop-in for the reference's ``models/frameworks/volsdf.py`` hot path: ``sdf_to_sigma`` (:16-35),
``error_bound`` (:38-74), ``fine_sample`` (:77-272), the ``VolSDF`` module (:274-331),
``volume_render`` (:334-551) and ``SingleRenderer`` (:554-560)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from ..base import ImplicitSurface, RadianceNet, query_radiance


class VolSDF(nn.Module):
    """volsdf.py:274-331 (same constructor, parameters and methods)."""

    def __init__(self, beta_init=0.1, speed_factor=1.0, input_ch=3, W_geo_feat=-1, obj_bounding_radius=3.0,
                 use_nerfplusplus=False, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        self.speed_factor = speed_factor
        self.ln_beta = nn.Parameter(data=torch.Tensor([np.log(beta_init) / self.speed_factor]), requires_grad=True)
        self.use_sphere_bg = not use_nerfplusplus
        self.obj_bounding_radius = obj_bounding_radius
        self.implicit_surface = ImplicitSurface(
            W_geo_feat=W_geo_feat, input_ch=input_ch, obj_bounding_size=obj_bounding_radius, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)
        if use_nerfplusplus:
            from ..base import NeRF
            self.nerf_outside = NeRF(input_ch=4, multires=10, multires_view=4, use_view_dirs=True)

    def forward_ab(self):
        beta = torch.exp(self.ln_beta * self.speed_factor)
        return 1. / beta, beta

    def forward_surface(self, x):
        sdf = self.implicit_surface.forward(x)
        if self.use_sphere_bg:
            return torch.min(sdf, self.obj_bounding_radius - x.norm(dim=-1))
        return sdf

    def forward_surface_with_nablas(self, x):
        sdf, nablas, h = self.implicit_surface.forward_with_nablas(x)
        if self.use_sphere_bg:
            d_bg = self.obj_bounding_radius - x.norm(dim=-1)
            sdf = torch.where(d_bg < sdf, d_bg, sdf)
        return sdf, nablas, h

    def forward(self, x, view_dirs):
        sdf, nablas, geometry_feature = self.forward_surface_with_nablas(x)
        radiances = self.radiance_net.forward(x, view_dirs, nablas, geometry_feature)
        return radiances, sdf, nablas
