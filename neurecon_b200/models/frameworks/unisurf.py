"""Drop-in for the reference's ``models/frameworks/unisurf.py`` hot path: the ``UNISURF`` module
(:16-62), ``volume_render`` (:64-283) and ``SingleRenderer`` (:285-291)."""
from collections import OrderedDict
from typing import Union

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import _lib
from ..base import ImplicitSurface, RadianceNet, query_radiance


class UNISURF(nn.Module):
    """unisurf.py:16-62 (same constructor, parameters and methods)."""

    def __init__(self, input_ch=3, W_geo_feat=-1, surface_cfg=dict(), radiance_cfg=dict()):
        super().__init__()
        self.implicit_surface = ImplicitSurface(input_ch=input_ch, W_geo_feat=W_geo_feat, **surface_cfg)
        if W_geo_feat < 0:
            W_geo_feat = self.implicit_surface.W
        self.radiance_net = RadianceNet(W_geo_feat=W_geo_feat, **radiance_cfg)

    def forward(self, x, view_dirs):
        occ, nablas, geometry_feature = self.implicit_surface.forward_with_nablas(x)
        # NOTE: like the reference (unisurf.py:36) F.normalize has no dim => dim=1, the point axis of a
        # batched [B, points, 3] chunk.
        normals = F.normalize(nablas)
        radiances = self.radiance_net.forward(x, view_dirs, normals, geometry_feature)
        return radiances, occ, nablas

    @staticmethod
    def get_surface_from_opacity(opacity: Union[torch.Tensor, np.ndarray], eps=1e-4):
        if isinstance(opacity, torch.Tensor):
            opacity = torch.clamp(opacity, min=eps, max=1 - eps)
            return -1. * torch.log(opacity / (1 - opacity))
        opacity = np.clip(opacity, a_min=eps, a_max=1 - eps)
        return -1. * np.log(opacity / (1 - opacity))

    @staticmethod
    def get_opacity_from_surface(imp_surface: Union[torch.Tensor, np.ndarray]):
        if isinstance(imp_surface, torch.Tensor):
            odds = torch.exp(-1. * imp_surface)
        else:
            odds = np.exp(-1. * imp_surface)
        return odds / (1 + odds)
