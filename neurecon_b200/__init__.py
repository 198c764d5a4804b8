"""neurecon_b200 -- B200-native drop-in for neurecon's ray-marched SDF volume-rendering path.

Mirrors the reference's module layout for the hot path only:

    neurecon_b200.models.base                 ImplicitSurface, RadianceNet, NeRF, Embedder
    neurecon_b200.models.frameworks.neus      NeuS, volume_render, SingleRenderer
    neurecon_b200.models.frameworks.volsdf    VolSDF, volume_render, ...
    neurecon_b200.models.frameworks.unisurf   UNISURF, volume_render, ...
    neurecon_b200.models.ray_casting          root_finding_surface_points
    neurecon_b200.utils.rend_util             sample_pdf, sample_cdf, near_far_from_sphere, ...
    neurecon_b200.utils.train_util            batchify_query

Everything numerical runs in hand-written sm_100a CUDA behind the C-ABI library declared in
include/neurecon_b200.h.  There is no CPU fallback: calling the hot path with CPU tensors, or
without the built library, raises.
"""
from ._lib import get_lib, library_path, set_precision, get_precision  # noqa: F401

__version__ = "0.1.0"
