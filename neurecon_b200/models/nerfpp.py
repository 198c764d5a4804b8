"""NeRF++ background network forward (models/base.py:426-453) -- built with the VolSDF milestone."""


def nerf_forward(module, input_pts, input_views):
    raise NotImplementedError("neurecon_b200: the NeRF++ background MLP kernels are not built yet")
