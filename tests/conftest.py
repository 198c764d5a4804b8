import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim > 0 else z[k].item() for k in z.files}


def rel_err(a, b):
    """max|a-b| / max|b| over the tensor (SURVEY.md section 7 convention)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def frac_close(a, b, tol):
    """fraction of elements with |a-b| <= tol * max|b| -- for per-sample tensors, whose sample
    positions may legitimately move when an inverse-CDF bin flips by an ulp."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs() <= tol * b.abs().max()).double().mean().item()


NEUS_CFG = dict(multires=6, multires_view=4, rad_multires=-1, skips=[4], D=8, D_rad=4, speed_factor=10.0)


def build_neus(seed=1, device="cpu"):
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import synthetic
    torch.manual_seed(0)
    m = neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=seed)
    return m.to(device)


def build_volsdf(beta_init=0.1, nerfpp=False, seed=3, device="cpu"):
    from neurecon_b200.models.frameworks import volsdf
    from neurecon_b200.utils import synthetic
    torch.manual_seed(0)
    m = volsdf.VolSDF(**dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=beta_init, use_nerfplusplus=nerfpp))
    synthetic.reseed_parameters(m, seed=seed)
    return m.to(device)


def build_unisurf(seed=4, device="cpu"):
    from neurecon_b200.models.frameworks import unisurf
    from neurecon_b200.utils import synthetic
    torch.manual_seed(0)
    m = unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=seed)
    return m.to(device)


def build_neus_bg(seed=5, device="cpu"):
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import synthetic
    torch.manual_seed(0)
    m = neus.NeuS(**dict(synthetic.NEUS_MODEL_KWARGS, use_outside_nerf=True))
    synthetic.reseed_parameters(m, seed=seed)
    return m.to(device)


def cpu_state_dict(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
