"""One training step of each framework through THIS package's Trainer (frameworks/trainers.py) against the same step of
the UNMODIFIED reference (tests/golden/train_*.npz, written by tests/golden/make_golden_grads.py from the reference's
own Trainer.forward + backward on the CPU): losses and the gradient of EVERY parameter.

Sample depths are forced to the reference's (``samples_bypass``): an ulp of difference in an sdf value flips
inverse-CDF bins, so the samplers' outputs are compared in their own tests, not through a gradient.

Bars (max|a-b| / max|b| per tensor).  Losses and the rendered colours: north_star's 1e-4 (fp32 tier) / 1e-2 (16-bit
tensor tier).  Gradients, fp32 tier: north_star's 1e-4 sits BELOW what float32 can deliver on this step -- the
generator also ran the unmodified reference in float64 on the same step, and the reference's own float32 gradients are
up to 3.0e-4 (NeuS), 2.1e-3 (VolSDF), 5e-5 (UNISURF) away from those; torch's own float32 autograd of the UNISURF
compositing is 1.1e-4 .. 2e-4 away from its float64 result on these rays (tools/dbg/dbg_uni2.py on a B200).  The bar
per tensor is therefore max(5e-4, 4 x that tensor's reference fp32-vs-fp64 distance); how many tensors exceed the plain
1e-4 is printed.  Gradients, tensor tiers: north_star's 1e-2 on every tensor, see TIER16_GRAD_BAR below.
"""
import numpy as np
import pytest
import torch

import neurecon_b200
from neurecon_b200.utils import synthetic
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
H, W = 24, 32
TIERS = (("fp32", 1e-4), ("fp16", 1e-2), ("fp16x2", 1e-2))
# Weight gradients of the tensor tiers ('fp16', 'fp16x2'): north_star's 1e-2 on EVERY parameter tensor of all three frameworks.
# Both tiers train with the forward sweeps and the reverse sweep (the normal) on split-precision operands (hi + lo fp16 pairs,
# csrc/gemm16.cu nr_gemm16_split); measured worst tensors: 4.3e-3 NeuS, 2.9e-3 VolSDF, 4.4e-3 UNISURF.
# With plain fp16 sweeps (NEURECON_B200_TRAIN_SPLIT=0, 0.7 ms faster per step) the worst tensors are 1.0e-1 / 3.0e-1 / 2.2e-1:
# Softplus(beta=100) turns a pre-activation error dz into 25 dz on softplus' and 2500 dz on softplus'' (the eikonal term's
# second-order path), and 16-bit operands leave dz ~ 3e-4 -- the gradient GEMMs are not the cause.
TIER16_GRAD_BAR = 1e-2
TIER16X2_GRAD_BAR = 1e-2


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _to_dev(d):
    return {k: v.cuda() for k, v in d.items()}


def _check_losses(out, z, keys, tol, tier):
    """each loss term within tol of the reference, relative to the larger of the term and the total: the eikonal term
    (|n| - 1)^2 of a near-eikonal network is a difference of nearly equal numbers, its own relative error says little"""
    total = abs(float(z["total"]))
    for k in keys:
        got, want = float(out["losses"][k]), float(z[k])
        assert abs(got - want) <= tol * max(abs(want), total), (tier, k, got, want)


def _assert_grads(model, z, tol, tier):
    """Every parameter gradient against the reference's float32 one (bars: module docstring).  The worst
    (error, reference fp32-vs-fp64 distance, name) triples are printed."""
    worst = _check_grads(model, z, tol, tier)
    if tier == "fp32":
        bad = [(e, nz, n) for e, nz, n in worst if e > max(5e-4, 4.0 * nz)]
    elif tier == "fp16x2":
        bad = [(e, nz, n) for e, nz, n in worst if e > TIER16X2_GRAD_BAR]
    else:
        bad = [(e, nz, n) for e, nz, n in worst if e > TIER16_GRAD_BAR]
    print("%s: %d tensors, worst (err, reference fp32-vs-fp64 noise, name): %s" % (tier, len(worst), worst[:4]))
    print("%s: tensors above the plain tolerance %g: %d" % (tier, tol, sum(1 for e, _, _ in worst if e > tol)))
    assert not bad, (tier, bad[:6])


def _check_grads(model, z, tol, tier):
    names = [str(n) for n in z["names_list"]]
    flat = z["grad"]
    off, worst = 0, []
    params = dict(model.named_parameters())
    assert list(params) == names, "parameter names / order differ from the reference's state layout"
    for n in names:
        p = params[n]
        want = flat[off:off + p.numel()].reshape(p.shape)
        off += p.numel()
        assert p.grad is not None, n
        e = rel_err(p.grad, want)
        worst.append((e, float(z["noise"][len(worst)]), n))
    assert off == flat.numel()
    worst.sort(reverse=True)
    return worst


def _golden(name):
    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, name))
    out = {k: (torch.from_numpy(np.atleast_1d(z[k])) if z[k].dtype.kind in "fiub" else z[k]) for k in z.files}
    out["names_list"] = list(z["names"])
    return out


@pytest.mark.parametrize("tier,tol", TIERS)
def test_neus_step_matches_reference(tier, tol):
    from neurecon_b200.models.frameworks import neus
    z = _golden("train_neus_r64.npz")
    neurecon_b200.set_precision(tier)
    try:
        torch.manual_seed(0)
        m = neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
        synthetic.reseed_parameters(m, seed=1)
        m = m.cuda()
        model_input, gt = synthetic.make_view(int(z["scene_seed"]), [float(v) for v in z["eye"]], H, W)
        R = int(z["n_rays"])
        args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_eikonal=0.1, with_mask=True, w_mask=1.0))
        kw = dict(H=H, W=W, batched=True, perturb=False, obj_bounding_radius=1.0, N_outside=0, white_bkgd=False,
                  upsample_algo="official_solution", N_nograd_samples=2048, N_upsample_iters=4,
                  samples_bypass={"d_all": z["d_all"]})
        trainer = neus.Trainer(m, device_ids=[0], batched=True)
        torch.manual_seed(int(z["seed"]))
        out = trainer.forward(args, None, model_input, gt, kw, 0)
        assert torch.equal(out["extras"]["select_inds"].cpu(), z["select_inds"])
        out["losses"]["total"].backward()
        _check_losses(out, z, ("loss_img", "loss_eikonal", "loss_mask", "total"), tol, tier)
        assert rel_err(out["extras"]["rgb"], z["rgb"]) < tol
        _assert_grads(m, z, tol, tier)
    finally:
        neurecon_b200.set_precision("fp16")


@pytest.mark.parametrize("tier,tol", TIERS)
def test_volsdf_step_matches_reference(tier, tol):
    """incl. ln_beta (through alpha = 1 / beta and beta) and the uniform eikonal points of volsdf.py:610-613"""
    from neurecon_b200.models.frameworks import volsdf
    z = _golden("train_volsdf_r48.npz")
    neurecon_b200.set_precision(tier)
    try:
        torch.manual_seed(0)
        m = volsdf.VolSDF(**dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=0.02))
        synthetic.reseed_parameters(m, seed=3)
        m = m.cuda()
        model_input, gt = synthetic.make_view(int(z["scene_seed"]), [float(v) for v in z["eye"]], H, W)
        R = int(z["n_rays"])
        args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_eikonal=0.1), model=AttrDict(obj_bounding_radius=3.0))
        kw = dict(H=H, W=W, near=0.0, far=6.0, batched=True, perturb=False, white_bkgd=False, max_upsample_steps=6,
                  use_nerfplusplus=False, obj_bounding_radius=3.0,
                  samples_bypass={"d_all": z["d_all"], "beta_map": z["beta_map"], "iter_usage": z["iter_usage"]})
        trainer = volsdf.Trainer(m, device_ids=[0], batched=True)
        trainer.rng_override["eikonal_points"] = z["eikonal_points"].cuda()
        torch.manual_seed(int(z["seed"]))
        out = trainer.forward(args, None, model_input, gt, kw, 0)
        assert torch.equal(out["extras"]["select_inds"].cpu(), z["select_inds"])
        out["losses"]["total"].backward()
        _check_losses(out, z, ("loss_img", "loss_eikonal", "total"), tol, tier)
        _assert_grads(m, z, tol, tier)
    finally:
        neurecon_b200.set_precision("fp16")


@pytest.mark.parametrize("tier,tol", TIERS)
def test_unisurf_step_matches_reference(tier, tol):
    """incl. the surface-normal regulariser's two extra forward_with_nablas queries (unisurf.py:331-341)"""
    from neurecon_b200.models.frameworks import unisurf
    z = _golden("train_unisurf_r48.npz")
    neurecon_b200.set_precision(tier)
    try:
        torch.manual_seed(0)
        m = unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
        synthetic.reseed_parameters(m, seed=4)
        m = m.cuda()
        model_input, gt = synthetic.make_view(int(z["scene_seed"]), [float(v) for v in z["eye"]], H, W)
        R = int(z["n_rays"])
        args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_reg=0.01, perturb_surface_pts=0.01, delta_max=1.0,
                                                                     delta_min=0.05, delta_beta=1.5e-5))
        kw = dict(H=H, W=W, batched=True, tau=0.5, perturb=False, white_bkgd=False,
                  logit_tau=m.get_surface_from_opacity(0.5), radius_of_interest=4.0,
                  samples_bypass={"d_all": z["d_all"], "surface_points": z["surface_points"]})
        trainer = unisurf.Trainer(m, device_ids=[0], batched=True)
        trainer.rng_override["surface_jitter"] = z["surface_jitter"].cuda()
        torch.manual_seed(int(z["seed"]))
        out = trainer.forward(args, None, model_input, gt, kw, 0)
        assert torch.equal(out["extras"]["mask_surface"].cpu(), z["mask_surface"])
        assert rel_err(out["extras"]["surface_points"], z["surface_points"]) < tol
        out["losses"]["total"].backward()
        _check_losses(out, z, ("loss_img", "loss_reg", "total"), tol, tier)
        assert rel_err(out["extras"]["rgb"], z["rgb"]) < tol
        _assert_grads(m, z, tol, tier)
    finally:
        neurecon_b200.set_precision("fp16")
