"""Golden TRAINING steps: the UNMODIFIED reference's ``Trainer.forward`` + ``backward()`` (neus.py:408-478,
volsdf.py:562-634, unisurf.py:293-346) on a synthetic camera, run on the CPU in the build container:

    python tests/golden/make_golden_grads.py

Per framework one .npz with the step's inputs that a seed cannot reproduce on another device (the sample depths the
reference's samplers chose, its uniform eikonal points / surface jitter), the losses, and the gradient of EVERY parameter
(flattened in ``named_parameters`` order).  tests/test_gpu_train_golden.py replays the step through this package's
Trainer with those depths forced (``samples_bypass``) and compares loss and gradients at north_star's tolerances.

The sample depths are the reference's own: the result of its last ``torch.sort`` over the merged depths (neus.py:275,
unisurf.py:203; VolSDF returns them as ``d_vals``), captured by wrapping ``torch.sort`` during the call and verified
here, bit for bit, against the points the reference fed to its network (captured by wrapping ``forward_with_nablas``).
(The oracle's restated samplers land on the same composited image, but an ulp of difference in an sdf value flips
inverse-CDF bins in flat stretches of the CDF, so its depths are not the reference's on every ray.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from neurecon_b200.utils import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H, W = 24, 32


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def scene(seed, eye, n_rays):
    return synthetic.make_view(seed, eye, H, W)


def capture_nablas_inputs(surface):
    """wrap ImplicitSurface.forward_with_nablas: record the points of every call"""
    calls = []
    orig = surface.forward_with_nablas

    def wrapped(x, *a, **k):
        calls.append(x.detach().clone())
        return orig(x, *a, **k)
    surface.forward_with_nablas = wrapped
    return calls


class SortLog:
    """record the sorted values of every torch.sort call"""

    def __enter__(self):
        self.values, self._real = [], torch.sort

        def logged(*a, **k):
            out = self._real(*a, **k)
            self.values.append(out[0].detach().clone())
            return out
        torch.sort = logged
        return self

    def __exit__(self, *exc):
        torch.sort = self._real


class Fp64Replay:
    """Run the reference in float64 on the SAME step: default dtype double, ``Tensor.float()`` (neus.py:169-170) keeps
    double, the final depth sort returns the float32 run's depths, the device draws return the float32 run's numbers.
    Its gradients are the exact ones up to 1e-15; ``noise`` = how far the reference's own float32 gradients are from them."""

    def __init__(self, width, d_all, uniform=None, rand=None):
        self.width, self.d_all, self.uniform, self.rand = width, d_all.double(), uniform, rand

    def __enter__(self):
        self._sort, self._float, self._uniform, self._rand = torch.sort, torch.Tensor.float, torch.Tensor.uniform_, torch.rand
        torch.set_default_dtype(torch.float64)
        real_float = self._float
        torch.Tensor.float = lambda t: t.double() if t.is_floating_point() else real_float(t).double()
        me = self

        def sort(x, *a, **k):
            out = me._sort(x, *a, **k)
            if x.shape[-1] == me.width:
                return me.d_all.reshape(out[0].shape), out[1]
            return out

        def uniform_(t, *a, **k):
            if me.uniform is not None and t.numel() == me.uniform.numel():
                return t.copy_(me.uniform.reshape(t.shape).to(t.dtype))
            return me._uniform(t, *a, **k)

        def rand(*a, **k):
            out = me._rand(*a, **k)
            if me.rand is not None and out.numel() == me.rand.numel():
                return me.rand.reshape(out.shape).to(out.dtype)
            return out
        torch.sort, torch.Tensor.uniform_, torch.rand = sort, uniform_, rand
        return self

    def __exit__(self, *exc):
        torch.sort, torch.Tensor.float, torch.Tensor.uniform_, torch.rand = self._sort, self._float, self._uniform, self._rand
        torch.set_default_dtype(torch.float32)


def noise_floor(model32, model64):
    """per parameter tensor: max|g32 - g64| / max|g64|"""
    out = []
    for (n, p32), (_, p64) in zip(model32.named_parameters(), model64.named_parameters()):
        out.append(((p32.grad.double() - p64.grad).abs().max() / p64.grad.abs().max().clamp_min(1e-300)).item())
    return np.array(out)


def to64(d):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in d.items()}


def flat_grads(model):
    names, chunks = [], []
    for n, p in model.named_parameters():
        assert p.grad is not None, n
        names.append(n)
        chunks.append(p.grad.detach().reshape(-1))
    return names, torch.cat(chunks)


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def rays_of(ref, model_input, n_rays, seed):
    torch.manual_seed(seed)
    return ref.rend_util.get_rays(model_input["c2w"], model_input["intrinsics"], H, W, N_rays=n_rays)


def golden_neus(ref):
    R, seed = 64, 21
    torch.manual_seed(0)
    m = ref.neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=1)
    model_input, gt = scene(31, [2.0, 1.0, 1.1], R)
    args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_eikonal=0.1, with_mask=True, w_mask=1.0))
    kw = dict(H=H, W=W, batched=True, perturb=False, obj_bounding_radius=1.0, N_outside=0, white_bkgd=False,
              upsample_algo="official_solution", N_nograd_samples=2048, N_upsample_iters=4)
    calls = capture_nablas_inputs(m.implicit_surface)
    trainer = ref.neus.Trainer(m, device_ids=["cpu"], batched=True)
    torch.manual_seed(seed)
    with SortLog() as sorts:
        out = trainer.forward(args, None, model_input, gt, kw, 0, device="cpu")
    out["losses"]["total"].backward()
    names, grads = flat_grads(m)
    # the depths: the last sort of the up-sampler (neus.py:275), verified against the points of neus.py:294
    rays_o, rays_d, sel = rays_of(ref, model_input, R, seed)
    assert torch.equal(sel, out["extras"]["select_inds"])
    d_all = [v for v in sorts.values if v.shape[-1] == 128][-1][0]
    dirs = torch.nn.functional.normalize(rays_d[0], dim=-1)
    pts = rays_o[0][:, None, :] + dirs[:, None, :] * d_all[..., None]
    assert torch.equal(pts.reshape(-1, 3), calls[0].reshape(-1, 3)), "captured depths are not the ones the reference rendered"
    torch.manual_seed(0)
    m64 = ref.neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m64, seed=1)
    m64 = m64.double()
    with Fp64Replay(128, d_all):
        torch.manual_seed(seed)
        o64 = ref.neus.Trainer(m64, device_ids=["cpu"], batched=True).forward(args, None, to64(model_input), to64(gt), kw, 0, device="cpu")
        o64["losses"]["total"].backward()
    noise = noise_floor(m, m64)
    print("neus: loss fp32 %.9g fp64 %.9g, worst fp32-vs-fp64 gradient noise %.3g" % (out["losses"]["total"].item(), o64["losses"]["total"].item(), noise.max()))
    save("train_neus_r64.npz", noise=noise, total64=o64["losses"]["total"], seed=seed, n_rays=R, eye=[2.0, 1.0, 1.1], scene_seed=31, select_inds=sel, rays_o=rays_o,
         rays_d=rays_d, d_all=d_all, rgb=out["extras"]["rgb"], mask_volume=out["extras"]["mask_volume"],
         loss_img=out["losses"]["loss_img"], loss_eikonal=out["losses"]["loss_eikonal"], loss_mask=out["losses"]["loss_mask"],
         total=out["losses"]["total"], grad=grads, names=np.array(names))


def golden_volsdf(ref):
    R, seed = 48, 22
    kwm = dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=0.02)
    torch.manual_seed(0)
    m = ref.volsdf.VolSDF(**kwm)
    synthetic.reseed_parameters(m, seed=3)
    model_input, gt = scene(32, [1.8, 1.5, 1.4], R)
    args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_eikonal=0.1), model=AttrDict(obj_bounding_radius=3.0))
    kw = dict(H=H, W=W, near=0.0, far=6.0, batched=True, perturb=False, white_bkgd=False, max_upsample_steps=6,
              use_nerfplusplus=False, obj_bounding_radius=3.0)
    calls = capture_nablas_inputs(m.implicit_surface)
    trainer = ref.volsdf.Trainer(m, device_ids=["cpu"], batched=True)
    torch.manual_seed(seed)
    out = trainer.forward(args, None, model_input, gt, kw, 0)
    out["losses"]["total"].backward()
    names, grads = flat_grads(m)
    rays_o, rays_d, sel = rays_of(ref, model_input, R, seed)
    assert torch.equal(sel, out["extras"]["select_inds"])
    ex = out["extras"]
    d_all = ex["d_vals"][0]                                    # the reference returns its depths (volsdf.py:523)
    dirs = torch.nn.functional.normalize(rays_d[0], dim=-1)
    pts = rays_o[0][:, None, :] + dirs[:, None, :] * d_all[..., None]
    assert torch.equal(pts.reshape(-1, 3), calls[0].reshape(-1, 3)) and calls[1].shape == (1, R, 1, 3)
    torch.manual_seed(0)
    m64 = ref.volsdf.VolSDF(**kwm)
    synthetic.reseed_parameters(m64, seed=3)
    m64 = m64.double()
    with Fp64Replay(192, d_all, uniform=calls[1]):
        torch.manual_seed(seed)
        o64 = ref.volsdf.Trainer(m64, device_ids=["cpu"], batched=True).forward(args, None, to64(model_input), to64(gt), kw, 0)
        o64["losses"]["total"].backward()
    noise = noise_floor(m, m64)
    print("volsdf: loss fp32 %.9g fp64 %.9g, worst fp32-vs-fp64 gradient noise %.3g" % (out["losses"]["total"].item(), o64["losses"]["total"].item(), noise.max()))
    save("train_volsdf_r48.npz", noise=noise, total64=o64["losses"]["total"], seed=seed, n_rays=R, eye=[1.8, 1.5, 1.4], scene_seed=32, select_inds=sel, rays_o=rays_o,
         rays_d=rays_d, d_all=d_all, beta_map=ex["beta_map"][0], iter_usage=ex["iter_usage"][0], eikonal_points=calls[1],
         rgb=ex["rgb"], loss_img=out["losses"]["loss_img"], loss_eikonal=out["losses"]["loss_eikonal"],
         total=out["losses"]["total"], grad=grads, names=np.array(names))


def golden_unisurf(ref):
    R, seed = 48, 23
    torch.manual_seed(0)
    m = ref.unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=4)
    model_input, gt = scene(33, [2.2, -1.6, 1.2], R)
    args = AttrDict(data=AttrDict(N_rays=R), training=AttrDict(w_reg=0.01, perturb_surface_pts=0.01, delta_max=1.0,
                                                                 delta_min=0.05, delta_beta=1.5e-5))
    kw = dict(H=H, W=W, batched=True, tau=0.5, perturb=False, white_bkgd=False,
              logit_tau=m.get_surface_from_opacity(0.5), radius_of_interest=4.0)
    calls = capture_nablas_inputs(m.implicit_surface)
    draws = []
    real_rand = torch.rand

    def rand_logged(*a, **k):
        t = real_rand(*a, **k)
        draws.append(t.clone())
        return t
    trainer = ref.unisurf.Trainer(m, device_ids=["cpu"], batched=True)
    torch.manual_seed(seed)
    torch.rand = rand_logged
    try:
        with SortLog() as sorts:
            out = trainer.forward(args, None, model_input, gt, kw, 0, device="cpu")
    finally:
        torch.rand = real_rand
    out["losses"]["total"].backward()
    names, grads = flat_grads(m)
    rays_o, rays_d, sel = rays_of(ref, model_input, R, seed)
    assert len(draws) == 1 and draws[0].shape == (1, R, 3)
    d_all = [v for v in sorts.values if v.shape[-1] == 96][-1][0]       # unisurf.py:203
    dirs = torch.nn.functional.normalize(rays_d[0], dim=-1)
    pts = rays_o[0][:, None, :] + dirs[:, None, :] * d_all[..., None]
    # UNISURF.forward sees the flattened [1, R*M, 3] chunk (batchify_query, unisurf.py:214)
    assert torch.equal(pts.reshape(-1, 3), calls[0].reshape(-1, 3)), "captured depths are not the ones the reference rendered"
    ex = out["extras"]
    torch.manual_seed(0)
    m64 = ref.unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
    synthetic.reseed_parameters(m64, seed=4)
    m64 = m64.double()
    with Fp64Replay(96, d_all, rand=draws[0]):
        torch.manual_seed(seed)
        o64 = ref.unisurf.Trainer(m64, device_ids=["cpu"], batched=True).forward(args, None, to64(model_input), to64(gt), kw, 0, device="cpu")
        o64["losses"]["total"].backward()
    noise = noise_floor(m, m64)
    print("unisurf: loss fp32 %.9g fp64 %.9g, worst fp32-vs-fp64 gradient noise %.3g" % (out["losses"]["total"].item(), o64["losses"]["total"].item(), noise.max()))
    save("train_unisurf_r48.npz", noise=noise, total64=o64["losses"]["total"], seed=seed, n_rays=R, eye=[2.2, -1.6, 1.2], scene_seed=33, select_inds=sel, rays_o=rays_o,
         rays_d=rays_d, d_all=d_all, surface_jitter=draws[0], surface_points=ex["surface_points"], mask_surface=ex["mask_surface"],
         rgb=ex["rgb"], loss_img=out["losses"]["loss_img"], loss_reg=out["losses"]["loss_reg"], total=out["losses"]["total"],
         grad=grads, names=np.array(names))


if __name__ == "__main__":
    ref = ref_loader.load()
    which = sys.argv[1:] or ["neus", "volsdf", "unisurf"]
    for w in which:
        {"neus": golden_neus, "volsdf": golden_volsdf, "unisurf": golden_unisurf}[w](ref)
