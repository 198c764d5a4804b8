"""Drop-in for the hot-path part of the reference's ``utils/train_util.py``."""
import torch


def batchify_query(query_fn, *args, chunk, dim_batchify):
    """utils/train_util.py:23-71: flatten ``[(B), N_rays, N_pts, ...]`` to points, call
    ``query_fn`` on slices of ``chunk`` points and restore the ``[(B), N_rays, N_pts, ...]``
    shape.  Kept for callers that chunk explicitly; the fused kernels tile internally, so
    the neurecon_b200 renderers do not need it to bound memory."""
    n_rays = args[0].shape[dim_batchify]
    n_pts = args[0].shape[dim_batchify + 1]
    flat = [a.flatten(dim_batchify, dim_batchify + 1) for a in args]
    total = flat[0].shape[dim_batchify]
    if dim_batchify not in (0, 1, 2):
        raise NotImplementedError
    pieces = []
    for i in range(0, total, chunk):
        sl = [a.narrow(dim_batchify, i, min(chunk, total - i)) for a in flat]
        r = query_fn(*sl)
        pieces.append(r if isinstance(r, tuple) else [r])

    def restore(v):
        return v.reshape([*v.shape[:dim_batchify], n_rays, n_pts, *v.shape[dim_batchify + 1:]])

    out = []
    for entry in zip(*pieces):
        if isinstance(entry[0], dict):
            merged = {}
            for d in entry:
                for k, v in d.items():
                    merged.setdefault(k, []).append(v)
            out.append({k: restore(torch.cat(v, dim=dim_batchify)) for k, v in merged.items()})
        else:
            out.append(restore(torch.cat(entry, dim=dim_batchify)))
    return out[0] if len(out) == 1 else tuple(out)


# ------------------------------------------------------------------------------------------------------------
# the step after the path in training -- losses, gradient norm, Adam -- on the device.
# ------------------------------------------------------------------------------------------------------------
class _NeusLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, nablas, mask_volume, target_rgb, target_mask, mask_ignore, w_eikonal, w_mask):
        from .. import _lib
        _lib.require_cuda(rgb, nablas)
        lib = _lib.get_lib()
        dev = rgb.device
        rgb_c, nab_c, tgt_c = _lib.f32c(rgb.detach().reshape(-1, 3)), _lib.f32c(nablas.detach()), _lib.f32c(target_rgb.reshape(-1, 3))
        R = rgb_c.shape[0]
        P = nab_c.numel() // (3 * R)
        acc_c = _lib.f32c(mask_volume.detach().reshape(-1)) if mask_volume is not None else None
        tm = target_mask.reshape(-1).to(torch.uint8).contiguous() if target_mask is not None else None
        mi = mask_ignore.reshape(-1).to(torch.uint8).contiguous() if mask_ignore is not None else None
        sums = torch.empty(4, dtype=torch.float32, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        g_rgb, g_nab = torch.empty_like(rgb_c), torch.empty_like(nab_c)
        g_acc = torch.empty_like(acc_c) if tm is not None else None
        with torch.cuda.device(dev):
            _lib.check(lib.nr_neus_loss(_lib.ptr(rgb_c), _lib.ptr(tgt_c), _lib.ptr(nab_c), _lib.ptr(acc_c), _lib.ptr(tm),
                                        _lib.ptr(mi), R, P, float(w_eikonal), float(w_mask), _lib.ptr(sums), _lib.ptr(losses),
                                        _lib.ptr(g_rgb), _lib.ptr(g_nab), _lib.ptr(g_acc), _lib.stream_ptr(dev)), "neus_loss")
        ctx.save_for_backward(g_rgb, g_nab, g_acc if g_acc is not None else torch.empty(0, device=dev))
        ctx.shapes = (rgb.shape, nablas.shape, None if mask_volume is None else mask_volume.shape, g_acc is not None)
        ctx.mark_non_differentiable(losses)
        return losses[3].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        g_rgb, g_nab, g_acc = ctx.saved_tensors
        s_rgb, s_nab, s_acc, has_acc = ctx.shapes
        ga = (g_acc * g_total).reshape(s_acc) if (has_acc and s_acc is not None) else None
        return (g_rgb * g_total).reshape(s_rgb), (g_nab * g_total).reshape(s_nab), ga, None, None, None, None, None


def neus_losses(rgb, target_rgb, nablas, mask_volume=None, target_mask=None, mask_ignore=None, w_eikonal=0.1, w_mask=0.0):
    """The losses of the reference's NeuS ``Trainer.forward`` (neus.py:443-478) in two launches, no host sync:
    returns OrderedDict(loss_img, loss_eikonal[, loss_mask], total); ``total`` carries the gradient to ``rgb``,
    ``nablas`` and (with a target mask) ``mask_volume``; the individual terms are detached device scalars."""
    from collections import OrderedDict
    total, parts = _NeusLossFn.apply(rgb, nablas, mask_volume if target_mask is not None else None, target_rgb, target_mask,
                                     mask_ignore, w_eikonal, w_mask)
    out = OrderedDict([("loss_img", parts[0]), ("loss_eikonal", parts[1])])
    if target_mask is not None:
        out["loss_mask"] = parts[2]
    out["total"] = total
    return out


def _tensor_table(entries, device):
    """Device table of {param*, grad*, exp_avg*, exp_avg_sq*, numel} records for the multi-tensor kernels."""
    rows = []
    for p, g, m, v in entries:
        rows.append([p.data_ptr(), g.data_ptr(), 0 if m is None else m.data_ptr(), 0 if v is None else v.data_ptr(), p.numel()])
    return torch.tensor(rows, dtype=torch.int64).to(device, non_blocking=True)


def grad_norm_device(*modules):
    """2-norm of all gradients of the modules as a 0-dim device tensor (one launch, no sync)."""
    from .. import _lib
    ps = [p for m in modules for p in m.parameters() if p.requires_grad and p.grad is not None]
    if not ps:
        return torch.zeros(())
    dev = ps[0].device
    for p in ps:
        assert p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.device == dev
    tab = _tensor_table([(p.grad, p.grad, None, None) for p in ps], dev)
    out = torch.empty(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.get_lib().nr_grad_sqsum(_lib.ptr(tab), len(ps), _lib.ptr(out), _lib.stream_ptr(dev)), "grad_sqsum")
    return out.sqrt()[0]


def calc_grad_norm(norm_type=2.0, **named_models):
    """train_util.py:5-16 (same return: dict of Python floats incl. 'total'), one device reduction and one sync per
    model instead of one ``.item()`` per parameter."""
    assert float(norm_type) == 2.0, "neurecon_b200 calc_grad_norm supports the 2-norm"
    norms = {"total": 0.0}
    vals = {name: grad_norm_device(model) for name, model in named_models.items()}
    for name, v in vals.items():
        norms[name] = float(v)
        norms["total"] += norms[name] ** 2
    norms["total"] = norms["total"] ** 0.5
    return norms


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (betas, eps; no weight decay / amsgrad) with one multi-tensor launch per parameter group,
    so the per-module learning-rate groups of the reference's ``get_optimizer`` (base.py:486-521) carry over."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        from .. import _lib
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.get_lib()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            entries = []
            for p in ps:
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                assert p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32
                entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            group["step"] = group.get("step", 0) + 1
            key = tuple(t.data_ptr() for e in entries for t in e)
            if group.get("_table_key") != key:
                group["_table"], group["_table_key"] = _tensor_table(entries, dev), key
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                _lib.check(lib.nr_adam_step(_lib.ptr(group["_table"]), len(entries), float(group["lr"]), float(b1), float(b2),
                                            float(group["eps"]), int(group["step"]), _lib.stream_ptr(dev)), "adam_step")
        return loss
