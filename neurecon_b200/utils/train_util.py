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


_NORM_TABLES = {}


def grad_norm_device(*modules):
    """2-norm of all gradients of the modules as a 0-dim device tensor (one launch, no sync)."""
    from .. import _lib
    ps = [p for m in modules for p in m.parameters() if p.requires_grad and p.grad is not None]
    if not ps:
        return torch.zeros(())
    dev = ps[0].device
    for p in ps:
        assert p.grad.is_contiguous() and p.grad.dtype == torch.float32 and p.grad.device == dev
    key = tuple((p.grad.data_ptr(), p.numel()) for p in ps)
    # The table is uploaded outside any graph capture (the copy comes from pageable memory) and cached by gradient
    # addresses.  A table a capture has used is pinned for good -- the graph holds its pointer; the others are a small
    # LRU, so that a loop with zero_grad(set_to_none=True) (fresh gradient tensors every step) does not pile tables up.
    capturing = torch.cuda.is_current_stream_capturing()
    entry = _NORM_TABLES.pop(key, None)
    if entry is None:
        if capturing:
            raise RuntimeError("neurecon_b200: grad_norm_device inside a CUDA-graph capture needs a warm-up call on the same "
                               "gradient tensors first (use zero_grad(set_to_none=False))")
        entry = [_tensor_table([(p.grad, p.grad, None, None) for p in ps], dev), False]
    entry[1] = entry[1] or capturing
    _NORM_TABLES[key] = entry                               # most recently used last
    loose = [k for k, e in _NORM_TABLES.items() if not e[1]]
    for k in loose[:-8]:
        del _NORM_TABLES[k]
    tab = entry[0]
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
    so the per-module learning-rate groups of the reference's ``get_optimizer`` (base.py:486-521) carry over.
    ``capturable=True`` keeps the step count and each group's lr in device memory (``nr_adam_step_dev``), which is
    what a CUDA graph of the step needs: `CapturedStep` refreshes the device lr from ``group["lr"]`` (where the
    reference's schedulers write it, train.py:210) before every replay."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        # the group keys torch.optim.Adam reads, with the only values this optimiser implements, so that a state_dict
        # saved here loads into torch.optim.Adam and back (the reference checkpoints optimizer.state_dict());
        # foreach: zero_grad() as one multi-tensor launch
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=True, capturable=bool(capturable), differentiable=False, fused=None,
                                      decoupled_weight_decay=False))
        self.capturable = bool(capturable)
        self._dev = {}          # per group index: pointer table, its key, device lr / step count -- never part of state_dict()

    def push_lr(self):
        """Copy every group's host ``lr`` into its device slot (stream-ordered, no sync).  Not recorded while a CUDA
        graph is being captured: a replay must see the lr of its own iteration, not the capture's."""
        if not self.capturable or torch.cuda.is_current_stream_capturing():
            return
        for gi, group in enumerate(self.param_groups):
            d = self._dev.get(gi)
            if d is not None and "lr" in d:
                d["lr"].fill_(float(group["lr"]))

    def state_dict(self):
        """Checkpointing (checkpoints.py of the reference saves ``optimizer.state_dict()``) in torch.optim.Adam's layout:
        the step count lives per group here (one bias correction per multi-tensor launch; graph replays advance it on
        the device only, so it is read back first -- one sync per group) and is mirrored into ``state[p]["step"]`` as
        the float tensor torch keeps, so that the dict loads into ``torch.optim.Adam`` with warm moments AND the right
        bias correction."""
        for gi, group in enumerate(self.param_groups):
            d = self._dev.get(gi)
            if self.capturable and d is not None and "step" in d:
                group["step"] = int(d["step"].item())
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = torch.tensor(float(group.get("step", 0)), dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Accepts this class's dicts and torch.optim.Adam's: the group's step is taken from the loaded per-parameter
        steps (their maximum; torch keeps them equal for parameters that always have a gradient)."""
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            steps = [int(float(self.state[p]["step"])) for p in group["params"] if p in self.state and "step" in self.state[p]]
            if steps:
                group["step"] = max(steps)
            group["capturable"] = self.capturable
            group.setdefault("foreach", True)
        self._dev = {}          # device counters and tables are rebuilt from the loaded groups at the next step

    @torch.no_grad()
    def step(self, closure=None):
        from .. import _lib
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.get_lib()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            entries = []
            for p in ps:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                assert p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32
                entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
            group["step"] = group.get("step", 0) + 1
            d = self._dev.setdefault(gi, {})
            key = tuple(t.data_ptr() for e in entries for t in e)
            if d.get("key") != key:
                d["table"], d["key"] = _tensor_table(entries, dev), key
            b1, b2 = group["betas"]
            if self.capturable:
                if "lr" not in d:
                    d["lr"] = torch.full((1,), float(group["lr"]), dtype=torch.float32, device=dev)
                    d["step"] = torch.full((1,), group["step"] - 1, dtype=torch.int64, device=dev)
                elif not torch.cuda.is_current_stream_capturing():
                    d["lr"].fill_(float(group["lr"]))
                with torch.cuda.device(dev):
                    _lib.check(lib.nr_adam_step_dev(_lib.ptr(d["table"]), len(entries), _lib.ptr(d["lr"]), float(b1), float(b2),
                                                    float(group["eps"]), _lib.ptr(d["step"]), _lib.stream_ptr(dev)), "adam_step_dev")
            else:
                with torch.cuda.device(dev):
                    _lib.check(lib.nr_adam_step(_lib.ptr(d["table"]), len(entries), float(group["lr"]), float(b1), float(b2),
                                                float(group["eps"]), int(group["step"]), _lib.stream_ptr(dev)), "adam_step")
            # the kernel wrote through raw pointers: tell torch, the packed-weight caches of models/base.py key on _version
            torch.autograd.graph.increment_version(ps)
        return loss


class CapturedStep:
    """A whole training iteration -- ``volume_render`` under autograd, the losses, ``backward()``, the gradient
    all-reduce if any, ``optimizer.step()`` (train.py:196-210) -- captured ONCE into a CUDA graph and replayed: the
    ~1 100 launches of a 512-ray NeuS step become one graph launch (22.0 -> 19.3 ms on a B200).

        step = CapturedStep(step_fn, (rays_o, rays_d, target_rgb), optimizer=opt)
        for it in range(...):
            scheduler.step(it)                                     # writes group["lr"] on the host
            losses = step(rays_o_it, rays_d_it, target_it)         # copies into the static inputs, replays

    ``step_fn(*inputs)`` must be free of host reads (no ``.item()``; `neus_losses`, `grad_norm_device` and
    `FusedAdam(capturable=True)` / ``torch.optim.Adam(capturable=True)`` are) and must zero the gradients with
    ``zero_grad(set_to_none=False)`` so that the buffers the graph writes stay the same.  Its return value (tensors or
    a dict / tuple of tensors) is static storage overwritten by every replay.  Input shapes are fixed; random
    numbers inside (``perturb=True``) advance per replay through torch's graph-safe generator.  With a gradient
    all-reduce inside (NCCL captures), drop the CapturedStep before ``destroy_process_group()``."""

    def __init__(self, step_fn, example_inputs, optimizer=None, warmup=3):
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("neurecon_b200: CapturedStep needs CUDA tensors (there is no CPU fallback)")
        self.optimizer = optimizer
        self._params = [p for g in optimizer.param_groups for p in g["params"]] if optimizer is not None else []
        self.inputs = [t.clone() for t in example_inputs]
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                          # allocator / lazy-init warm-up outside the capture
                for _ in range(max(int(warmup), 1)):
                    step_fn(*self.inputs)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = step_fn(*self.inputs)
        self.warmup_steps = max(int(warmup), 1)

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        if self.optimizer is not None and hasattr(self.optimizer, "push_lr"):
            self.optimizer.push_lr()
        self.graph.replay()
        if self._params:                                            # the replay changed them behind torch's back (see FusedAdam.step)
            torch.autograd.graph.increment_version(self._params)
        return self.outputs
