"""Training path of the MLPs: ``torch.autograd.Function``s over the fp32 CUDA building blocks of
``csrc/mlp_f32.cu`` (nr_gemm_f32, nr_gemm_tn_f32, nr_colsum_f32, nr_sdf_bwd_act_f32, nr_act_bwd_f32,
nr_embed_f32).

The reference differentiates ``ImplicitSurface.forward_with_nablas`` twice (``autograd.grad(...,
create_graph=True)``, models/base.py:265-282, then ``loss.backward()`` through the eikonal term,
neus.py:443-458).  Here the normal is produced by forward-mode differentiation inside the network
(value rows h, tangent rows t_c sharing the weights), so ONE hand-written first-order backward of
that extended network covers the whole second-order path:

    forward:   z = W h + b,  u_c = W t_c,  h' = sp(z),  t'_c = sp'(z) u_c
    backward:  g_u_c = g_t'_c sp'(z)
               g_z   = g_h' sp'(z) + sum_c g_t'_c u_c sp''(z)
               g_W   = g_z^T h + sum_c g_u_c^T t_c,   g_b = colsum(g_z)
               g_h   = g_z W,  g_t_c = g_u_c W

Weight-norm (``W = g v/||v||``) stays in PyTorch on top of the effective weights, so ``weight_g`` /
``weight_v`` receive their gradients through ordinary autograd.  The sample points are constants
(they come from the no-grad up-sampling), so no gradient w.r.t. x is produced.
"""
import math

import torch

from .. import _lib

MODE_NONE, MODE_SOFTPLUS, MODE_RELU, MODE_SIGMOID, MODE_TANGENT, MODE_LINEAR = 0, 1, 2, 3, 4, 5


def _pad4(k):
    return (k + 3) & ~3


def _padded(W):
    """[out, in] -> contiguous fp32 [out, pad4(in)] (zero filled)."""
    out_d, in_d = W.shape
    if in_d % 4 == 0:
        return W.float().contiguous()
    Wp = torch.zeros(out_d, _pad4(in_d), dtype=torch.float32, device=W.device)
    Wp[:, :in_d] = W
    return Wp


import os

# Gradient operands of the tensor tier.  Round 1 converted them to bf16 ("for their range"): 8 bits of mantissa per
# rounding put 8e-2 on the first layers' weight gradients of a NeuS step (tests/test_gpu_train_golden.py against the
# reference's own gradients).  They now go in as fp16 (11 bits) behind a power-of-two loss scale applied INSIDE each
# backward: incoming gradients x 2^k, outgoing gradients x 2^-k -- exact, and invisible to the caller.  2^k is chosen
# for the magnitudes of the reference's losses (mean over rays and samples: upstream gradients of 1e-7 .. 1e-3, which
# 2^14 lifts into fp16's normal range with a factor ~1e3 of headroom below 65504).
_GRAD_SCALE = float(2 ** int(os.environ.get("NEURECON_B200_GRAD_SCALE_LOG2", "14")))


# fp16 tier: the SDF network trains through the reverse-mode 16-bit path of models/autograd_rev.py
# (NEURECON_B200_TRAIN=forward: round 1's forward-mode tangent path on fp32 rows)
_REVERSE_TRAINING = os.environ.get("NEURECON_B200_TRAIN", "reverse") != "forward"


def _tc():
    """Training GEMMs on the tensor cores (csrc/gemm_tc.cu) in the fp16 / bf16 tiers, fp32 SIMT in the fp32 tier."""
    return _lib.tensor_tier()


def _gemm(A, K, W, bias, N, mode, S=None, aux=None, m_val=0, grad=False, out=None):
    """Y[M, pad4(N)] = epilogue(A[:, :K] @ W[:N, :K]^T + bias); pad columns are zero.  ``grad``: the operands are
    gradients (tensor tier: bf16 operands for their range instead of fp16).  ``out``: a wider row-major buffer whose
    first pad4(N) columns receive the result (the head of a skip layer's concatenated input)."""
    lib = _lib.get_lib()
    M = A.shape[0]
    if out is not None:
        assert out.shape[0] == M and out.shape[1] >= _pad4(N) and out.is_contiguous()
        Y, ldy = out, out.shape[1]
    else:
        ldy = _pad4(N)
        Y = torch.zeros(M, ldy, dtype=torch.float32, device=A.device) if ldy != N else torch.empty(M, ldy, dtype=torch.float32, device=A.device)
    st = _lib.stream_ptr(A.device)
    if _tc():
        f16 = 1 if _lib.operand() == "fp16" else 0       # gradient operands too: fp16 behind the loss scale
        for n0 in range(0, N, 256):                       # the MMA's N is at most 256 (the last SDF layer has 257 rows)
            nn = min(256, N - n0)
            off = lambda t, ld: None if t is None else t[:, n0:]
            _lib.check(lib.nr_gemm_tc(
                _lib.ptr(A), A.shape[1], _lib.ptr(W[n0:]), W.shape[1], _lib.ptr(None if bias is None else bias[n0:]), M, nn, K,
                _lib.ptr(Y[:, n0:]), ldy, mode, _lib.ptr(off(S, 0)), 0 if S is None else S.shape[1], _lib.ptr(off(aux, 0)),
                0 if aux is None else aux.shape[1], m_val, f16, st), "gemm_tc")
        return Y
    _lib.check(lib.nr_gemm_f32(_lib.ptr(A), A.shape[1], _lib.ptr(W), W.shape[1], _lib.ptr(bias), M, N, K, _lib.ptr(Y), ldy,
                               mode, _lib.ptr(S), 0 if S is None else S.shape[1], _lib.ptr(aux),
                               0 if aux is None else aux.shape[1], m_val, st), "gemm_f32")
    return Y


def _gemm_tn(G, N, X, K, dW):
    """dW[:N, :K] += G[:, :N]^T X[:, :K]"""
    lib = _lib.get_lib()
    if _tc():
        _lib.check(lib.nr_gemm_tn_tc(_lib.ptr(G), G.shape[1], _lib.ptr(X), X.shape[1], G.shape[0], N, K, _lib.ptr(dW),
                                     dW.shape[1], 1 if _lib.operand() == "fp16" else 0, _lib.stream_ptr(G.device)),
                   "gemm_tn_tc")
        return
    _lib.check(lib.nr_gemm_tn_f32(_lib.ptr(G), G.shape[1], _lib.ptr(X), X.shape[1], G.shape[0], N, K, _lib.ptr(dW),
                                  dW.shape[1], _lib.stream_ptr(G.device)), "gemm_tn_f32")


def _scale_in(*gs):
    """incoming gradients times the loss scale (tensor tier with fp16 operands only)"""
    if not (_tc() and _lib.operand() == "fp16"):
        return gs + (1.0,)
    return tuple(None if g is None else g * _GRAD_SCALE for g in gs) + (1.0 / _GRAD_SCALE,)


def _unscale(grads, inv):
    return grads if inv == 1.0 else [None if g is None else g * inv for g in grads]


def _colsum(G, N):
    lib = _lib.get_lib()
    out = torch.zeros(N, dtype=torch.float32, device=G.device)
    _lib.check(lib.nr_colsum_f32(_lib.ptr(G), G.shape[1], G.shape[0], N, _lib.ptr(out), _lib.stream_ptr(G.device)), "colsum_f32")
    return out


class _SdfFn(torch.autograd.Function):
    """(x, W_0, b_0, ..., W_D, b_D) -> (sdf [n], nabla [n,3], feat [n,F]).  W_l are the EFFECTIVE weights
    (the skip layer's already divided by sqrt 2)."""

    @staticmethod
    def forward(ctx, x, multires, skip, want_nablas, *wb):
        lib = _lib.get_lib()
        dev = x.device
        n = x.shape[0]
        L = len(wb) // 2
        Ws = [_padded(w.detach().float()) for w in wb[0::2]]
        bs = [b.detach().float().contiguous() for b in wb[1::2]]
        dims = [(w.shape[0], w.shape[1]) for w in wb[0::2]]  # (out, in)
        pe_dim = 3 if multires < 0 else 3 * (1 + 2 * multires)
        ldpe = _pad4(pe_dim)
        f = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            pe = torch.zeros(n, ldpe, **f)
            tpe = torch.zeros(3 * n, ldpe, **f) if want_nablas else None
            _lib.check(lib.nr_embed_f32(_lib.ptr(x), n, 3, multires, _lib.ptr(pe), ldpe, 0, _lib.ptr(tpe), ldpe, 0,
                                        _lib.stream_ptr(dev)), "embed_f32")
            h, t = pe, tpe
            saved = []
            for l in range(L - 1):
                N, K = dims[l]
                S = torch.zeros(n, _pad4(N), **f) if _pad4(N) != N else torch.empty(n, N, **f)   # the epilogue writes every column < N
                h_buf = t_buf = None
                if l + 1 == skip:   # this layer's output is the head of the skip layer's input [h | pe]: the GEMMs write it in place
                    ldn = _pad4(dims[skip][1])
                    h_buf = torch.empty(n, ldn, **f)
                    t_buf = torch.empty(3 * n, ldn, **f) if want_nablas else None
                h_out = _gemm(h, K, Ws[l], bs[l], N, MODE_SOFTPLUS, S=S, out=h_buf)
                t_out = None
                if want_nablas:   # t' = sp'(z) * (W t), the scaling in the GEMM's epilogue (row r of t belongs to point r % n)
                    t_out = _gemm(t, K, Ws[l], None, N, MODE_TANGENT, aux=S, m_val=n, out=t_buf)
                for buf, src in ((h_buf, pe), (t_buf, tpe)):
                    if buf is not None:
                        buf[:, N:N + pe_dim] = src[:, :pe_dim]
                        if buf.shape[1] > N + pe_dim:
                            buf[:, N + pe_dim:] = 0
                saved.append((h, t, S))
                h, t = h_out, t_out
            N, K = dims[L - 1]
            out = _gemm(h, K, Ws[L - 1], bs[L - 1], N, MODE_NONE)
            sdf = out[:, 0].contiguous()
            feat = out[:, 1:N].contiguous()
            if want_nablas:
                tout = _gemm(t, K, Ws[L - 1], None, 1, MODE_LINEAR)       # [3n, 4], column 0 valid
                nabla = tout[:, 0].reshape(3, n).t().contiguous()
            else:
                nabla = torch.zeros(n, 3, **f)
        ctx.saved = saved
        ctx.last = (h, t)
        ctx.Ws, ctx.dims, ctx.skip, ctx.want_nablas, ctx.n = Ws, dims, skip, want_nablas, n
        ctx.mark_non_differentiable(*([nabla] if not want_nablas else []))
        return sdf, nabla, feat

    @staticmethod
    def backward(ctx, g_sdf, g_nabla, g_feat):
        lib = _lib.get_lib()
        Ws, dims, skip, wn, n = ctx.Ws, ctx.dims, ctx.skip, ctx.want_nablas, ctx.n
        L = len(Ws)
        h_last, t_last = ctx.last
        dev = h_last.device
        f = dict(dtype=torch.float32, device=dev)
        grads = [None] * (2 * L)
        g_sdf, g_nabla, g_feat, inv = _scale_in(g_sdf, g_nabla, g_feat)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            # ---- last (linear) layer: rows = [sdf | feat] ----
            N, K = dims[L - 1]
            g_out = torch.zeros(n, _pad4(N), **f)
            if g_sdf is not None:
                g_out[:, 0] = g_sdf
            if g_feat is not None:
                g_out[:, 1:N] = g_feat
            dW = torch.zeros(N, _pad4(K), **f)
            _gemm_tn(g_out, N, h_last, K, dW)
            grads[2 * (L - 1) + 1] = _colsum(g_out, N)
            Wt = _padded(Ws[L - 1][:, :K].t().contiguous())
            g_h = _gemm(g_out, N, Wt, None, K, MODE_LINEAR, grad=True)
            g_t = None
            if wn:
                g_tout = torch.zeros(3 * n, 4, **f)
                if g_nabla is not None:
                    g_tout[:, 0] = g_nabla.t().reshape(-1)
                _gemm_tn(g_tout, 1, t_last, K, dW)                      # adds to row 0
                w_row = torch.zeros(1, _pad4(K), **f)
                w_row[:, :K] = Ws[L - 1][0:1, :K]
                g_t = g_tout[:, :1] * w_row                            # [3n, pad4(K)], pad columns zero
            grads[2 * (L - 1)] = dW[:, :K]
            # ---- hidden softplus layers ----
            for l in range(L - 2, -1, -1):
                N, K = dims[l]
                h_in, t_in, S = ctx.saved[l]
                if wn:
                    # the layer's output tangent S * u is the next layer's saved input (its first N columns at the skip)
                    t_scaled = ctx.saved[l + 1][1] if l + 1 <= L - 2 else t_last
                    g_b = torch.zeros(N, **f)                          # bias gradient = column sums of g_z, from the same pass
                    _lib.check(lib.nr_sdf_bwd_act_f32(_lib.ptr(g_h), g_h.shape[1], _lib.ptr(g_t), g_t.shape[1], _lib.ptr(S),
                                                      S.shape[1], _lib.ptr(t_scaled), t_scaled.shape[1], n, N, 1, _lib.ptr(g_b), st),
                               "sdf_bwd_act")
                else:
                    g_h[:, :N] *= S[:, :N]
                    g_b = None
                dW = torch.zeros(N, _pad4(K), **f)
                _gemm_tn(g_h, N, h_in, K, dW)
                if wn:
                    _gemm_tn(g_t, N, t_in, K, dW)
                grads[2 * l] = dW[:, :K]
                grads[2 * l + 1] = g_b if g_b is not None else _colsum(g_h, N)
                if l > 0:
                    Wt = _padded(Ws[l][:, :K].t().contiguous())
                    g_h = _gemm(g_h, N, Wt, None, K, MODE_LINEAR, grad=True)      # [n, pad4(K)]; for the skip layer only the
                    if wn:                                             # first out_{l-1} columns are used below
                        g_t = _gemm(g_t, N, Wt, None, K, MODE_LINEAR, grad=True)
        ctx.saved = ctx.last = None
        return (None, None, None, None, *_unscale(grads, inv))


class _RadianceFn(torch.autograd.Function):
    """(x, view, normals, feat, W_0, b_0, ...) -> rgb [n,3]; gradients for normals, feat and the weights."""

    @staticmethod
    def forward(ctx, x, view, normals, feat, multires, multires_view, *wb):
        lib = _lib.get_lib()
        dev = x.device
        n = x.shape[0]
        L = len(wb) // 2
        Ws = [_padded(w.detach().float()) for w in wb[0::2]]
        bs = [b.detach().float().contiguous() for b in wb[1::2]]
        dims = [(w.shape[0], w.shape[1]) for w in wb[0::2]]
        px = 3 if multires < 0 else 3 * (1 + 2 * multires)
        pv = 3 if multires_view < 0 else 3 * (1 + 2 * multires_view)
        in0 = px + pv + 3 + feat.shape[1]
        assert in0 == dims[0][1], "RadianceNet layer 0 width"
        f = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            a0 = torch.zeros(n, _pad4(in0), **f)
            _lib.check(lib.nr_embed_f32(_lib.ptr(x), n, 3, multires, _lib.ptr(a0), a0.shape[1], 0, None, 0, 0, st), "embed")
            _lib.check(lib.nr_embed_f32(_lib.ptr(view), n, 3, multires_view, _lib.ptr(a0), a0.shape[1], px, None, 0, 0, st), "embed")
            a0[:, px + pv:px + pv + 3] = normals
            a0[:, px + pv + 3:in0] = feat
            acts = [a0]
            h = a0
            for l in range(L):
                N, K = dims[l]
                h = _gemm(h, K, Ws[l], bs[l], N, MODE_SIGMOID if l == L - 1 else MODE_RELU)
                acts.append(h)
        ctx.acts, ctx.Ws, ctx.dims, ctx.n = acts, Ws, dims, n
        ctx.split = (px + pv, in0)
        return acts[-1][:, :3].contiguous()

    @staticmethod
    def backward(ctx, g_rgb):
        lib = _lib.get_lib()
        acts, Ws, dims, n = ctx.acts, ctx.Ws, ctx.dims, ctx.n
        L = len(Ws)
        dev = acts[0].device
        f = dict(dtype=torch.float32, device=dev)
        grads = [None] * (2 * L)
        g_rgb, inv = _scale_in(g_rgb)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            g = torch.zeros(n, 4, **f)
            g[:, :3] = g_rgb
            for l in range(L - 1, -1, -1):
                N, K = dims[l]
                _lib.check(lib.nr_act_bwd_f32(_lib.ptr(g), g.shape[1], _lib.ptr(acts[l + 1]), acts[l + 1].shape[1], n, N,
                                              1 if l == L - 1 else 0, st), "act_bwd")
                dW = torch.zeros(N, _pad4(K), **f)
                _gemm_tn(g, N, acts[l], K, dW)
                grads[2 * l] = dW[:, :K]
                grads[2 * l + 1] = _colsum(g, N)
                Wt = _padded(Ws[l][:, :K].t().contiguous())
                g = _gemm(g, N, Wt, None, K, MODE_LINEAR, grad=True)
        off, in0 = ctx.split
        g_normals = g[:, off:off + 3].contiguous() * inv
        g_feat = g[:, off + 3:in0].contiguous() * inv
        ctx.acts = None
        return (None, None, g_normals, g_feat, None, None, *_unscale(grads, inv))


class _WeightNormFn(torch.autograd.Function):
    """(scales, g_0, v_0, g_1, v_1, ...) -> (W_0, W_1, ...), W_l = scale_l g_l v_l / ||v_l||_row: the weight normalisation
    of every layer of a network in one launch, its backward in one more (csrc/weight_norm.cu) -- instead of three torch
    kernels per layer forward and half a dozen backward."""

    @staticmethod
    def _table(entries, dev):
        import struct
        rows, words = 0, b""
        for v, g, W, dW, dv, dg, scale in entries:
            words += struct.pack("<6q4qdq", v.data_ptr(), g.data_ptr(), W.data_ptr(), 0 if dW is None else dW.data_ptr(),
                                 0 if dv is None else dv.data_ptr(), 0 if dg is None else dg.data_ptr(), v.shape[0], v.shape[1],
                                 W.stride(0), 0 if dW is None else dW.stride(0), float(scale), rows)
            rows += v.shape[0]
        return _lib.C.create_string_buffer(words, len(words)), rows          # host memory: the kernel takes the table by value

    @staticmethod
    def forward(ctx, scales, *gv):
        gs, vs = gv[0::2], gv[1::2]
        dev = vs[0].device
        Ws = [torch.empty(v.shape[0], (v.shape[1] + 3) & ~3, dtype=torch.float32, device=dev) for v in vs]
        ent = [(v.detach(), g.detach(), W, None, None, None, sc) for v, g, W, sc in zip(vs, gs, Ws, scales)]
        tab, rows = _WeightNormFn._table(ent, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.get_lib().nr_weight_norm(tab, len(ent), rows, 0, _lib.stream_ptr(dev)), "weight_norm")
        ctx.save_for_backward(*gv)
        ctx.scales = scales
        return tuple(W[:, :v.shape[1]] for W, v in zip(Ws, vs))

    @staticmethod
    def backward(ctx, *dWs):
        gv = ctx.saved_tensors
        gs, vs = gv[0::2], gv[1::2]
        dev = vs[0].device
        ent, out = [], []
        for g, v, dW, sc in zip(gs, vs, dWs, ctx.scales):
            dW = torch.zeros_like(v) if dW is None else dW
            if dW.stride(1) != 1 or dW.dtype != torch.float32:
                dW = dW.float().contiguous()
            dg, dv = torch.empty_like(g), torch.empty_like(v)
            # W is only read by the forward kernel; the backward gets a dummy
            ent.append((v, g, dv, dW, dv, dg, sc))
            out += [dg, dv]
        tab, rows = _WeightNormFn._table(ent, dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.get_lib().nr_weight_norm(tab, len(ent), rows, 1, _lib.stream_ptr(dev)), "weight_norm")
        return (None, *out)


def _fused_effective_weights(layers, scales):
    """effective weights of weight-normed layers through one kernel (tensor tier); None if a layer is not weight-normed"""
    if not all(hasattr(l, "weight_g") for l in layers) or not layers[0].weight_v.is_cuda:
        return None
    gv = []
    for l in layers:
        gv += [l.weight_g, l.weight_v]
    return list(_WeightNormFn.apply(tuple(scales), *gv))


def _surface_weights(surface):
    from .base import _effective_weight
    layers = list(surface.surface_fc_layers)
    scales = [1.0 / math.sqrt(2) if i in surface.skips else 1.0 for i in range(len(layers))]
    Ws = _fused_effective_weights(layers, scales) if (_tc() and torch.is_grad_enabled()) else None
    wb = []
    for i, layer in enumerate(layers):
        if Ws is not None:
            W = Ws[i]
        else:
            W = _effective_weight(layer)
            if i in surface.skips:
                W = W / math.sqrt(2)
        wb += [W, layer.bias]
    return wb


def _run_sdf(surface, x, want_nablas):
    _lib.require_cuda(x)
    surface._check_supported()
    shape = x.shape[:-1]
    xf = _lib.f32c(x.detach().reshape(-1, 3))
    skip = surface.skips[0] if surface.skips else -1
    wb = _surface_weights(surface)
    if want_nablas and _tc() and _lib.operand() == "fp16" and _REVERSE_TRAINING:
        from . import autograd_rev
        if autograd_rev.supported([tuple(w.shape) for w in wb[0::2]], skip, surface.embed_multires):
            sdf, nabla, feat = autograd_rev.SdfRevFn.apply(xf, surface.embed_multires, skip, *wb)
            return sdf.reshape(shape), nabla.reshape(*shape, 3), feat.reshape(*shape, -1)
    sdf, nabla, feat = _SdfFn.apply(xf, surface.embed_multires, skip, want_nablas, *wb)
    return sdf.reshape(shape), nabla.reshape(*shape, 3), feat.reshape(*shape, -1)


def sdf_forward_autograd(surface, x, return_h=False):
    """ImplicitSurface.forward under autograd (models/base.py:243-263)."""
    sdf, _, feat = _run_sdf(surface, x, want_nablas=False)
    return (sdf, feat) if return_h else sdf


def sdf_forward_with_nablas_autograd(surface, x):
    """ImplicitSurface.forward_with_nablas with has_grad=True (models/base.py:265-282): the returned
    nabla is differentiable w.r.t. the weights (what create_graph=True provides in the reference)."""
    return _run_sdf(surface, x, want_nablas=True)


def radiance_forward_autograd(rad, x, view_dirs, normals, geometry_feature):
    """RadianceNet.forward under autograd (models/base.py:372-391)."""
    from .base import _effective_weight
    if rad.skips:
        raise NotImplementedError("neurecon_b200 RadianceNet supports skips=[]")
    _lib.require_cuda(x, view_dirs, normals, geometry_feature)
    shape = x.shape[:-1]
    xf = _lib.f32c(x.detach().reshape(-1, 3))
    if rad.use_view_dirs:
        vf = _lib.f32c(view_dirs.detach().expand(*shape, 3).reshape(-1, 3))
        nf = normals.reshape(-1, 3).float()
    else:       # base.py:383-384: neither enters the network; layer 0's weight has zero columns for them (base.py here)
        vf, nf = xf, xf
    ff = geometry_feature.reshape(-1, geometry_feature.shape[-1]).float()
    layers = list(rad.layers)
    Ws = _fused_effective_weights(layers, [1.0] * len(layers)) if (_tc() and torch.is_grad_enabled()) else None
    wb = []
    for i, layer in enumerate(layers):
        wb += [Ws[i] if Ws is not None else _effective_weight(layer), layer.bias]
    wb[0] = rad._layer0_weight(wb[0])
    mv = rad._multires_view_eff
    if _tc() and _lib.operand() == "fp16" and _REVERSE_TRAINING and all(tuple(w.shape)[0] == 256 for w in wb[0:-2:2]) \
            and wb[0].shape[1] > 256:
        from . import autograd_rev
        rgb = autograd_rev.RadianceRevFn.apply(xf, vf, nf, ff, rad.embed_multires, mv, *wb)
        return rgb.reshape(*shape, 3)
    rgb = _RadianceFn.apply(xf, vf, nf, ff, rad.embed_multires, mv, *wb)
    return rgb.reshape(*shape, 3)


class _NerfFn(torch.autograd.Function):
    """NeRF++ background net (models/base.py:426-453, use_view_dirs=True) under autograd:
    (x [n,dim], view [n,3], pts W/b ..., alpha W/b, feature W/b, views W/b, rgb W/b) -> (sigma [n], rgb [n,3]).
    Gradients w.r.t. the weights only (the inputs are sample positions)."""

    @staticmethod
    def forward(ctx, x, view, dim, multires, multires_view, skip, D, *wb):
        lib = _lib.get_lib()
        n, dev = x.shape[0], x.device
        Ws = [_padded(w.detach().float()) for w in wb[0::2]]
        bs = [b.detach().float().contiguous() for b in wb[1::2]]
        dims = [(w.shape[0], w.shape[1]) for w in wb[0::2]]
        npe = dim * (1 if multires < 0 else 1 + 2 * multires)
        npv = 3 if multires_view < 0 else 3 * (1 + 2 * multires_view)
        f = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            pe = torch.zeros(n, _pad4(npe), **f)
            _lib.check(lib.nr_embed_f32(_lib.ptr(x), n, dim, multires, _lib.ptr(pe), pe.shape[1], 0, None, 0, 0, st), "embed")
            pv = torch.zeros(n, _pad4(npv), **f)
            _lib.check(lib.nr_embed_f32(_lib.ptr(view), n, 3, multires_view, _lib.ptr(pv), pv.shape[1], 0, None, 0, 0, st), "embed")
            ins, outs = [], []           # input / output activation of every pts layer
            h = pe
            for i in range(D):
                N, K = dims[i]
                ins.append(h)
                h = _gemm(h, K, Ws[i], bs[i], N, MODE_RELU)
                outs.append(h)
                if i == skip:            # cat([PE(x), h])  (base.py:436)
                    cat = torch.zeros(n, _pad4(npe + N), **f)
                    cat[:, :npe] = pe[:, :npe]
                    cat[:, npe:npe + N] = h[:, :N]
                    h = cat
            W_ = dims[D - 1][0]
            ia, ife, iv, ir = D, D + 1, D + 2, D + 3
            sigma = _gemm(h, W_, Ws[ia], bs[ia], 1, MODE_NONE)
            feat = _gemm(h, W_, Ws[ife], bs[ife], dims[ife][0], MODE_NONE)
            vin = torch.zeros(n, _pad4(dims[ife][0] + npv), **f)
            vin[:, :dims[ife][0]] = feat[:, :dims[ife][0]]
            vin[:, dims[ife][0]:dims[ife][0] + npv] = pv[:, :npv]
            hv = _gemm(vin, dims[iv][1], Ws[iv], bs[iv], dims[iv][0], MODE_RELU)
            rgb = _gemm(hv, dims[ir][1], Ws[ir], bs[ir], 3, MODE_SIGMOID)
        ctx.state = (ins, outs, h, vin, hv, rgb, Ws, dims, npe, skip, D, n)
        return sigma[:, 0].contiguous(), rgb[:, :3].contiguous()

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        lib = _lib.get_lib()
        ins, outs, h_last, vin, hv, rgb, Ws, dims, npe, skip, D, n = ctx.state
        dev = h_last.device
        f = dict(dtype=torch.float32, device=dev)
        grads = [None] * (2 * (D + 4))
        ia, ife, iv, ir = D, D + 1, D + 2, D + 3
        g_sigma, g_rgb, inv = _scale_in(g_sigma, g_rgb)

        def layer_bwd(idx, g, x_in, act=None, sigmoid=False):
            """g [n, pad4(N)] = grad w.r.t. the layer's output -> weight / bias grads; returns grad w.r.t. its input."""
            N, K = dims[idx]
            if act is not None:
                _lib.check(lib.nr_act_bwd_f32(_lib.ptr(g), g.shape[1], _lib.ptr(act), act.shape[1], n, N, 1 if sigmoid else 0,
                                              _lib.stream_ptr(dev)), "act_bwd")
            dW = torch.zeros(N, _pad4(K), **f)
            _gemm_tn(g, N, x_in, K, dW)
            grads[2 * idx] = dW[:, :K]
            grads[2 * idx + 1] = _colsum(g, N)
            Wt = _padded(Ws[idx][:, :K].t().contiguous())
            return _gemm(g, N, Wt, None, K, MODE_LINEAR, grad=True)

        with torch.cuda.device(dev):
            g = torch.zeros(n, 4, **f)
            if g_rgb is not None:
                g[:, :3] = g_rgb
            g_hv = layer_bwd(ir, g, hv, act=rgb, sigmoid=True)
            g_vin = layer_bwd(iv, g_hv, vin, act=hv)
            nf = dims[ife][0]
            g_feat = torch.zeros(n, _pad4(nf), **f)
            g_feat[:, :nf] = g_vin[:, :nf]
            g_h = layer_bwd(ife, g_feat, h_last)
            ga = torch.zeros(n, 4, **f)
            if g_sigma is not None:
                ga[:, 0] = g_sigma
            g_h = g_h + layer_bwd(ia, ga, h_last)
            for i in range(D - 1, -1, -1):
                N, K = dims[i]
                if i == skip:            # the next layer's input was cat([PE, h_i]): keep the h part
                    gh = torch.zeros(n, _pad4(N), **f)
                    gh[:, :N] = g_h[:, npe:npe + N]
                    g_h = gh
                g_h = layer_bwd(i, g_h, ins[i], act=outs[i])
        ctx.state = None
        return (None, None, None, None, None, None, None, *_unscale(grads, inv))


def nerf_forward_autograd(module, input_pts, input_views):
    """NeRF.forward under autograd (models/base.py:426-453)."""
    if not module.use_view_dirs or len(module.skips) > 1:
        raise NotImplementedError("neurecon_b200 NeRF supports use_view_dirs=True with at most one skip")
    _lib.require_cuda(input_pts, input_views)
    shape = input_pts.shape[:-1]
    xf = _lib.f32c(input_pts.detach().reshape(-1, module.input_dim))
    vf = _lib.f32c(input_views.detach().expand(*shape, 3).reshape(-1, 3))
    wb = []
    for lin in list(module.pts_linears) + [module.alpha_linear, module.feature_linear, module.views_linears[0], module.rgb_linear]:
        wb += [lin.weight, lin.bias]
    skip = module.skips[0] if module.skips else -1
    sigma, rgb = _NerfFn.apply(xf, vf, module.input_dim, module.multires, module.multires_view, skip, module.D, *wb)
    return sigma.reshape(shape), rgb.reshape(*shape, 3)


class _ExclusiveCumprodFn(torch.autograd.Function):
    """T_i = prod_{j<i} p_j along the last axis: the transmittance of neus.py:346-347 / volsdf.py:487 /
    unisurf.py:210 (``cumprod(cat([1, p]))[..., :-1]``).  Same forward as torch; the backward is written
    out because ``torch.cumprod``'s reads an "any zero?" flag on the host, which keeps a training step
    from being captured into a CUDA graph.  Zeros are handled on the device: the first zero of a row
    gets the product of the others, everything behind it zero -- torch's result."""

    @staticmethod
    def forward(ctx, p):
        T = torch.cumprod(torch.cat([torch.ones_like(p[..., :1]), p[..., :-1]], dim=-1), dim=-1)
        ctx.save_for_backward(p, T)
        return T

    @staticmethod
    def backward(ctx, g):
        p, T = ctx.saved_tensors

        def suffix_sum_excl(a):                            # sum_{i>j} a_i
            return a.flip(-1).cumsum(-1).flip(-1) - a

        zero = p == 0
        first = zero & (zero.cumsum(-1) == 1)
        T1 = torch.cumprod(torch.cat([torch.ones_like(p[..., :1]), torch.where(first, torch.ones_like(p), p)[..., :-1]], dim=-1), dim=-1)
        regular = suffix_sum_excl(g * T) / torch.where(zero, torch.ones_like(p), p)
        return torch.where(first, suffix_sum_excl(g * T1), torch.where(zero, torch.zeros_like(p), regular))


def exclusive_cumprod(p):
    return _ExclusiveCumprodFn.apply(p)
