"""Training path of the SDF network on the tensor tier, REVERSE-MODE formulation with 16-bit tensors in HBM.

The reference differentiates ``ImplicitSurface.forward_with_nablas`` twice (``autograd.grad(..., create_graph=True)``,
models/base.py:265-282, then ``loss.backward()`` through the eikonal term, neus.py:443-458).  Round 1 produced the
normal by forward-mode differentiation (three tangent rows per point through every layer, fp32 rows in HBM).  Here
the normal is computed the way autograd does it -- a reverse sweep -- and the gradient of (sdf, nabla, feature) w.r.t.
the weights by the adjoint of BOTH sweeps, four GEMM chains over n rows instead of two over 4 n:

    forward sweep    z_l = W_l h_l + b_l,  h_{l+1} = sp(z_l),  S_l = sp'(z_l)                   l = 0 .. D-1
                     [sdf | feat] = W_D h_D + b_D
    reverse sweep    p_{D-1} = S_{D-1} * W_D[0, :],   p_{l-1} = S_{l-1} * (W_l^T p_l),   nabla = J^T (W_0^T p_0 + g_e)
    adjoint of it    gb_0 = J nb,   q_l = W_l gb_l,   gb_{l+1} = S_l * q_l,   zb2_l = sp''(z_l) g_{l+1} q_l
                                                                                  = 100 (1 - S_l) p_l q_l   (no division)
    backprop         zb_{D-1} = S_{D-1} * (W_D^T yb) + zb2_{D-1},   zb_{l-1} = S_{l-1} * (W_l^T zb_l) + zb2_{l-1}
    weights          dW_l = zb_l^T h_l + p_l^T gb_l,   db_l = colsum(zb_l),   dW_D = yb^T h_D (+ colsum(gb_D) on the sdf row)

(J = d PE / d x; the skip layer's input is [h | PE(x)] with its weight pre-divided by sqrt 2, so its reverse and
adjoint GEMMs carry the embedding part g_e / J nb in the last 39 columns of the 256-wide rows.)  Every matrix product
is one ``nr_gemm16`` launch (csrc/gemm16.cu) with the elementwise part in its epilogue; h, S, p, gb, zb2, zb are fp16
rows of 256; gradients run behind a power-of-two loss scale that the dW / db kernels take out again.
"""
import torch

from .. import _lib
from .autograd import _GRAD_SCALE

G_LINEAR, G_SOFTPLUS, G_SCALE, G_ADJ = 0, 1, 2, 3
WIDTH = 256


def _pad4c(W):
    """contiguous fp32 [out, pad4(in)]"""
    W = W.detach().float()
    out_d, in_d = W.shape
    if in_d % 4 == 0:
        return W.contiguous()
    Wp = torch.zeros(out_d, (in_d + 3) & ~3, dtype=torch.float32, device=W.device)
    Wp[:, :in_d] = W
    return Wp


import os
_PACK = os.environ.get("NEURECON_G16_PACK", "1") != "0"


class _Packed:
    """a weight matrix as the fp16 shared-memory image of nr_gemm16 (packed once, used by two GEMMs of the step)"""

    def __init__(self, W, N, K):
        lib = _lib.get_lib()
        self.img = torch.empty(int(lib.nr_gemm16_pack_w_bytes(N, K)), dtype=torch.uint8, device=W.device)
        _lib.check(lib.nr_gemm16_pack_w(_lib.ptr(W), W.stride(0), N, K, _lib.ptr(self.img), _lib.stream_ptr(W.device)), "gemm16_pack_w")


def _gemm16(A, W, bias, n, N, K, Y, y_half, mode, aux_a=None, aux_b=None, out2=None):
    lib = _lib.get_lib()
    packed = isinstance(W, _Packed)
    _lib.check(lib.nr_gemm16(
        _lib.ptr(A), A.stride(0), _lib.ptr(W.img if packed else W), 0 if packed else W.stride(0), _lib.ptr(bias), n, N, K,
        _lib.ptr(Y), Y.stride(0), int(y_half), mode, _lib.ptr(aux_a), 0 if aux_a is None else aux_a.stride(0), _lib.ptr(aux_b),
        0 if aux_b is None else aux_b.stride(0), _lib.ptr(out2), 0 if out2 is None else out2.stride(0), int(packed),
        _lib.stream_ptr(A.device)), "gemm16")


def split_forward():
    """Forward and reverse sweeps on split-precision operands (csrc/gemm16.cu nr_gemm16_split): the training path of the
    'fp16x2' tier and, by default, of the 'fp16' tier too; NEURECON_B200_TRAIN_SPLIT=0 keeps plain fp16 sweeps there (0.8 ms
    faster per 512-ray step).  Softplus(beta=100) turns a pre-activation error dz into 25 dz on softplus' and 2500 dz on
    softplus'' (the eikonal term's second-order path): with plain fp16 operands (dz ~ 3e-4) the worst weight gradient of a
    NeuS / VolSDF / UNISURF step is 1e-1 / 3e-1 / 2e-1 off the reference's, with split sweeps 4e-3 / 3e-3 / 4e-3 (measured)."""
    return _lib.split_tier() or os.environ.get("NEURECON_B200_TRAIN_SPLIT", "1") != "0"


class _PackedSplit:
    """[W_hi | W_hi | W_lo] (the K-concatenated split product's weight side) as nr_gemm16 images, one per block of 64 output
    columns; W: contiguous fp32 [>= N, pad4(K)]"""

    def __init__(self, W, N, K):
        lib = _lib.get_lib()
        self.img = torch.empty(int(lib.nr_gemm16_pack_w_split_bytes(N, K)), dtype=torch.uint8, device=W.device)
        _lib.check(lib.nr_gemm16_pack_w_split(_lib.ptr(W), W.stride(0), N, K, _lib.ptr(self.img), _lib.stream_ptr(W.device)),
                   "gemm16_pack_w_split")


def _gemm16_split(A, Wp, bias, n, N, K, Y, y_half, lo_off, mode, out2=None, aux_a=None):
    _lib.check(_lib.get_lib().nr_gemm16_split(
        _lib.ptr(A), A.stride(0), _lib.ptr(Wp.img), _lib.ptr(bias), n, N, K, _lib.ptr(Y), Y.stride(0), int(y_half), int(lo_off), mode,
        _lib.ptr(out2), 0 if out2 is None else out2.stride(0), _lib.ptr(aux_a), 0 if aux_a is None else aux_a.stride(0),
        _lib.stream_ptr(A.device)), "gemm16_split")


def _cast_cols16(src, scale, dst, lo_off=0):
    """dst[:, :ncols] = fp16(scale * src) (+ the lo part lo_off columns to the right) in one launch; src fp32 [n, ncols] with any
    row stride (unit column stride), dst a column-offset view of fp16 rows"""
    if src.dtype != torch.float32 or src.stride(-1) != 1:
        src = src.float().contiguous()
    if src.dim() == 1:
        src = src.unsqueeze(1)
    n, ncols = src.shape
    _lib.check(_lib.get_lib().nr_cast_cols16(_lib.ptr(src), src.stride(0), n, ncols, float(scale), _lib.ptr(dst), dst.stride(0),
                                             int(lo_off), _lib.stream_ptr(dst.device)), "cast_cols16")


def supported(dims, skip, multires):
    """the shapes this path is written for: hidden width 256, one skip whose input is [h | PE(x)] of width 256"""
    L = len(dims)
    pe = 3 if multires < 0 else 3 + 6 * multires
    if pe > 64 or dims[0][1] != pe or dims[L - 1][0] != WIDTH + 1:
        return False
    for l in range(L - 1):
        out_d, in_d = dims[l]
        want_out = WIDTH - pe if l + 1 == skip else WIDTH
        want_in = pe if l == 0 else WIDTH
        if out_d != want_out or in_d != want_in:
            return False
    return True


class SdfRevFn(torch.autograd.Function):
    """(x [n,3], W_0, b_0, ..., W_D, b_D) -> (sdf [n], nabla [n,3], feat [n,256]); W_l are the EFFECTIVE weights (the
    skip layer's already divided by sqrt 2)."""

    @staticmethod
    def forward(ctx, x, multires, skip, *wb):
        lib = _lib.get_lib()
        dev, n = x.device, x.shape[0]
        L = len(wb) // 2
        D = L - 1
        Ws = [_pad4c(w) for w in wb[0::2]]
        bs = [b.detach().float().contiguous() for b in wb[1::2]]
        dims = [(w.shape[0], w.shape[1]) for w in wb[0::2]]
        pe = 3 if multires < 0 else 3 + 6 * multires
        h16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            split = split_forward()
            if split:
                e = torch.empty(n, 128, **h16)                                # [PE hi | PE lo]
                _lib.check(lib.nr_pe16_split(_lib.ptr(x), n, multires, _lib.ptr(e), 128, 64, 64, None, 0, 0, 0, st), "pe16_split")
            else:
                e = torch.empty(n, 64, **h16)
                _lib.check(lib.nr_pe16(_lib.ptr(x), n, multires, _lib.ptr(e), 64, 64, None, 0, 0, st), "pe16")
            hs, S = [e], []
            Wp = [_Packed(Ws[l], dims[l][0], dims[l][1]) if _PACK else Ws[l] for l in range(D)]   # forward + adjoint sweeps
            for l in range(D):                                               # ---- forward sweep
                N, K = dims[l]
                Sl = torch.empty(n, WIDTH, **h16)
                if split:      # pre-activations from (hi, lo) operand pairs: softplus(beta=100) turns dz into 25 dz on softplus'
                    out = torch.empty(n, 2 * WIDTH, **h16)
                    _gemm16_split(hs[l], _PackedSplit(Ws[l], N, K), bs[l], n, N, K, out, 1, WIDTH, G_SOFTPLUS, out2=Sl)
                else:
                    out = torch.empty(n, WIDTH, **h16)
                    _gemm16(hs[l], Wp[l], bs[l], n, N, K, out, 1, G_SOFTPLUS, out2=Sl)
                if l + 1 == skip:                                            # the skip layer's input [h | PE(x)], S = 1 on the PE part
                    if split:
                        _lib.check(lib.nr_pe16_split(_lib.ptr(x), n, multires, _lib.ptr(e), 128, 64, 64, _lib.ptr(out), 2 * WIDTH, N,
                                                     WIDTH, st), "pe16_split")
                    else:
                        _lib.check(lib.nr_pe16(_lib.ptr(x), n, multires, _lib.ptr(e), 64, 64, _lib.ptr(out), WIDTH, N, st), "pe16")
                    Sl[:, N:] = 1.0
                hs.append(out)
                S.append(Sl)
            y = torch.empty(n, WIDTH + 16, **f32)                             # [sdf | feat]: 257 columns in two launches
            if split:
                _gemm16_split(hs[D], _PackedSplit(Ws[D], WIDTH + 1, WIDTH), bs[D], n, WIDTH + 1, WIDTH, y, 0, 0, G_LINEAR)
            else:
                _gemm16(hs[D], Ws[D][:WIDTH], bs[D][:WIDTH], n, WIDTH, WIDTH, y, 0, G_LINEAR)
                _gemm16(hs[D], Ws[D][WIDTH:], bs[D][WIDTH:], n, 1, WIDTH, y[:, WIDTH:], 0, G_LINEAR)
            sdf = y[:, 0]                      # views of y: the consumers (reshape, the radiance net's input rows) take strides
            feat = y[:, 1:WIDTH + 1]
            # ---- reverse sweep: the normal
            Wts = [None] * L
            P = [None] * D
            # split: the normal feeds the radiance net, whose ReLU masks flip on 1e-4 errors of a pre-activation -- the reverse
            # sweep's rows are (hi, lo) pairs too ([n, 512]; everything downstream reads the hi halves through the row stride)
            if split:
                P[D - 1] = torch.empty(n, 2 * WIDTH, **h16)
                _cast_cols16(S[D - 1] * Ws[D][0:1, :WIDTH], 1.0, P[D - 1], lo_off=WIDTH)      # fp16 x fp32 -> fp32 in one kernel
            else:
                P[D - 1] = S[D - 1] * Ws[D][0:1, :WIDTH].half()           # one fp16 kernel (the rows are rounded to fp16 anyway)
            for l in range(D - 1, 0, -1):
                out_l, in_l = dims[l]
                Wt = _pad4c(Ws[l][:, :in_l].t())                                # the reverse sweep and the backprop use these
                Wts[l] = _Packed(Wt, in_l, out_l) if _PACK else Wt
                if split:
                    P[l - 1] = torch.empty(n, 2 * WIDTH, **h16)
                    _gemm16_split(P[l], _PackedSplit(Wt, in_l, out_l), None, n, in_l, out_l, P[l - 1], 1, WIDTH, G_SCALE, aux_a=S[l - 1])
                else:
                    P[l - 1] = torch.empty(n, WIDTH, **h16)
                    _gemm16(P[l], Wts[l], None, n, in_l, out_l, P[l - 1], 1, G_SCALE, aux_a=S[l - 1])
            Wts[0] = _pad4c(Ws[0][:, :pe].t())
            g0 = torch.empty(n, (pe + 15) & ~15, **f32)
            if split:
                _gemm16_split(P[0], _PackedSplit(Wts[0], pe, dims[0][0]), None, n, pe, dims[0][0], g0, 0, 0, G_LINEAR)
            else:
                _gemm16(P[0], Wts[0], None, n, pe, dims[0][0], g0, 0, G_LINEAR)
            nabla = torch.empty(n, 3, **f32)
            ge = P[skip - 1][:, dims[skip - 1][0]:] if skip > 0 else None
            _lib.check(lib.nr_pe_jac_t(_lib.ptr(x), n, multires, _lib.ptr(g0), g0.stride(0), _lib.ptr(ge),
                                       0 if ge is None else ge.stride(0), WIDTH if (split and ge is not None) else 0,
                                       _lib.ptr(nabla), st), "pe_jac_t")
        ctx.state = (x, hs, S, P, Ws, Wp, Wts, dims, multires, skip, pe, n)
        return sdf, nabla, feat

    @staticmethod
    def backward(ctx, g_sdf, g_nabla, g_feat):
        lib = _lib.get_lib()
        x, hs, S, P, Ws, Wp, Wts, dims, multires, skip, pe, n = ctx.state
        L = len(Ws)
        D = L - 1
        dev = x.device
        h16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        scale, inv = _GRAD_SCALE, 1.0 / _GRAD_SCALE
        grads = [None] * (2 * L)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            # ---- adjoint of the reverse sweep (only if the normal has a consumer)
            G, Z2 = [None] * (D + 1), [None] * D
            adj = g_nabla is not None
            if adj:
                G[0] = torch.empty(n, 64, **h16)
                gn = _lib.f32c(g_nabla)
                _lib.check(lib.nr_pe_jac(_lib.ptr(x), n, multires, _lib.ptr(gn), scale, _lib.ptr(G[0]), 64, 64, None, 0, 0, st), "pe_jac")
                for l in range(D):
                    N, K = dims[l]
                    full = N == WIDTH
                    G[l + 1] = torch.empty(n, WIDTH, **h16) if full else torch.zeros(n, WIDTH, **h16)
                    Z2[l] = torch.empty(n, WIDTH, **h16) if full else torch.zeros(n, WIDTH, **h16)
                    _gemm16(G[l], Wp[l], None, n, N, K, G[l + 1], 1, G_ADJ, aux_a=S[l], aux_b=P[l], out2=Z2[l])
                    if l + 1 == skip:
                        G[l + 1][:, N:N + pe] = G[0][:, :pe]
            # ---- yb = scale * [g_sdf | g_feat] as fp16 rows
            yb = torch.zeros(n, 320, **h16)
            if g_sdf is not None:
                _cast_cols16(g_sdf.reshape(n), scale, yb)
            if g_feat is not None:
                _cast_cols16(g_feat, scale, yb[:, 1:])
            # ---- backprop through the forward sweep
            Z = [None] * D
            Wt_out = _pad4c(Ws[D][:, :WIDTH].t())                              # [256, 257 -> 260]
            Z[D - 1] = torch.empty(n, WIDTH, **h16)
            _gemm16(yb, Wt_out, None, n, WIDTH, WIDTH + 1, Z[D - 1], 1, G_SCALE, aux_a=S[D - 1], aux_b=Z2[D - 1])
            for l in range(D - 1, 0, -1):
                out_l, in_l = dims[l]
                Z[l - 1] = torch.empty(n, WIDTH, **h16)
                _gemm16(Z[l], Wts[l], None, n, in_l, out_l, Z[l - 1], 1, G_SCALE, aux_a=S[l - 1], aux_b=Z2[l - 1])
            # ---- weight and bias gradients (accumulated by atomics into ONE zero-filled buffer: a fill per tensor was 36 launches)
            sizes = []
            for l in range(D):
                N, K = dims[l]
                sizes += [N * ((K + 3) & ~3), (N + 3) & ~3]
            sizes += [(WIDTH + 1) * WIDTH, (WIDTH + 1 + 3) & ~3]
            flat = torch.zeros(sum(sizes), **f32)
            segs, off = [], 0
            for sz in sizes:
                segs.append(flat[off:off + sz])
                off += sz
            for l in range(D):
                N, K = dims[l]
                dW = segs[2 * l].view(N, (K + 3) & ~3)
                if adj:      # zb^T h + p^T gb: one launch, one pass of atomics
                    _lib.check(lib.nr_gemm16_tn2(_lib.ptr(Z[l]), WIDTH, _lib.ptr(hs[l]), hs[l].stride(0), _lib.ptr(P[l]), P[l].stride(0),
                                                 _lib.ptr(G[l]), G[l].stride(0), n, N, K, _lib.ptr(dW), dW.stride(0), inv, st), "gemm16_tn2")
                else:
                    _lib.check(lib.nr_gemm16_tn(_lib.ptr(Z[l]), WIDTH, _lib.ptr(hs[l]), hs[l].stride(0), n, N, K, _lib.ptr(dW),
                                                dW.stride(0), inv, st), "gemm16_tn")
                db = segs[2 * l + 1][:N]
                _lib.check(lib.nr_colsum16(_lib.ptr(Z[l]), WIDTH, n, N, inv, _lib.ptr(db), st), "colsum16")
                grads[2 * l], grads[2 * l + 1] = dW[:, :K], db
            dWo = segs[2 * D].view(WIDTH + 1, WIDTH)
            _lib.check(lib.nr_gemm16_tn(_lib.ptr(yb), 320, _lib.ptr(hs[D]), hs[D].stride(0), n, WIDTH + 1, WIDTH, _lib.ptr(dWo), WIDTH, inv, st),
                       "gemm16_tn")
            if adj:
                _lib.check(lib.nr_colsum16(_lib.ptr(G[D]), WIDTH, n, WIDTH, inv, _lib.ptr(dWo), st), "colsum16")   # sdf row
            dbo = segs[2 * D + 1][:WIDTH + 1]                                 # column sums of yb (fp16, scaled): two launches
            _lib.check(lib.nr_colsum16(_lib.ptr(yb), 320, n, WIDTH, inv, _lib.ptr(dbo), st), "colsum16")
            _lib.check(lib.nr_colsum16(_lib.ptr(yb[:, WIDTH:]), 320, n, 1, inv, _lib.ptr(dbo[WIDTH:]), st), "colsum16")
            grads[2 * D], grads[2 * D + 1] = dWo, dbo
        ctx.state = None
        return (None, None, None, *grads)


G_RELU, G_SIGMOID, G_MASK = 4, 5, 6


def _r16(k):
    return (k + 15) & ~15


class RadianceRevFn(torch.autograd.Function):
    """RadianceNet.forward (base.py:372-391) on the 16-bit training GEMMs: cat([PE(x), PE(view), normals, feat]) as fp16
    rows -> ReLU layers -> sigmoid; the backward masks with the saved activations (h > 0) in the GEMM epilogue.
    (x, view, normals [n,3], feat [n,F], W_0, b_0, ...) -> rgb [n,3]; gradients for normals, feat and the weights."""

    @staticmethod
    def forward(ctx, x, view, normals, feat, multires, multires_view, *wb):
        lib = _lib.get_lib()
        dev, n = x.device, x.shape[0]
        L = len(wb) // 2
        Ws = [_pad4c(w) for w in wb[0::2]]
        bs = [b.detach().float().contiguous() for b in wb[1::2]]
        dims = [(w.shape[0], w.shape[1]) for w in wb[0::2]]
        px = 3 if multires < 0 else 3 * (1 + 2 * multires)
        pv = 3 if multires_view < 0 else 3 * (1 + 2 * multires_view)
        in0 = px + pv + 3 + feat.shape[1]
        assert in0 == dims[0][1], "RadianceNet layer 0 width"
        ld0 = (in0 + 63) // 64 * 64
        h16 = dict(dtype=torch.float16, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            split = split_forward()
            a0 = torch.zeros(n, 2 * ld0 if split else ld0, **h16)        # split: [hi (ld0 columns) | lo (ld0 columns)]
            if split:
                _lib.check(lib.nr_pe16_split(_lib.ptr(x), n, multires, _lib.ptr(a0), 2 * ld0, px, ld0, None, 0, 0, 0, st), "pe16_split")
                _lib.check(lib.nr_pe16_split(_lib.ptr(view), n, multires_view, _lib.ptr(a0[:, px:]), 2 * ld0, pv, ld0, None, 0, 0, 0, st),
                           "pe16_split")
                _cast_cols16(normals.detach(), 1.0, a0[:, px + pv:], lo_off=ld0)
                _cast_cols16(feat.detach(), 1.0, a0[:, px + pv + 3:], lo_off=ld0)
            else:
                _lib.check(lib.nr_pe16(_lib.ptr(x), n, multires, _lib.ptr(a0), ld0, px, None, 0, 0, st), "pe16")
                _lib.check(lib.nr_pe16(_lib.ptr(view), n, multires_view, _lib.ptr(a0[:, px:]), ld0, pv, None, 0, 0, st), "pe16")
                a0[:, px + pv:px + pv + 3] = normals
                a0[:, px + pv + 3:in0] = feat
            acts = [a0]
            for l in range(L - 1):
                N, K = dims[l]
                if split:
                    out = torch.empty(n, 2 * WIDTH, **h16)
                    _gemm16_split(acts[l], _PackedSplit(Ws[l], N, K), bs[l], n, N, K, out, 1, WIDTH, G_RELU)
                else:
                    out = torch.empty(n, WIDTH, **h16)
                    _gemm16(acts[l], _Packed(Ws[l], N, K), bs[l], n, N, K, out, 1, G_RELU)
                acts.append(out)
            N, K = dims[L - 1]
            y = torch.empty(n, _r16(N), dtype=torch.float32, device=dev)
            if split:
                _gemm16_split(acts[L - 1], _PackedSplit(Ws[L - 1], N, K), bs[L - 1], n, N, K, y, 0, 0, G_SIGMOID)
            else:
                _gemm16(acts[L - 1], Ws[L - 1], bs[L - 1], n, N, K, y, 0, G_SIGMOID)
            rgb = y[:, :N].contiguous()
        ctx.state = (acts, rgb, Ws, dims, n, px + pv, in0)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        lib = _lib.get_lib()
        acts, rgb, Ws, dims, n, off, in0 = ctx.state
        L = len(Ws)
        dev = rgb.device
        h16 = dict(dtype=torch.float16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        scale, inv = _GRAD_SCALE, 1.0 / _GRAD_SCALE
        grads = [None] * (2 * L)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            N, K = dims[L - 1]
            z = torch.zeros(n, 64, **h16)
            z[:, :N] = g_rgb * rgb * (1.0 - rgb) * scale                       # through the sigmoid
            sizes = []
            for l in range(L):
                N, K = dims[l]
                sizes += [N * ((K + 3) & ~3), (N + 3) & ~3]
            flat = torch.zeros(sum(sizes), **f32)                              # all dW / db of the net: one fill
            segs, o_ = [], 0
            for sz in sizes:
                segs.append(flat[o_:o_ + sz])
                o_ += sz
            for l in range(L - 1, -1, -1):
                N, K = dims[l]
                dW = segs[2 * l].view(N, (K + 3) & ~3)
                _lib.check(lib.nr_gemm16_tn(_lib.ptr(z), z.stride(0), _lib.ptr(acts[l]), acts[l].stride(0), n, N, K, _lib.ptr(dW),
                                            dW.stride(0), inv, st), "gemm16_tn")
                db = segs[2 * l + 1][:N]
                _lib.check(lib.nr_colsum16(_lib.ptr(z), z.stride(0), n, N, inv, _lib.ptr(db), st), "colsum16")
                grads[2 * l], grads[2 * l + 1] = dW[:, :K], db
                Wt = _pad4c(Ws[l][:, :K].t())                                    # [K, N]
                if l > 0:
                    zn = torch.empty(n, WIDTH, **h16)
                    _gemm16(z, Wt, None, n, K, N, zn, 1, G_MASK, aux_a=acts[l])    # relu'(h) = (h > 0)
                    z = zn
                else:                                                            # gradient of the input row: two launches over its columns
                    ga = torch.empty(n, _r16(K - WIDTH) + WIDTH, **f32)
                    _gemm16(z, Wt[:WIDTH], None, n, WIDTH, N, ga, 0, G_LINEAR)
                    _gemm16(z, Wt[WIDTH:], None, n, K - WIDTH, N, ga[:, WIDTH:], 0, G_LINEAR)
        g_normals = ga[:, off:off + 3] * inv
        g_feat = ga[:, off + 3:in0] * inv
        ctx.state = None
        return (None, None, g_normals, g_feat, None, None, *grads)
