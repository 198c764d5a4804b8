"""Host-side packing of MLP weights into the bf16, pre-swizzled shared-memory images the
tcgen05 kernels stream with bulk async copies (layout: csrc/umma.cuh)."""
import numpy as np
import torch

_A_IDX = None


def _a_tile_index():
    """uint16 index of element (r, k) inside a [128 x 64] K-major SWIZZLE_128B tile."""
    global _A_IDX
    if _A_IDX is None:
        r = np.arange(128)[:, None]
        k = np.arange(64)[None, :]
        off = r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1)
        _A_IDX = torch.from_numpy((off // 2).astype(np.int64))
    return _A_IDX


_A_IDX_DEV = {}


def _a_tile_index_on(device):
    """The tile index on `device`, uploaded once (the weights are re-packed after every optimiser step; a host-to-device
    copy per call would also break CUDA-graph capture of a training step)."""
    key = str(device)
    if key not in _A_IDX_DEV:
        _A_IDX_DEV[key] = _a_tile_index().to(device).reshape(-1)
    return _A_IDX_DEV[key]


def _pack_a_device(W, n_mt, kp, mode):
    """nr_umma_pack_a: one launch, any strides (csrc/umma_pack.cu)"""
    from . import _lib
    lib = _lib.get_lib()
    Wf = W.detach()
    if Wf.dtype != torch.float32:
        Wf = Wf.float()
    rows, K = Wf.shape
    kp64 = (kp + 63) // 64 * 64
    img = torch.empty(int(lib.nr_umma_pack_a_bytes(n_mt, kp64, mode)) // 2, dtype=torch.int16, device=Wf.device)
    with torch.cuda.device(Wf.device):
        _lib.check(lib.nr_umma_pack_a(_lib.ptr(Wf), Wf.stride(0), Wf.stride(1), rows, K, n_mt, kp64, mode, _lib.ptr(img),
                                      _lib.stream_ptr(Wf.device)), "umma_pack_a")
    return img.view(-1, 8192)


def pack_a_tiles(W, n_mtiles=None, k_pad=None, dtype=torch.bfloat16):
    """W: [rows, K] float tensor -> int16 tensor [n_kchunks * n_mtiles, 8192] of A tiles ordered
    (k-chunk major, M-tile minor: the order the two MMA-issuing warps consume them);
    rows/K zero-padded to 128 / 64 multiples.  CUDA tensors are packed by nr_umma_pack_a; the torch composition below
    remains for CPU tensors (layout tests)."""
    rows, K = W.shape
    n_mt = n_mtiles if n_mtiles is not None else (rows + 127) // 128
    kp = k_pad if k_pad is not None else K
    if W.is_cuda:
        return _pack_a_device(W, n_mt, kp, 0 if dtype == torch.float16 else 1)
    n_kc = (kp + 63) // 64
    full = torch.zeros(n_mt * 128, n_kc * 64, dtype=torch.float32, device=W.device)
    full[:rows, :K] = W.float()
    bf = full.to(dtype).view(torch.int16)
    tiles = bf.reshape(n_mt, 128, n_kc, 64).permute(2, 0, 1, 3).reshape(n_kc * n_mt, 128 * 64)
    idx = _a_tile_index_on(W.device)
    out = torch.empty_like(tiles)
    out[:, idx] = tiles
    return out.contiguous()


def pack_a_tiles_split(W, n_mtiles=None, k_pad=None):
    """Split-precision image (csrc/mlp_rev_split.cu): W = hi + 2^-12 lo with hi = fp16(W), lo = fp16((W - hi) 2^12) (scaled so
    that it is a normal fp16 number); the A tiles of both parts interleaved (hi, lo) per (k-chunk, M-tile):
    [2 * n_kchunks * n_mtiles, 8192] int16."""
    if W.is_cuda:
        rows, K = W.shape
        return _pack_a_device(W, n_mtiles if n_mtiles is not None else (rows + 127) // 128, k_pad if k_pad is not None else K, 2)
    Wf = W.float()
    hi = Wf.to(torch.float16)
    lo = ((Wf - hi.float()) * 4096.0).to(torch.float16)
    th = pack_a_tiles(hi, n_mtiles=n_mtiles, k_pad=k_pad, dtype=torch.float16)
    tl = pack_a_tiles(lo, n_mtiles=n_mtiles, k_pad=k_pad, dtype=torch.float16)
    return torch.stack([th, tl], dim=1).reshape(-1, th.shape[1]).contiguous()


# --------------------------------------------------------------------------------------------
# Programs for the fused tcgen05 MLP kernel (csrc/mlp_umma.cu, include/neurecon_b200.h)
# --------------------------------------------------------------------------------------------
EPI_HIDDEN, EPI_SDF_OUT, EPI_FEAT, EPI_RELU, EPI_RGB, EPI_EXTRAS = 0, 1, 2, 3, 4, 5
EPI_BWD, EPI_NABLA = 7, 8


def _pe_dim(multires):
    return 3 if multires < 0 else 3 + 6 * multires


class UmmaNet:
    """bf16 weight image + bias table + step templates for one (surface[, radiance]) pair."""

    def __init__(self, surface_W, surface_b, multires, skip_layer, rad_W=None, rad_b=None, rad_multires=-1,
                 rad_multires_view=-1, operand="fp16", split=False):
        dev = surface_W[0].device
        self.operand = operand
        # split: every weight chunk as a (hi, lo) fp16 pair for csrc/mlp_rev_split.cu (reverse-mode programs only)
        self.split = bool(split)
        assert not (split and (rad_W is not None or operand != "fp16")), "split images hold the SDF net only, fp16 parts"
        op_dtype = torch.float16 if operand == "fp16" else torch.bfloat16
        self.multires = multires
        self.rad_multires, self.rad_multires_view = rad_multires, rad_multires_view
        chunks, biases = [], []
        self._n_chunks, self._n_bias = 0, 0

        def add(W, b, k_steps, n_mt):
            assert k_steps % 4 == 0
            if self.split:
                img = pack_a_tiles_split(W, n_mtiles=n_mt, k_pad=k_steps * 16)
                assert img.shape[0] == 2 * n_mt * (k_steps // 4)
            else:
                img = pack_a_tiles(W, n_mtiles=n_mt, k_pad=k_steps * 16, dtype=op_dtype)
                assert img.shape[0] == n_mt * (k_steps // 4)
            bt = torch.zeros(n_mt * 128, dtype=torch.float32, device=dev)
            if b is not None:
                bt[: b.numel()] = b.float()
            c0, b0 = self._n_chunks, self._n_bias
            chunks.append(img)
            biases.append(bt)
            self._n_chunks += img.shape[0]
            self._n_bias += bt.numel()
            return c0, b0

        L = len(surface_W)
        pe = _pe_dim(multires)
        if pe > 64:
            raise NotImplementedError("tensor tier: embedding wider than 64 is not supported (use set_precision('fp32'))")
        if pe > 40 and 0 < skip_layer < L - 1:
            raise NotImplementedError("tensor tier: a skip connection needs an embedding of at most 40 rows "
                                      "(embed_multires <= 6); use set_precision('fp32') for this network")
        self.hidden = []
        for l in range(L - 1):
            W, b = surface_W[l], surface_b[l]
            out_d, in_d = W.shape
            if in_d > 256 or out_d > 256:
                raise NotImplementedError("bf16 tier supports hidden widths up to 256")
            k_steps, n_mt = (in_d + 63) // 64 * 4, (out_d + 127) // 128
            c0, b0 = add(W, b, k_steps, n_mt)
            pe_fill = 1 if (l + 1 == skip_layer) else 0
            if pe_fill and out_d + pe > 256:
                raise NotImplementedError("bf16 tier: skip operand wider than 256")
            self.hidden.append(dict(chunk_begin=c0, n_mt=n_mt, k_steps=k_steps, epi=EPI_HIDDEN, bias_off=b0,
                                    out_rows=out_d, pe_fill=pe_fill, to_rad=0))
        Wl, bl = surface_W[L - 1], surface_b[L - 1]
        in_d = Wl.shape[1]
        k_steps = (in_d + 63) // 64 * 4
        c0, b0 = add(Wl[0:1].expand(32, in_d), bl[0:1].expand(32), k_steps, 1)
        self.sdf_out = dict(chunk_begin=c0, n_mt=1, k_steps=k_steps, epi=EPI_SDF_OUT, bias_off=b0, out_rows=1,
                            pe_fill=0, to_rad=0)
        # CTA-pair kernel: the replicated sdf row in BOTH M-tiles, so each CTA of a pair finds it in its own TMEM lanes
        rep = torch.zeros(256, in_d, dtype=Wl.dtype, device=dev)
        rep[0:32] = Wl[0:1]
        rep[128:160] = Wl[0:1]
        brep = torch.zeros(256, dtype=bl.dtype, device=dev)
        brep[0:32] = bl[0]
        brep[128:160] = bl[0]
        c0, b0 = add(rep, brep, k_steps, 2)
        self.sdf_out2 = dict(chunk_begin=c0, n_mt=2, k_steps=k_steps, epi=EPI_SDF_OUT, bias_off=b0, out_rows=1,
                             pe_fill=0, to_rad=0)
        # reverse-mode normals (csrc/mlp_rev.cu): the sdf row's weights as an fp32 table for the start of the backward
        # sweep, W_l^T chunks for l = L-2 .. 1 (rows = the layer's inputs, k = its outputs) and W_0^T for the embedding
        aux = torch.zeros(256, dtype=torch.float32, device=dev)
        aux[:in_d] = Wl[0].float()
        self._aux_off = self._n_bias
        biases.append(aux)
        self._n_bias += 256
        self.bwd = []
        for l in range(L - 2, 0, -1):
            Wt = surface_W[l].t()                      # [in_d, out_d]
            m_d, k_d = Wt.shape
            k_steps, n_mt = (k_d + 63) // 64 * 4, (m_d + 127) // 128
            c0, _ = add(Wt, None, k_steps, n_mt)
            skip = 1 if l == skip_layer else 0
            self.bwd.append(dict(chunk_begin=c0, n_mt=n_mt, k_steps=k_steps, epi=EPI_BWD, bias_off=0,
                                 out_rows=surface_W[l - 1].shape[0], pe_fill=skip, to_rad=0, sig_slot=l - 1))
        W0t = surface_W[0].t()
        c0, _ = add(W0t, None, (W0t.shape[1] + 63) // 64 * 4, 1)
        self.bwd.append(dict(chunk_begin=c0, n_mt=1, k_steps=(W0t.shape[1] + 63) // 64 * 4, epi=EPI_NABLA, bias_off=0,
                             out_rows=pe, pe_fill=1 if 0 < skip_layer < L - 1 else 0, to_rad=0))
        self.feat_dim = Wl.shape[0] - 1
        if self.feat_dim > 0:
            if self.feat_dim > 256:
                raise NotImplementedError("bf16 tier: geometry feature wider than 256")
            n_mt = (self.feat_dim + 127) // 128
            c0, b0 = add(Wl[1:], bl[1:], k_steps, n_mt)
            self.feat = dict(chunk_begin=c0, n_mt=n_mt, k_steps=k_steps, epi=EPI_FEAT, bias_off=b0,
                             out_rows=self.feat_dim, pe_fill=0, to_rad=0)
        self.rad = None
        self.rad_extra_rows = 0
        if rad_W is not None:
            if self.feat_dim != 256:
                raise NotImplementedError("bf16 tier: the fused radiance path needs W_geo_feat == 256")
            px, pv = _pe_dim(rad_multires), _pe_dim(rad_multires_view)
            n_extra = px + pv + 3
            self.rad_extra_rows = (n_extra + 63) // 64 * 64
            if 256 + self.rad_extra_rows > 512:
                raise NotImplementedError("bf16 tier: radiance input too wide")
            W0 = rad_W[0]
            assert W0.shape[1] == n_extra + 256, "RadianceNet layer 0 width"
            # reference order [PE(x) | PE(view) | normals | feat] (base.py:382) -> operand order [feat | ... ]
            W0p = torch.cat([W0[:, n_extra:], W0[:, :n_extra]], dim=1)
            steps = []
            Ls = len(rad_W)
            for l in range(Ls):
                W = W0p if l == 0 else rad_W[l]
                out_d, in_d = W.shape
                if l > 0 and in_d > 256 or out_d > 256:
                    raise NotImplementedError("bf16 tier supports radiance widths up to 256 (no skips)")
                k_steps = (256 + self.rad_extra_rows) // 16 if l == 0 else (in_d + 63) // 64 * 4
                n_mt = (out_d + 127) // 128
                c0, b0 = add(W, rad_b[l], k_steps, n_mt)
                steps.append(dict(chunk_begin=c0, n_mt=n_mt, k_steps=k_steps, epi=EPI_RGB if l == Ls - 1 else EPI_RELU,
                                  bias_off=b0, out_rows=out_d, pe_fill=0, to_rad=0))
            if steps[-1]["out_rows"] != 3:
                raise NotImplementedError("bf16 tier: radiance output must be rgb")
            self.rad = steps
            # the geometry-feature layer is linear and feeds only radiance layer 0: folded into it for the radiance pass
            # that starts from the last hidden SDF activations (program 'radiance'):
            #   W0[:, feat] (W_f h + b_f) = (W0[:, feat] W_f) h + W0[:, feat] b_f      -- one K = 256 layer less per point
            W0f = W0[:, n_extra:].double()
            Wfold = (W0f @ Wl[1:].double()).float()
            bfold = (rad_b[0].double() + W0f @ bl[1:].double()).float()
            kf = (Wfold.shape[1] + 63) // 64 * 4
            c0, b0 = add(Wfold, bfold, kf, (Wfold.shape[0] + 127) // 128)
            self.rad_fold = dict(steps[0], chunk_begin=c0, bias_off=b0, k_steps=kf)
        self.image = torch.cat(chunks, 0).contiguous()
        self.bias = torch.cat(biases, 0).contiguous()

    def rev_ok(self, want_feat=False):
        """The reverse-mode program (csrc/mlp_rev.cu) keeps the embedding rows in the 40-row stash and has 2 L + 1 (+ 1)
        steps; anything else runs on the forward-mode tangent tiles."""
        from . import _lib
        pe = _pe_dim(self.multires)
        n_steps = 2 * len(self.hidden) + 1 + (1 if want_feat else 0)
        return pe <= 40 and pe % 3 == 0 and n_steps <= _lib.NR_UMMA_MAX_STEPS

    def pair_ok(self):
        """The CTA-pair kernel needs every step as an M-tile pair (hidden width > 128) and the feature, if any, too."""
        return all(h["n_mt"] == 2 for h in self.hidden) and (self.feat_dim == 0 or self.feat["n_mt"] == 2)

    def program(self, mode, want_feat=False, pair=False):
        """mode: 'sdf' (128-point tiles), 'nablas' (32-point tangent tiles; want_feat: fp32 feature rows or, for
        'nablas_img', the radiance operand image), 'radiance' (the radiance net alone on 128-point tiles fed by that
        image: layer 0 = feature part, then the [PE(x)|PE(view)|normals] part accumulated onto it),
        'fused' (everything in one launch, radiance steps 32 columns wide)."""
        from . import _lib
        if mode in ("radiance", "radiancef"):
            # 'radiance': the image holds the last hidden SDF activations -> the geometry-feature layer (linear) first;
            # 'radiancef': the image holds the feature itself (written by an EPI_FEAT step, 'nablas_imgf')
            assert self.rad is not None
            r0 = self.rad[0]
            extras = dict(r0, chunk_begin=r0["chunk_begin"] + r0["n_mt"] * 4, k_steps=self.rad_extra_rows // 16,
                          accumulate=1, n_cols=128)
            if mode == "radiance":     # image = last hidden activations: feature layer folded into layer 0
                steps = [dict(self.rad_fold, epi=EPI_EXTRAS, n_cols=128), dict(extras, bias_off=self.rad_fold["bias_off"])]
            else:
                steps = [dict(r0, k_steps=16, epi=EPI_EXTRAS, n_cols=128), extras]
            steps += [dict(s, n_cols=128) for s in self.rad[1:]]
            return self._finish(steps, tang=0, input_mode=1)
        if mode == "rev_sdf":
            # split-precision kernel, sdf (and optionally the feature) only: the forward sweep of 'rev'
            steps = [dict(s, n_cols=128, sig_slot=i) for i, s in enumerate(self.hidden)]
            if want_feat:
                steps.append(dict(self.feat, n_cols=128))
            steps.append(dict(self.sdf_out, n_cols=128, sig_slot=len(self.hidden) - 1, aux_off=self._aux_off))
            return self._finish(steps, tang=0, reverse=1)
        assert not self.split or mode in ("rev", "rev_img"), "split images run reverse-mode programs only"
        if mode in ("rev", "rev_img"):
            # reverse-mode normals on 128-point tiles (nr_mlp_umma_reverse); 'rev_img': the last hidden activations
            # also go to the radiance pass's operand image
            steps = [dict(s, n_cols=128, sig_slot=i) for i, s in enumerate(self.hidden)]
            if mode == "rev_img":
                steps[-1]["to_rad"] = 1
            elif want_feat:
                steps.append(dict(self.feat, n_cols=128))
            steps.append(dict(self.sdf_out, n_cols=128, sig_slot=len(self.hidden) - 1, aux_off=self._aux_off))
            steps += [dict(s, n_cols=128) for s in self.bwd]
            return self._finish(steps, tang=0, reverse=1)
        tang = 0 if mode == "sdf" else 1
        steps = [dict(s, n_cols=128) for s in self.hidden]
        steps.append(dict(self.sdf_out2 if pair else self.sdf_out, n_cols=128))
        if mode == "nablas_img":       # the last hidden layer's value activations also go to the operand image
            steps[len(self.hidden) - 1]["to_rad"] = 1
        elif mode == "nablas_imgf":    # the geometry feature (one more, 32-column step) goes to the operand image
            steps.append(dict(self.feat, n_cols=32))
        elif mode == "fused":
            assert self.rad is not None
            steps.append(dict(self.feat, n_cols=32, to_rad=1))
            steps += [dict(s, n_cols=32) for s in self.rad]
        elif want_feat:
            steps.append(dict(self.feat, n_cols=32 if tang else 128))
        return self._finish(steps, tang)

    def _finish(self, steps, tang, input_mode=0, reverse=0):
        from . import _lib
        P = _lib.UmmaProgram()
        assert len(steps) <= _lib.NR_UMMA_MAX_STEPS
        P.n_steps, P.tangents, P.multires, P.input_mode = len(steps), tang, self.multires, input_mode
        P.reverse = reverse
        P.rad_multires, P.rad_multires_view, P.rad_extra_rows = self.rad_multires, self.rad_multires_view, self.rad_extra_rows
        P.operand_f16 = 1 if self.operand == "fp16" else 0
        for i, s in enumerate(steps):
            for k, v in s.items():
                setattr(P.steps[i], k, int(v))
        return P


EPI_LINEAR = 6


class UmmaNerfNet:
    """16-bit weight image + bias table + program of the NeRF++ background net (models/base.py:395-453, the
    ``use_view_dirs=True`` form) for the fused tcgen05 kernel, 128-point value tiles (input_mode 2).

    Layers whose input is a concatenation run as split-K pairs of steps over two operands that live in the same rows of
    the shared-memory buffer one after the other: ``pts_linears[s+1](cat([PE(x), h]))`` = W[:, npe:] h (then the rows
    are refilled with PE(x)) + W[:, :npe] PE(x); ``views_linears[0](cat([feature, PE(view)]))`` likewise."""

    def __init__(self, module, operand="fp16"):
        dev = module.alpha_linear.weight.device
        self.operand = operand
        op_dtype = torch.float16 if operand == "fp16" else torch.bfloat16
        if not module.use_view_dirs:
            raise NotImplementedError("tensor tier: NeRF needs use_view_dirs=True")
        W, D = module.W, module.D
        if W != 256 or len(module.skips) > 1:
            raise NotImplementedError("tensor tier: NeRF++ net must be 256 wide with at most one skip")
        self.multires, self.multires_view, self.input_dim = module.multires, module.multires_view, module.input_dim
        npe = self.input_dim * (1 if self.multires < 0 else 1 + 2 * self.multires)
        npv = 3 if self.multires_view < 0 else 3 + 6 * self.multires_view
        if npe > 256 or npv > 256:
            raise NotImplementedError("tensor tier: embedding wider than 256")
        chunks, biases, steps = [], [], []
        self._n_chunks = self._n_bias = 0

        def add(Wm, b, epi, n_mt=None, **kw):
            out_d, in_d = Wm.shape
            k_steps = (in_d + 63) // 64 * 4
            n_mt = (out_d + 127) // 128 if n_mt is None else n_mt
            img = pack_a_tiles(Wm, n_mtiles=n_mt, k_pad=k_steps * 16, dtype=op_dtype)
            bt = torch.zeros(n_mt * 128, dtype=torch.float32, device=dev)
            if b is not None:
                bt[: b.numel()] = b.float()
            steps.append(dict(chunk_begin=self._n_chunks, n_mt=n_mt, k_steps=k_steps, n_cols=128, epi=epi,
                              bias_off=self._n_bias, out_rows=out_d, pe_fill=0, to_rad=0, accumulate=0, **kw))
            chunks.append(img)
            biases.append(bt)
            self._n_chunks += img.shape[0]
            self._n_bias += bt.numel()

        lin = [(l.weight.detach().float(), l.bias.detach().float()) for l in module.pts_linears]
        for i in range(D):
            Wi, bi = lin[i]
            if i > 0 and (i - 1) in module.skips:      # input = cat([PE(x), h])  (base.py:436)
                add(Wi[:, npe:], None, EPI_EXTRAS)
                steps[-1]["to_rad"] = 1
                add(Wi[:, :npe], bi, EPI_RELU)
                steps[-1]["accumulate"] = 1
            else:
                add(Wi, bi, EPI_RELU)
        if (D - 1) in module.skips:
            raise NotImplementedError("tensor tier: skip after the last NeRF++ layer")
        a = module.alpha_linear
        add(a.weight.detach().float().expand(32, W), a.bias.detach().float().expand(32), EPI_SDF_OUT, n_mt=1)
        f = module.feature_linear
        add(f.weight.detach().float(), f.bias.detach().float(), EPI_LINEAR)
        v = module.views_linears[0]
        Wv, bv = v.weight.detach().float(), v.bias.detach().float()
        add(Wv[:, :W], None, EPI_EXTRAS)               # cat([feature, PE(view)])  (base.py:441)
        steps[-1]["to_rad"] = 2
        add(Wv[:, W:], bv, EPI_RELU)
        steps[-1]["accumulate"] = 1
        r = module.rgb_linear
        add(r.weight.detach().float(), r.bias.detach().float(), EPI_RGB)
        self.steps = steps
        self.image = torch.cat(chunks, 0).contiguous()
        self.bias = torch.cat(biases, 0).contiguous()

    def program(self):
        from . import _lib
        P = _lib.UmmaProgram()
        assert len(self.steps) <= _lib.NR_UMMA_MAX_STEPS
        P.n_steps, P.tangents, P.multires, P.input_mode, P.input_dim = len(self.steps), 0, self.multires, 2, self.input_dim
        P.rad_multires, P.rad_multires_view, P.rad_extra_rows = -1, self.multires_view, 0
        P.operand_f16 = 1 if self.operand == "fp16" else 0
        for i, s in enumerate(self.steps):
            for k, v in s.items():
                setattr(P.steps[i], k, int(v))
        return P
