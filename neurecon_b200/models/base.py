"""Drop-in for the hot-path part of the reference's ``models/base.py``.

Same class names, constructor signatures, ``state_dict`` keys/shapes and call signatures as
the reference (SuwoongHeo/neurecon ``models/base.py``):

* ``Embedder`` / ``get_embedder``   -- base.py:14-81
* ``DenseLayer``                    -- base.py:118-129
* ``ImplicitSurface``               -- base.py:131-282 (``forward``, ``forward_with_nablas``)
* ``RadianceNet``                   -- base.py:312-391
* ``NeRF`` (NeRF++ background)      -- base.py:395-453

Parameters stay ordinary ``nn.Parameter``s (``*.weight_g / *.weight_v / *.bias``) so reference
checkpoints load unchanged; the arithmetic runs in the sm_100a kernels behind the C-ABI.
The SIREN variant (``use_siren=True``) is outside the hot-path scope.
"""
import math
import warnings

import numpy as np
import os

import torch
import torch.nn as nn

from .. import _lib
from .._lib import C


# --------------------------------------------------------------------------------------------
# Embedder (base.py:14-81).  Stand-alone use only; inside the MLP kernels the encoding is fused.
# --------------------------------------------------------------------------------------------
class Embedder(nn.Module):
    def __init__(self, input_dim, max_freq_log2, N_freqs, log_sampling=True, include_input=True,
                 periodic_fns=(torch.sin, torch.cos)):
        super().__init__()
        self.input_dim = input_dim
        self.include_input = include_input
        self.periodic_fns = periodic_fns
        self.out_dim = (input_dim if include_input else 0) + input_dim * N_freqs * len(periodic_fns)
        if log_sampling:
            bands = 2.0 ** torch.linspace(0.0, max_freq_log2, N_freqs)
        else:
            bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq_log2, N_freqs)
        self.freq_bands = bands.numpy().tolist()

    def forward(self, input):
        assert input.shape[-1] == self.input_dim
        parts = [input] if self.include_input else []
        for f in self.freq_bands:
            for fn in self.periodic_fns:
                parts.append(fn(input * f))
        return torch.cat(parts, dim=-1)


def get_embedder(multires, input_dim=3):
    if multires < 0:
        return nn.Identity(), input_dim
    emb = Embedder(input_dim=input_dim, max_freq_log2=multires - 1, N_freqs=multires,
                   log_sampling=True, include_input=True, periodic_fns=[torch.sin, torch.cos])
    return emb, emb.out_dim


class DenseLayer(nn.Linear):
    """nn.Linear + activation (base.py:118-129).  Only the parameter container and the
    activation tag matter here; the product is evaluated by the fused kernels."""

    def __init__(self, input_dim, out_dim, *args, activation=None, **kwargs):
        super().__init__(input_dim, out_dim, *args, **kwargs)
        self.activation = nn.ReLU(inplace=True) if activation is None else activation

    def forward(self, x):
        return self.activation(super().forward(x))


def _weight_norm(layer):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return nn.utils.weight_norm(layer)


def _effective_weight(layer):
    """W = g * v / ||v||_row for weight-normed layers (old-style weight_norm, dim=0), else .weight.
    Evaluated in PyTorch so that autograd owns g and v."""
    if hasattr(layer, "weight_g"):
        v, g = layer.weight_v, layer.weight_g
        return v * (g / v.norm(dim=1, keepdim=True))
    return layer.weight


def _effective_weights_nograd(layers, scales):
    """effective weights of all layers (detached fp32, scale folded in): ONE launch of nr_weight_norm for weight-normed CUDA
    layers (the images are re-packed after every optimiser step), else the torch composition"""
    with torch.no_grad():
        from .autograd import _fused_effective_weights
        Ws = _fused_effective_weights(layers, scales)       # None unless every layer is weight-normed and on a CUDA device
        if Ws is not None:
            return [W.detach() for W in Ws]
        out = []
        for l, sc in zip(layers, scales):
            W = _effective_weight(l).detach().float()
            out.append(W * sc if sc != 1.0 else W)
        return out


def _param_key(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


def _cached(module, slot, key, build):
    """Packed weights / descriptors of a module, rebuilt when a parameter changed.  ``nn.DataParallel`` replicas are
    shallow copies that SHARE the module's ``_cache`` dict and run in concurrent threads on different devices
    (the reference renders that way, neus.py:413-414): entries are keyed by (slot, device), built into a local and
    stored with one dict assignment, so a replica can neither see a half-built entry nor another device's pointers."""
    dev = next(module.parameters()).device
    k = (slot, dev.type, dev.index)
    hit = module._cache.get(k)
    if hit is None or hit[0] != key:
        hit = (key, build())
        module._cache[k] = hit
    return hit[1]


def _pad4(k):
    return (k + 3) & ~3


def _pack_layers(weights, biases, scales=None):
    """fp32 [out, pad4(in)] zero-padded copies + descriptor arrays."""
    Ws, bs = [], []
    for i, (W, b) in enumerate(zip(weights, biases)):
        W = W.detach().float()
        if scales is not None and scales[i] != 1.0:
            W = W * scales[i]
        out_d, in_d = W.shape
        Wp = torch.zeros(out_d, _pad4(in_d), dtype=torch.float32, device=W.device)
        Wp[:, :in_d] = W
        Ws.append(Wp.contiguous())
        bs.append(b.detach().float().contiguous())
    return Ws, bs


_MAX_POINTS_PER_CALL = 1 << 18
_SPLIT_POINTS = 1 << 21          # points per (geometry, radiance) launch pair of the tensor tier: 1 GiB feature image
_FUSED_MODE = os.environ.get("NEURECON_B200_FUSED", "split")   # 'split' | 'fused' (single launch)
# SDF net on CTA pairs (tcgen05.mma.cta_group::2, csrc/mlp_umma2.cu): numerically identical, but measured slower on
# B200 (DSMEM activation exchange at ~20 B/clk/SM, weight round trip through a relay) -- opt-in until that is fixed
_PAIR_KERNEL = os.environ.get("NEURECON_B200_PAIR", "0") != "0"
# Normals of the tensor tier: reverse mode (csrc/mlp_rev.cu: forward + backward sweep on 128-point tiles, half the
# tensor work) or, with NEURECON_B200_NABLAS=forward, the forward-mode tangent tiles of csrc/mlp_umma.cu
_SDF_VIA_REV = os.environ.get("NEURECON_B200_SDF_VIA_REV", "1") != "0"
_REVERSE_NABLAS = os.environ.get("NEURECON_B200_NABLAS", "reverse") != "forward"


class ImplicitSurface(nn.Module):
    """Geometry MLP (base.py:131-282): PE -> D softplus(beta=100) layers with a skip
    ``cat([h, pe])/sqrt(2)`` -> linear(1 + W_geo_feat)."""

    def __init__(self, W=256, D=8, skips=[4], W_geo_feat=256, input_ch=3, radius_init=1.0,
                 obj_bounding_size=2.0, geometric_init=True, embed_multires=6, weight_norm=True,
                 use_siren=False):
        super().__init__()
        if use_siren:
            raise NotImplementedError("SIREN surfaces are outside the neurecon_b200 hot-path scope")
        self.radius_init = radius_init
        self.register_buffer("obj_bounding_size", torch.tensor([obj_bounding_size]).float())
        self.geometric_init = geometric_init
        self.D, self.W, self.W_geo_feat = D, W, W_geo_feat
        self.skips = list(skips)
        self.use_siren = False
        self.embed_multires = embed_multires
        self.embed_fn, input_ch = get_embedder(embed_multires)
        self._pe_dim = input_ch

        layers = []
        for l in range(D + 1):
            if l == D:
                out_dim = 1 + W_geo_feat if W_geo_feat > 0 else 1
            elif (l + 1) in self.skips:
                out_dim = W - input_ch
            else:
                out_dim = W
            in_dim = input_ch if l == 0 else W
            if l != D:
                layer = DenseLayer(in_dim, out_dim, activation=nn.Softplus(beta=100))
            else:
                layer = nn.Linear(in_dim, out_dim)
            if geometric_init:
                self._sphere_init(layer, l, in_dim, out_dim, input_ch)
            if weight_norm:
                layer = _weight_norm(layer)
            layers.append(layer)
        self.surface_fc_layers = nn.ModuleList(layers)
        self._cache = {}

    def _sphere_init(self, layer, l, in_dim, out_dim, input_ch):
        """SAL / IDR sphere initialisation (same distributions as base.py:207-224)."""
        with torch.no_grad():
            if l == self.D:
                layer.weight.normal_(mean=math.sqrt(math.pi) / math.sqrt(in_dim), std=0.0001)
                layer.bias.fill_(-self.radius_init)
            elif self.embed_multires > 0 and l == 0:
                layer.bias.zero_()
                layer.weight[:, 3:].zero_()
                layer.weight[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(out_dim))
            elif self.embed_multires > 0 and l in self.skips:
                layer.bias.zero_()
                layer.weight.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim))
                layer.weight[:, -(input_ch - 3):].zero_()
            else:
                layer.bias.zero_()
                layer.weight.normal_(0.0, math.sqrt(2) / math.sqrt(out_dim))

    def pretrain_hook(self, configs={}):
        configs["target_radius"] = self.radius_init
        configs["obj_bounding_size"] = self.obj_bounding_size.item()
        return False

    # ---- packed weights / descriptor -------------------------------------------------------
    def _check_supported(self):
        if len(self.skips) > 1:
            raise NotImplementedError("neurecon_b200 supports at most one skip connection")
        if self.W_geo_feat <= 0:
            raise NotImplementedError("neurecon_b200 supports the IDR-style geometry feature (W_geo_feat > 0)")
        for s in self.skips:
            if not (0 < s < self.D):
                raise NotImplementedError("skip layer index must be in (0, D)")

    def _descriptor(self):
        """Device-resident packed weights + C descriptor, rebuilt only when a parameter changed."""
        self._check_supported()
        def build():
            Wl = [_effective_weight(l) for l in self.surface_fc_layers]
            bl = [l.bias for l in self.surface_fc_layers]
            scales = [1.0 / math.sqrt(2) if i in self.skips else 1.0 for i in range(self.D + 1)]
            Ws, bs = _pack_layers(Wl, bl, scales)
            d = _lib.SdfNet()
            d.n_layers = self.D + 1
            d.multires = self.embed_multires
            d.skip_layer = self.skips[0] if self.skips else -1
            d.width = self.W
            for i, l in enumerate(self.surface_fc_layers):
                out_d, in_d = Wl[i].shape
                d.in_dim[i], d.out_dim[i] = in_d, out_d
                d.W[i], d.b[i] = Ws[i].data_ptr(), bs[i].data_ptr()
            d.umma_image = None
            d.umma_bias = None
            return d, (Ws, bs)            # the descriptor holds raw pointers: keep the tensors alive with it
        return _cached(self, "desc", _param_key(self), build)[0]

    def _umma_net(self, radiance_net=None):
        """bf16 image / bias table / step templates for the tcgen05 tier (cached like _descriptor)."""
        self._check_supported()
        from .. import umma_pack
        key = (_param_key(self), None if radiance_net is None else _param_key(radiance_net), _lib.get_precision())
        slot = "umma_fused" if radiance_net is not None else "umma"

        def build():
            with torch.no_grad():
                layers = list(self.surface_fc_layers)
                Wl = _effective_weights_nograd(layers, [1.0 / math.sqrt(2) if i in self.skips else 1.0 for i in range(len(layers))])
                bl = [l.bias.detach().float() for l in self.surface_fc_layers]
                kw = {}
                if radiance_net is not None:
                    if radiance_net.skips:
                        raise NotImplementedError("tensor tier: RadianceNet needs skips=[]")
                    rl = list(radiance_net.layers)
                    rW = _effective_weights_nograd(rl, [1.0] * len(rl))
                    rW[0] = radiance_net._layer0_weight(rW[0])
                    kw = dict(rad_W=rW,
                              rad_b=[l.bias.detach().float() for l in radiance_net.layers],
                              rad_multires=radiance_net.embed_multires,
                              rad_multires_view=radiance_net._multires_view_eff)
                return umma_pack.UmmaNet(Wl, bl, self.embed_multires, self.skips[0] if self.skips else -1,
                                         operand=_lib.operand(), **kw)
        return _cached(self, slot, key, build)

    def _umma_split_net(self):
        """(hi, lo) fp16 weight image of the SDF net for the split-precision kernel (precision 'fp16x2')."""
        self._check_supported()
        from .. import umma_pack

        def build():
            with torch.no_grad():
                layers = list(self.surface_fc_layers)
                Wl = _effective_weights_nograd(layers, [1.0 / math.sqrt(2) if i in self.skips else 1.0 for i in range(len(layers))])
                bl = [l.bias.detach().float() for l in self.surface_fc_layers]
                net = umma_pack.UmmaNet(Wl, bl, self.embed_multires, self.skips[0] if self.skips else -1, operand="fp16",
                                        split=True)
            if not net.rev_ok(want_feat=True):
                raise NotImplementedError("precision 'fp16x2' needs an embedding of at most 40 rows (embed_multires <= 6) and "
                                          "at most 11 hidden layers; use set_precision('fp32') for this network")
            return net
        return _cached(self, "umma_split", _param_key(self), build)

    def forward_level_set(self, x):
        """sdf only, for iterations that walk onto the zero level set (ray_casting.sphere_tracing_surface_points): always
        mlp_umma_kernel's 'sdf' program in the fp16 tier.  The forward sweep of mlp_rev_kernel that `forward` uses is as close
        to the fp32 reference (tools/check_rev_err.py; surface normals of 4096 sphere-traced rays: rms 1.07e-2 against 1.24e-2,
        tools/check_surface_render_err.py), but the 48 rays of the reference's surface-rendering golden were pinned with this
        program (normals 4.6e-3; 1.12e-2 with the other, over the tier's 1e-2 bar on that sample)."""
        if self._needs_grad(x, torch.is_grad_enabled()) or not _lib.tensor_tier():
            return self.forward(x)
        _lib.require_cuda(x)
        return self._run_umma(x, "sdf", sdf_via_rev=False)[0]

    def _run_umma(self, x, mode, want_feat=False, radiance_net=None, view_dirs=None, want_sdf=True,
                  want_nablas=True, normal_scale=None, sdf_via_rev=True):
        """The fused tcgen05 kernel.  mode: 'sdf' | 'nablas' | 'fused' (one launch, radiance steps 32 columns wide) |
        'split' (two launches per <= 2M points: SDF net + normals + feature image, then the radiance net on 128-point
        tiles -- the radiance MMAs run 128 columns wide, 2.5x the rate of the 32-column ones)."""
        _lib.require_cuda(x, view_dirs)
        lib = _lib.get_lib()
        shape = x.shape[:-1]
        xf = _lib.f32c(x.detach().reshape(-1, 3))
        n, dev = xf.shape[0], xf.device
        with_rad = mode in ("fused", "split")
        net = self._umma_net(radiance_net if with_rad else None)
        f = dict(dtype=torch.float32, device=dev)
        sdf = torch.empty(n, **f) if want_sdf else None
        nabla = torch.empty(n, 3, **f) if (mode != "sdf" and (want_nablas or mode == "split")) else None
        feat = torch.empty(n, net.feat_dim, **f) if (want_feat and not with_rad) else None
        rgb = torch.empty(n, 3, **f) if with_rad else None
        vf = None
        if with_rad:      # view_dirs None (use_view_dirs=False): any finite value, it meets zero weights
            vf = xf if view_dirs is None else _lib.f32c(view_dirs.detach().expand(*shape, 3).reshape(-1, 3))

        pair = _PAIR_KERNEL and net.pair_ok()

        def launch_pair(prog, i0, m, sdf_o, nabla_o, feat_o, img):
            sl = lambda t: None if t is None else t[i0:i0 + m]
            _lib.check(lib.nr_mlp_umma2_forward(
                C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias), net.bias.numel(),
                _lib.ptr(xf[i0:i0 + m]), m, _lib.ptr(sl(sdf_o)), _lib.ptr(sl(nabla_o)), _lib.ptr(sl(feat_o)),
                net.feat_dim, _lib.ptr(img), _lib.stream_ptr(dev)), "mlp_umma2_forward")

        def launch(prog, i0, m, sdf_o, nabla_o, feat_o, rgb_o, img):
            sl = lambda t: None if t is None else t[i0:i0 + m]
            _lib.check(lib.nr_mlp_umma_forward(
                C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias), net.bias.numel(),
                _lib.ptr(xf[i0:i0 + m]), _lib.ptr(sl(vf)), m, _lib.ptr(sl(sdf_o)), _lib.ptr(sl(nabla_o)), _lib.ptr(sl(feat_o)),
                net.feat_dim, _lib.ptr(sl(rgb_o)), _lib.ptr(normal_scale), _lib.ptr(img), _lib.stream_ptr(dev)),
                "mlp_umma_forward")

        def launch_rev(prog, i0, m, sdf_o, nabla_o, feat_o, img):
            sl = lambda t: None if t is None else t[i0:i0 + m]
            need = lib.nr_mlp_umma_reverse_workspace(C.byref(prog), m)
            ws = _lib.workspace(need, dev, slot=2)
            _lib.check(lib.nr_mlp_umma_reverse(
                C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias), net.bias.numel(),
                _lib.ptr(xf[i0:i0 + m]), m, _lib.ptr(sl(sdf_o)), _lib.ptr(sl(nabla_o)), _lib.ptr(sl(feat_o)),
                net.feat_dim, _lib.ptr(img), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "mlp_umma_reverse")

        def launch_split(snet, prog, i0, m, sdf_o, nabla_o, feat_o, img):
            sl = lambda t: None if t is None else t[i0:i0 + m]
            need = lib.nr_mlp_split_reverse_workspace(C.byref(prog), m)
            ws = _lib.workspace(need, dev, slot=2)
            _lib.check(lib.nr_mlp_split_reverse(
                C.byref(prog), _lib.ptr(snet.image), snet.image.numel() * 2, _lib.ptr(snet.bias), snet.bias.numel(),
                _lib.ptr(xf[i0:i0 + m]), m, _lib.ptr(sl(sdf_o)), _lib.ptr(sl(nabla_o)), _lib.ptr(sl(feat_o)),
                snet.feat_dim, _lib.ptr(img), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "mlp_split_reverse")

        if _lib.split_tier():
            # 'fp16x2': the SDF net (where softplus(beta=100) amplifies operand rounding) on split-precision operands,
            # csrc/mlp_rev_split.cu; the radiance pass on plain fp16 operands, fed by the same operand image
            snet = self._umma_split_net()
            with torch.cuda.device(dev):
                if with_rad:
                    p_geo, p_rad = snet.program("rev_img"), net.program("radiance")
                    step = _SPLIT_POINTS
                    img = _lib.workspace((min(n, step) + 127) // 128 * 65536, dev, slot=1)
                    for i0 in range(0, n, step):
                        m = min(step, n - i0)
                        launch_split(snet, p_geo, i0, m, sdf, nabla, None, img)
                        launch(p_rad, i0, m, None, nabla, None, rgb, img)
                elif mode == "sdf":
                    launch_split(snet, snet.program("rev_sdf", want_feat=want_feat), 0, n, sdf, None, feat, None)
                else:
                    launch_split(snet, snet.program("rev", want_feat=want_feat), 0, n, sdf, nabla, feat, None)
            rs = lambda t, *tail: None if t is None else t.reshape(*shape, *tail)
            return rs(sdf), rs(nabla if want_nablas else None, 3), rs(feat, net.feat_dim), rs(rgb, 3)

        # bf16 operands keep the tangent tiles: the backward sweep rounds every layer's gradient to 8 bits of mantissa,
        # which lands on the tier's 1e-2 bar (1.06e-2 measured), the forward-mode tangents stay under it
        rev = _REVERSE_NABLAS and not pair and net.operand == "fp16" and net.rev_ok(want_feat)
        with torch.cuda.device(dev):
            if mode == "split" and rev:
                p_geo, p_rad = net.program("rev_img"), net.program("radiance")
                step = _SPLIT_POINTS
                img = _lib.workspace((min(n, step) + 127) // 128 * 65536, dev, slot=1)
                for i0 in range(0, n, step):
                    m = min(step, n - i0)
                    launch_rev(p_geo, i0, m, sdf, nabla, None, img)
                    launch(p_rad, i0, m, None, nabla, None, rgb, img)
            elif mode == "nablas" and rev:
                launch_rev(net.program("rev", want_feat=want_feat), 0, n, sdf, nabla, feat, None)
            elif mode == "sdf" and rev and _SDF_VIA_REV and sdf_via_rev:
                # sdf (+ feature) only: the forward sweep of the reverse-mode kernel alone (its one-tanh activation on 16
                # shared epilogue warps is faster than mlp_umma_kernel's 'sdf' program)
                launch_rev(net.program("rev_sdf", want_feat=want_feat), 0, n, sdf, None, feat, None)
            elif mode == "split":
                # one-CTA kernel: image = last hidden activations, the radiance pass applies the feature layer 128 columns
                # wide; pair kernel: image = the feature, from its own 32-column step
                p_geo = net.program("nablas_imgf" if pair else "nablas_img", pair=pair)
                p_rad = net.program("radiancef" if pair else "radiance")
                step = _SPLIT_POINTS
                img = _lib.workspace((min(n, step) + 127) // 128 * 65536, dev, slot=1)
                for i0 in range(0, n, step):
                    m = min(step, n - i0)
                    if pair:
                        launch_pair(p_geo, i0, m, sdf, nabla, None, img)
                    else:
                        launch(p_geo, i0, m, sdf, nabla, None, None, img)
                    launch(p_rad, i0, m, None, nabla, None, rgb, img)
            elif pair and mode in ("sdf", "nablas"):
                launch_pair(net.program(mode, want_feat=want_feat, pair=True), 0, n, sdf, nabla, feat, None)
            else:
                launch(net.program(mode, want_feat=want_feat), 0, n, sdf, nabla, feat, rgb, None)
        rs = lambda t, *tail: None if t is None else t.reshape(*shape, *tail)
        return rs(sdf), rs(nabla if want_nablas else None, 3), rs(feat, net.feat_dim), rs(rgb, 3)

    # ---- forward -----------------------------------------------------------------------------
    def _needs_grad(self, x, has_grad):
        return has_grad and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def _run(self, x, want_nablas, want_feat):
        _lib.require_cuda(x)
        if _lib.tensor_tier():
            sdf, nabla, feat, _ = self._run_umma(x, "nablas" if want_nablas else "sdf", want_feat=want_feat)
            return sdf, nabla, feat
        lib = _lib.get_lib()
        shape = x.shape[:-1]
        xf = _lib.f32c(x.detach().reshape(-1, 3))
        n = xf.shape[0]
        dev = xf.device
        desc = self._descriptor()
        sdf = torch.empty(n, dtype=torch.float32, device=dev)
        nabla = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_nablas else None
        feat = torch.empty(n, self.W_geo_feat, dtype=torch.float32, device=dev) if want_feat else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            for i0 in range(0, n, _MAX_POINTS_PER_CALL):
                m = min(_MAX_POINTS_PER_CALL, n - i0)
                xs = xf[i0:i0 + m]
                fptr = _lib.ptr(feat[i0:i0 + m]) if want_feat else None
                if want_nablas:
                    need = lib.nr_sdf_forward_nablas_f32_workspace(C.byref(desc), m)
                    ws = _lib.workspace(need, dev)
                    _lib.check(lib.nr_sdf_forward_nablas_f32(
                        C.byref(desc), _lib.ptr(xs), m, _lib.ptr(sdf[i0:i0 + m]), _lib.ptr(nabla[i0:i0 + m]),
                        fptr, self.W_geo_feat, _lib.ptr(ws), ws.numel(), st), "sdf_forward_nablas")
                else:
                    need = lib.nr_sdf_forward_f32_workspace(C.byref(desc), m)
                    ws = _lib.workspace(need, dev)
                    _lib.check(lib.nr_sdf_forward_f32(
                        C.byref(desc), _lib.ptr(xs), m, _lib.ptr(sdf[i0:i0 + m]), fptr, self.W_geo_feat,
                        _lib.ptr(ws), ws.numel(), st), "sdf_forward")
        sdf = sdf.reshape(shape)
        if nabla is not None:
            nabla = nabla.reshape(*shape, 3)
        if feat is not None:
            feat = feat.reshape(*shape, self.W_geo_feat)
        return sdf, nabla, feat

    def forward(self, x, return_h=False):
        """base.py:243-263."""
        if self._needs_grad(x, torch.is_grad_enabled()):
            from .autograd import sdf_forward_autograd
            return sdf_forward_autograd(self, x, return_h)
        sdf, _, feat = self._run(x, want_nablas=False, want_feat=return_h)
        return (sdf, feat) if return_h else sdf

    def forward_with_nablas(self, x, has_grad_bypass=None):
        """base.py:265-282: (sdf, d sdf / d x, geometry feature).  The normal is the analytic
        forward-mode derivative evaluated inside the kernel instead of ``autograd.grad``."""
        has_grad = torch.is_grad_enabled() if has_grad_bypass is None else has_grad_bypass
        if self._needs_grad(x, has_grad):
            from .autograd import sdf_forward_with_nablas_autograd
            return sdf_forward_with_nablas_autograd(self, x)
        return self._run(x, want_nablas=True, want_feat=True)


class RadianceNet(nn.Module):
    """Appearance MLP (base.py:312-391): cat([PE(x), PE(view), normals, feature]) -> D ReLU
    layers -> Sigmoid(3)."""

    def __init__(self, D=4, W=256, skips=[], W_geo_feat=256, embed_multires=6, embed_multires_view=4,
                 use_view_dirs=True, weight_norm=True, use_siren=False):
        super().__init__()
        if use_siren:
            raise NotImplementedError("SIREN radiance nets are outside the neurecon_b200 hot-path scope")
        self.skips = list(skips)
        self.D, self.W = D, W
        self.use_view_dirs = use_view_dirs
        self.W_geo_feat = W_geo_feat
        self.embed_multires, self.embed_multires_view = embed_multires, embed_multires_view
        self.embed_fn, in_pts = get_embedder(embed_multires)
        if use_view_dirs:
            self.embed_fn_view, in_views = get_embedder(embed_multires_view)
            in0 = in_pts + in_views + 3 + W_geo_feat
        else:
            in0 = in_pts + W_geo_feat
        layers = []
        for l in range(D + 1):
            out_dim = 3 if l == D else W
            in_dim = in0 if l == 0 else (in0 + W if l in self.skips else W)
            act = nn.Sigmoid() if l == D else nn.ReLU(inplace=True)
            layer = DenseLayer(in_dim, out_dim, activation=act)
            if weight_norm:
                layer = _weight_norm(layer)
            layers.append(layer)
        self.layers = nn.ModuleList(layers)
        self._cache = {}

    # ``use_view_dirs=False`` (base.py:335-336,383-384: layer 0 sees cat([PE(x), feature]) only) runs on the same kernels:
    # layer 0's weight gets zero columns where the kernels place [view (un-embedded) | normals], so whatever is passed
    # for those six inputs is multiplied by zero.
    @property
    def _multires_view_eff(self):
        return self.embed_multires_view if self.use_view_dirs else -1

    def _layer0_weight(self, W0):
        """Layer-0 weight in the kernels' input order [PE(x) | PE(view) | normals | feature] (differentiable)."""
        if self.use_view_dirs:
            return W0
        px = W0.shape[1] - self.W_geo_feat
        return torch.cat([W0[:, :px], W0.new_zeros(W0.shape[0], 6), W0[:, px:]], dim=1)

    def _effective_weights(self):
        Wl = [_effective_weight(l) for l in self.layers]
        Wl[0] = self._layer0_weight(Wl[0])
        return Wl

    def _descriptor(self):
        if self.skips:
            raise NotImplementedError("neurecon_b200 RadianceNet supports skips=[] (every shipped reference config)")
        def build():
            Wl = self._effective_weights()
            Ws, bs = _pack_layers(Wl, [l.bias for l in self.layers])
            d = _lib.RadianceNetDesc()
            d.n_layers = self.D + 1
            d.multires, d.multires_view, d.feat_dim = self.embed_multires, self._multires_view_eff, self.W_geo_feat
            for i in range(self.D + 1):
                out_d, in_d = Wl[i].shape
                d.in_dim[i], d.out_dim[i] = in_d, out_d
                d.W[i], d.b[i] = Ws[i].data_ptr(), bs[i].data_ptr()
            d.umma_image = None
            d.umma_bias = None
            return d, (Ws, bs)
        return _cached(self, "desc", _param_key(self), build)[0]

    def forward(self, x, view_dirs, normals, geometry_feature):
        """base.py:372-391."""
        needs_grad = torch.is_grad_enabled() and (
            (normals is not None and normals.requires_grad) or geometry_feature.requires_grad
            or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            from .autograd import radiance_forward_autograd
            return radiance_forward_autograd(self, x, view_dirs, normals, geometry_feature)
        _lib.require_cuda(x, view_dirs, normals, geometry_feature)
        lib = _lib.get_lib()
        shape = x.shape[:-1]
        xf = _lib.f32c(x.detach().reshape(-1, 3))
        # use_view_dirs=False: the reference ignores view_dirs and normals (base.py:383-384); they meet zero weights here
        vf = xf if (view_dirs is None or not self.use_view_dirs) else _lib.f32c(view_dirs.detach().expand(*shape, 3).reshape(-1, 3))
        nf = xf if (normals is None or not self.use_view_dirs) else _lib.f32c(normals.detach().reshape(-1, 3))
        ff = _lib.f32c(geometry_feature.detach().reshape(-1, self.W_geo_feat))
        n, dev = xf.shape[0], xf.device
        desc = self._descriptor()
        rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            for i0 in range(0, n, _MAX_POINTS_PER_CALL):
                m = min(_MAX_POINTS_PER_CALL, n - i0)
                need = lib.nr_radiance_forward_f32_workspace(C.byref(desc), m)
                ws = _lib.workspace(need, dev)
                _lib.check(lib.nr_radiance_forward_f32(
                    C.byref(desc), _lib.ptr(xf[i0:i0 + m]), _lib.ptr(vf[i0:i0 + m]), _lib.ptr(nf[i0:i0 + m]),
                    _lib.ptr(ff[i0:i0 + m]), self.W_geo_feat, m, _lib.ptr(rgb[i0:i0 + m]), _lib.ptr(ws),
                    ws.numel(), st), "radiance_forward")
        return rgb.reshape(*shape, 3)


def query_radiance(surface, radiance_net, x, view_dirs, chunk_normalize=False):
    """Inference composition of ``forward_with_nablas`` + ``RadianceNet.forward`` at the same
    points (NeuS.forward_radiance neus.py:103-106, VolSDF.forward volsdf.py:327-331, UNISURF.forward
    unisurf.py:34-38): returns (radiance, sdf, nablas).  Tensor tier: ONE fused kernel, the 256-wide
    feature and the normal never leave the SM; fp32 tier: the two library calls.

    ``chunk_normalize`` reproduces UNISURF's ``F.normalize(nablas)`` (no ``dim`` => dim=1 => the POINT
    axis of the whole chunk) for a flat chunk x [n, 3]: every component is
    divided by its L2 norm over all n points.  That couples the points, so it takes two passes: normals
    first, then the (3-float) column norms, then the radiance with the scale applied inside the kernel."""
    if not chunk_normalize:
        if _lib.tensor_tier():
            sdf, nabla, _, rgb = surface._run_umma(x, _FUSED_MODE, radiance_net=radiance_net, view_dirs=view_dirs)
            return rgb, sdf, nabla
        sdf, nabla, feat = surface._run(x, want_nablas=True, want_feat=True)
        return radiance_net.forward(x, view_dirs, nabla, feat), sdf, nabla
    if _lib.tensor_tier():
        sdf, nabla, _, _ = surface._run_umma(x, "nablas")
        scale = (1.0 / nabla.reshape(-1, 3).norm(dim=0).clamp_min(1e-12)).contiguous()
        _, _, _, rgb = surface._run_umma(x, _FUSED_MODE, radiance_net=radiance_net, view_dirs=view_dirs, want_sdf=False,
                                         want_nablas=False, normal_scale=scale)
        return rgb, sdf, nabla
    sdf, nabla, feat = surface._run(x, want_nablas=True, want_feat=True)
    normals = nabla / nabla.reshape(-1, 3).norm(dim=0).clamp_min(1e-12)
    return radiance_net.forward(x, view_dirs, normals, feat), sdf, nabla


class NeRF(nn.Module):
    """NeRF++ background MLP (base.py:395-453): same constructor, parameter names
    (``pts_linears.{i}``, ``views_linears.0``, ``feature_linear``, ``alpha_linear``,
    ``rgb_linear``; plain ``weight``/``bias``) and ``forward(input_pts, input_views) -> (sigma, rgb)``."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_view=3, multires=-1, multires_view=-1, output_ch=4,
                 skips=[4], use_view_dirs=False):
        super().__init__()
        self.D, self.W = D, W
        self.skips = list(skips)
        self.use_view_dirs = use_view_dirs
        self.multires, self.multires_view = multires, multires_view
        self.input_dim, self.input_dim_view = input_ch, input_ch_view
        self.embed_fn, input_ch = get_embedder(multires, input_dim=input_ch)
        self.embed_fn_view, input_ch_view = get_embedder(multires_view, input_dim=input_ch_view)
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)]
            + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_view + W, W // 2)])
        if use_view_dirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self._cache = {}

    def forward(self, input_pts, input_views):
        from .nerfpp import nerf_forward
        return nerf_forward(self, input_pts, input_views)
