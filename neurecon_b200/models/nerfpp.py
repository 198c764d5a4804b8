"""NeRF++ background network forward (models/base.py:426-453) on the fp32 GEMM kernels."""
import torch

from .. import _lib
from .._lib import C
from .base import _cached, _pack_layers, _param_key, _MAX_POINTS_PER_CALL


def _descriptor(module):
    if not module.use_view_dirs:
        raise NotImplementedError("neurecon_b200 NeRF supports use_view_dirs=True (the NeRF++ background configuration)")
    if len(module.skips) > 1:
        raise NotImplementedError("neurecon_b200 NeRF supports one skip")
    def build():
        lin = list(module.pts_linears)
        Ws, bs = _pack_layers([l.weight for l in lin], [l.bias for l in lin])
        others = [module.alpha_linear, module.feature_linear, module.views_linears[0], module.rgb_linear]
        Wo, bo = _pack_layers([l.weight for l in others], [l.bias for l in others])
        d = _lib.NerfNet()
        d.depth, d.width, d.input_dim = module.D, module.W, module.input_dim
        d.multires, d.multires_view = module.multires, module.multires_view
        d.skip = module.skips[0] if module.skips else -1
        for i in range(module.D):
            d.pts_W[i], d.pts_b[i] = Ws[i].data_ptr(), bs[i].data_ptr()
        d.alpha_W, d.alpha_b = Wo[0].data_ptr(), bo[0].data_ptr()
        d.feature_W, d.feature_b = Wo[1].data_ptr(), bo[1].data_ptr()
        d.views_W, d.views_b = Wo[2].data_ptr(), bo[2].data_ptr()
        d.rgb_W, d.rgb_b = Wo[3].data_ptr(), bo[3].data_ptr()
        return d, (Ws, bs, Wo, bo)        # the descriptor holds raw pointers: keep the tensors alive with it
    return _cached(module, "desc", _param_key(module), build)[0]


def _umma_net(module):
    from ..umma_pack import UmmaNerfNet
    return _cached(module, "umma", (_param_key(module), _lib.operand()),
                   lambda: UmmaNerfNet(module, operand=_lib.operand()))


def nerf_forward(module, input_pts, input_views):
    """NeRF.forward(input_pts [..., 4], input_views [..., 3]) -> (sigma [...], rgb [..., 3])."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        from .autograd import nerf_forward_autograd
        return nerf_forward_autograd(module, input_pts, input_views)
    _lib.require_cuda(input_pts, input_views)
    lib = _lib.get_lib()
    shape = input_pts.shape[:-1]
    xf = _lib.f32c(input_pts.detach().reshape(-1, module.input_dim))
    vf = _lib.f32c(input_views.detach().expand(*shape, 3).reshape(-1, 3))
    n, dev = xf.shape[0], xf.device
    sigma = torch.empty(n, dtype=torch.float32, device=dev)
    rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
    if _lib.tensor_tier() and n > 0:
        # tensor tier: the whole net in one launch of the fused tcgen05 kernel (csrc/mlp_umma.cu, input_mode 2)
        net = _umma_net(module)
        prog = net.program()
        with torch.cuda.device(dev):
            _lib.check(lib.nr_mlp_umma_forward(
                C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias), net.bias.numel(),
                _lib.ptr(xf), _lib.ptr(vf), n, _lib.ptr(sigma), None, None, 0, _lib.ptr(rgb), None, None,
                _lib.stream_ptr(dev)), "mlp_umma_forward (NeRF++)")
        return sigma.reshape(shape), rgb.reshape(*shape, 3)
    desc = _descriptor(module)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        for i0 in range(0, n, _MAX_POINTS_PER_CALL):
            m = min(_MAX_POINTS_PER_CALL, n - i0)
            need = lib.nr_nerf_forward_f32_workspace(C.byref(desc), m)
            ws = _lib.workspace(need, dev)
            _lib.check(lib.nr_nerf_forward_f32(C.byref(desc), _lib.ptr(xf[i0:i0 + m]), _lib.ptr(vf[i0:i0 + m]), m,
                                               _lib.ptr(sigma[i0:i0 + m]), _lib.ptr(rgb[i0:i0 + m]), _lib.ptr(ws),
                                               ws.numel(), st), "nerf_forward")
    return sigma.reshape(shape), rgb.reshape(*shape, 3)
