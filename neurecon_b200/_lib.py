"""ctypes binding of include/neurecon_b200.h (the C-ABI drop-in boundary)."""
import ctypes as C
import os
import threading

import torch

NR_MAX_LAYERS = 16
ACT_NONE, ACT_SOFTPLUS100, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libneurecon_b200.so")
_lib = None
_lock = threading.Lock()
_precision = os.environ.get("NEURECON_B200_PRECISION", "fp16")


class SdfNet(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("multires", C.c_int32), ("skip_layer", C.c_int32), ("width", C.c_int32),
        ("in_dim", C.c_int32 * NR_MAX_LAYERS), ("out_dim", C.c_int32 * NR_MAX_LAYERS),
        ("W", C.c_void_p * NR_MAX_LAYERS), ("b", C.c_void_p * NR_MAX_LAYERS),
        ("umma_image", C.c_void_p), ("umma_bias", C.c_void_p),
    ]


class RadianceNetDesc(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("multires", C.c_int32), ("multires_view", C.c_int32), ("feat_dim", C.c_int32),
        ("in_dim", C.c_int32 * NR_MAX_LAYERS), ("out_dim", C.c_int32 * NR_MAX_LAYERS),
        ("W", C.c_void_p * NR_MAX_LAYERS), ("b", C.c_void_p * NR_MAX_LAYERS),
        ("umma_image", C.c_void_p), ("umma_bias", C.c_void_p),
    ]


class NerfNet(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("depth", "width", "input_dim", "multires", "multires_view", "skip")] + [
        ("pts_W", C.c_void_p * NR_MAX_LAYERS), ("pts_b", C.c_void_p * NR_MAX_LAYERS),
        ("alpha_W", C.c_void_p), ("alpha_b", C.c_void_p), ("feature_W", C.c_void_p), ("feature_b", C.c_void_p),
        ("views_W", C.c_void_p), ("views_b", C.c_void_p), ("rgb_W", C.c_void_p), ("rgb_b", C.c_void_p)]


NR_UMMA_MAX_STEPS = 24


class UmmaStep(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("chunk_begin", "n_mt", "k_steps", "n_cols", "epi", "bias_off", "out_rows",
                                         "pe_fill", "to_rad", "accumulate", "sig_slot", "aux_off")]


class UmmaProgram(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_steps", "tangents", "multires", "rad_multires", "rad_multires_view",
                                         "rad_extra_rows", "operand_f16", "debug_flags", "input_mode", "input_dim", "reverse")] + [("steps", UmmaStep * NR_UMMA_MAX_STEPS)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

_SIGNATURES = {
    "nr_version": (C.c_int, []),
    "nr_last_error": (C.c_int, [C.c_char_p, _SZ]),
    "nr_launch_count": (C.c_longlong, []),
    "nr_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "nr_sdf_forward_f32_workspace": (_SZ, [C.POINTER(SdfNet), _I64]),
    "nr_sdf_forward_f32": (C.c_int, [C.POINTER(SdfNet), _P, _I64, _P, _P, _I64, _P, _SZ, _P]),
    "nr_sdf_forward_nablas_f32_workspace": (_SZ, [C.POINTER(SdfNet), _I64]),
    "nr_sdf_forward_nablas_f32": (C.c_int, [C.POINTER(SdfNet), _P, _I64, _P, _P, _P, _I64, _P, _SZ, _P]),
    "nr_radiance_forward_f32_workspace": (_SZ, [C.POINTER(RadianceNetDesc), _I64]),
    "nr_radiance_forward_f32": (C.c_int, [C.POINTER(RadianceNetDesc), _P, _P, _P, _P, _I64, _I64, _P, _P, _SZ, _P]),
    "nr_gemm_f32": (C.c_int, [_P, _I32, _P, _I32, _P, _I64, _I32, _I32, _P, _I32, _I32, _P, _I32, _P, _I32, _I64, _P]),
    "nr_gemm_tc": (C.c_int, [_P, _I32, _P, _I32, _P, _I64, _I32, _I32, _P, _I32, _I32, _P, _I32, _P, _I32, _I64, _I32, _P]),
    "nr_gemm_tn_f32": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _I32, _P, _I32, _P]),
    "nr_gemm_tn_tc": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _I32, _P, _I32, _I32, _P]),
    "nr_colsum_f32": (C.c_int, [_P, _I32, _I64, _I32, _P, _P]),
    "nr_sdf_bwd_act_f32": (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _P, _I32, _I64, _I32, _I32, _P, _P]),
    "nr_act_bwd_f32": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _I32, _P]),
    "nr_embed_f32": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _I32, _P, _I32, _I32, _P]),
    "nr_nerf_forward_f32_workspace": (_SZ, [C.POINTER(NerfNet), _I64]),
    "nr_nerf_forward_f32": (C.c_int, [C.POINTER(NerfNet), _P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "nr_near_far_from_sphere": (C.c_int, [_P, _P, _I64, _F, _P, _P, _P]),
    "nr_grid_points": (C.c_int, [_I64, _I64, _I32, C.c_double, _I32, _P, _P]),
    "nr_get_rays": (C.c_int, [_P, _P, _P, _I32, _I32, _I64, _P, _P, _P]),
    "nr_sample_pdf": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _P, _P, _P]),
    "nr_neus_ray_setup": (C.c_int, [_P, _P, _I64, _F, _F, _F, _I32, _P, _P, _P, _P, _P, _P]),
    "nr_neus_upsample_step": (C.c_int, [_P, _P, _I64, _P, _P, _I32, _I32, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nr_mc_count_workspace": (_SZ, [_I32, _I32, _I32]),
    "nr_mc_blocks": (_I64, [_I32, _I32, _I32]),
    "nr_mc_count": (C.c_int, [_P, _I32, _I32, _I32, _F, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "nr_mc_generate": (C.c_int, [_P, _I32, _I32, _I32, _F, _F, _F, _F, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nr_neus_sdf_to_w": (C.c_int, [_P, _F, _I64, _I32, _P, _P]),
    "nr_neus_outside_points": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P]),
    "nr_neus_composite_bg": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P,
                                       _P, _P, _P, _P]),
    "nr_volsdf_error_bound": (C.c_int, [_P, _P, _I64, _I32, _P, _I32, _P, _I32, _P, _P, _P]),
    "nr_volsdf_ray_setup": (C.c_int, [_P, _P, _I64, _F, _F, _F, _I32, _P, _P, _P, _P, _I32, _P, _P]),
    "nr_sphere_min": (C.c_int, [_P, _P, _I64, _F, _P]),
    "nr_volsdf_fine_iter": (C.c_int, [_P, _P, _P, _I64, _P, _P, _I32, _I32, _P, _I32, _P, _P, _P, _F, _I32, _I32, _I32,
                                      _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nr_volsdf_merge": (C.c_int, [_P, _P, _P, _I64, _F, _I32, _P, _I32, _P, _P, _P]),
    "nr_volsdf_composite": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P,
                                      _P, _P, _P]),
    "nr_unisurf_ray_setup": (C.c_int, [_P, _P, _I64, _F, _F, _F, _I32, _P, _P, _P, _P, _P]),
    "nr_unisurf_first_crossing": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _F, _P, _P, _P, _P, _P, _P]),
    "nr_unisurf_secant_step": (C.c_int, [_P, _F, _P, _P, _P, _I64, _P, _P, _P]),
    "nr_unisurf_sample": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _F, _F, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "nr_sphere_trace_step": (C.c_int, [_P, _P, _P, _F, _I64, _P, _P, _P, _P]),
    "nr_unisurf_composite": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "nr_mlp_umma_forward": (C.c_int, [C.POINTER(UmmaProgram), _P, _SZ, _P, _SZ, _P, _P, _I64, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "nr_mlp_umma_reverse_workspace": (_SZ, [C.POINTER(UmmaProgram), _I64]),
    "nr_mlp_umma_reverse": (C.c_int, [C.POINTER(UmmaProgram), _P, _SZ, _P, _SZ, _P, _I64, _P, _P, _P, _I64, _P, _P, _SZ, _P]),
    "nr_neus_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _F, _F, _P, _P, _P, _P, _P, _P]),
    "nr_grad_sqsum": (C.c_int, [_P, _I32, _P, _P]),
    "nr_adam_step": (C.c_int, [_P, _I32, _F, _F, _F, _F, _I64, _P]),
    "nr_adam_step_dev": (C.c_int, [_P, _I32, _P, _F, _F, _F, _P, _P]),
    "nr_mlp_umma2_forward": (C.c_int, [C.POINTER(UmmaProgram), _P, _SZ, _P, _SZ, _P, _I64, _P, _P, _P, _I64, _P, _P]),
    "nr_mlp_umma_set_trace": (C.c_int, [_P]),
    "nr_umma_pack_a_bytes": (_SZ, [_I32, _I32, _I32]),
    "nr_umma_pack_a": (C.c_int, [_P, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "nr_mlp_split_reverse_workspace": (_SZ, [C.POINTER(UmmaProgram), _I64]),
    "nr_mlp_split_reverse": (C.c_int, [C.POINTER(UmmaProgram), _P, _SZ, _P, _SZ, _P, _I64, _P, _P, _P, _I64, _P, _P, _SZ, _P]),
    "nr_gemm16": (C.c_int, [_P, _I32, _P, _I32, _P, _I64, _I32, _I32, _P, _I32, _I32, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _P]),
    "nr_gemm16_pack_w_bytes": (_SZ, [_I32, _I32]),
    "nr_gemm16_pack_w": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "nr_gemm16_tn": (C.c_int, [_P, _I32, _P, _I32, _I64, _I32, _I32, _P, _I32, _F, _P]),
    "nr_gemm16_tn2": (C.c_int, [_P, _I32, _P, _I32, _P, _I32, _P, _I32, _I64, _I32, _I32, _P, _I32, _F, _P]),
    "nr_colsum16": (C.c_int, [_P, _I32, _I64, _I32, _F, _P, _P]),
    "nr_gemm16_split": (C.c_int, [_P, _I32, _P, _P, _I64, _I32, _I32, _P, _I32, _I32, _I32, _I32, _P, _I32, _P, _I32, _P]),
    "nr_gemm16_pack_w_split_bytes": (_SZ, [_I32, _I32]),
    "nr_gemm16_pack_w_split": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "nr_pe16_split": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _P]),
    "nr_cast_cols16": (C.c_int, [_P, _I64, _I64, _I32, _F, _P, _I64, _I32, _P]),
    "nr_pe16": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _I32, _I32, _P]),
    "nr_pe_jac_t": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, _I32, _I32, _P, _P]),
    "nr_pe_jac": (C.c_int, [_P, _I64, _I32, _P, _F, _P, _I32, _I32, _P, _I32, _I32, _P]),
    "nr_weight_norm": (C.c_int, [_P, _I32, _I64, _I32, _P]),
    "nr_sphere_intersection": (C.c_int, [_P, _P, _I64, C.c_double, _P, _P, _P, _P]),
    "nr_dvals_from_radius": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _P, _P]),
    "nr_volsdf_outside_points": (C.c_int, [_P, _P, _I64, _F, _I32, _P, _P, _P, _P, _P]),
    "nr_neus_composite_bwd": (C.c_int, [_P] * 10 + [_I64, _I32, _I32, _P, _P, _P, _F, _I32] + [_P] * 12),
    "nr_volsdf_composite_bwd": (C.c_int, [_P] * 11 + [_I64, _I32, _P, _P, _I32, _I32] + [_P] * 13),
    "nr_unisurf_composite_bwd": (C.c_int, [_P] * 8 + [_I64, _I32, _I32] + [_P] * 9),
    "nr_neus_composite": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
}


# self-tests and probes: include/neurecon_b200_devtools.h, lib/libneurecon_b200_devtools.so (not in the production library)
_DEVTOOLS_SIGNATURES = {
    "nr_bench_ldtm": (C.c_int, [_I32, _I32, _I32, _P, _P, _P]),
    "nr_probe_alu": (C.c_int, [_I32, _I32, _I32, _I32, _P, _P, _P]),
    "nr_probe_epi": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P]),
    "nr_probe_mix": (C.c_int, [_I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "nr_bench_umma": (C.c_int, [_I32, _I32, _I32, _I32, _P, _I32, _P, _P]),
    "nr_selftest_umma2": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _P]),
    "nr_selftest_umma": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _P]),
}
_devtools = None


def get_devtools():
    """ctypes handle on the development-tools library (self-tests, micro-architecture probes)."""
    global _devtools
    if _devtools is None:
        get_lib()                                   # builds everything if stale
        from . import build as _build
        lib = C.CDLL(_build.DEVTOOLS_LIB_PATH)
        for name, (res, args) in _DEVTOOLS_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        lib.nr_last_error.restype, lib.nr_last_error.argtypes = C.c_int, [C.c_char_p, _SZ]
        _devtools = lib
    return _devtools


def declared_symbols():
    """Every symbol include/neurecon_b200.h declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def library_path():
    return _LIB_PATH


def get_lib():
    """Load (building first if the sources are newer and nvcc is present) the C-ABI library.
    Raises RuntimeError when the library cannot be had -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if _build._stale():
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            if os.path.exists(nvcc):
                _build.build()
            elif not os.path.exists(_LIB_PATH):
                raise RuntimeError(
                    "neurecon_b200: CUDA library %s is missing and nvcc is not available to build it; "
                    "run `python -m neurecon_b200.build`. There is no CPU fallback." % _LIB_PATH)
        lib = C.CDLL(os.environ.get("NEURECON_B200_LIB", _LIB_PATH))  # override: kernel experiments only
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    get_lib().nr_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        msg = last_error()
        if rc == -1:
            raise ValueError("neurecon_b200 %s: %s" % (what, msg))
        raise RuntimeError("neurecon_b200 %s failed (code %d): %s" % (what, rc, msg))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "neurecon_b200: the hot path runs only on CUDA tensors (got a %s tensor); there is no CPU fallback"
                % t.device.type)


def f32c(t):
    """contiguous fp32 view/copy of a tensor (mirrors the reference's .float() at neus.py:169-170)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_workspaces = {}


def workspace(nbytes, device, slot=0):
    """Grow-only per-(device, stream, slot) scratch buffer; all kernels using it are ordered on
    the caller's current stream, so reuse across calls is safe."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def set_precision(p):
    """'fp32' (SIMT tier, <=1e-4 vs the reference), or the tcgen05 tier with 'fp16' or 'bf16'
    operands (fp32 accumulation; same tensor rate -- fp16 carries 3 more mantissa bits)."""
    global _precision
    if p not in ("fp32", "bf16", "fp16", "fp16x2"):
        raise ValueError("precision must be 'fp32', 'fp16x2', 'fp16' or 'bf16'")
    _precision = p


def get_precision():
    return _precision


def tensor_tier():
    """True when the MLPs run on the fused tcgen05 kernels."""
    return _precision in ("fp16", "bf16", "fp16x2")


def split_tier():
    """'fp16x2': the SDF net on split-precision operands (hi + lo fp16 pairs, csrc/mlp_rev_split.cu) -- the <= 1e-4 tier on
    the tensor pipe.  The radiance and NeRF++ nets, which have no softplus(beta=100) to amplify operand rounding, stay on
    plain fp16 operands; training takes the fp32 path."""
    return _precision == "fp16x2"


def operand():
    """16-bit operand type of the tensor tier's plain kernels"""
    return "bf16" if _precision == "bf16" else "fp16"
