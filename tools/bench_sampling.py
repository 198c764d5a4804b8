"""HBM roofline of the sampling / compositing kernels: algorithmic bytes / CUDA-event time against
the measured copy bandwidth.  Usage: python tools/bench_sampling.py [rays]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib  # noqa: E402
from neurecon_b200.models.frameworks import neus, volsdf, unisurf  # noqa: E402
from neurecon_b200.utils import rend_util  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 442368
dev = torch.device("cuda:0")
peak = 6533.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2


def timeit(fn, reps=5, inner=4):
    """ms per call: `inner` back-to-back calls per timing so the host-side launch cost pipelines away; the inputs
    (hundreds of MB) exceed the L2, which is additionally flushed before every timing."""
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return min(ts)


def report(name, ms, bytes_per_ray):
    gbs = R * bytes_per_ray / (ms * 1e-3) / 1e9
    print("%-34s %8.3f ms  %7.1f GB/s algorithmic  = %4.1f %% of the measured %.0f GB/s copy bandwidth (%d B/ray)" % (
        name, ms, gbs, 100 * gbs / peak, peak, bytes_per_ray), flush=True)


g = torch.Generator(device=dev).manual_seed(0)
f = dict(device=dev, generator=g)
with torch.no_grad():
    # NeuS compositing (a16): sdf[128] + nablas[128,3] + radiance[127,3] + d_mid[127] in, 32 B out
    M = 128
    sdf = torch.randn(R, M, **f) * 0.3
    nab = torch.randn(R, M, 3, **f)
    rad = torch.rand(R, M - 1, 3, **f)
    dmid = torch.rand(R, M - 1, **f).sort(-1).values
    s = torch.tensor([20.0], device=dev)
    report("nr_neus_composite (normals)", timeit(lambda: neus._composite(sdf, nab, rad, dmid, s, False, True, False)), 4084 + 32)
    # sample_pdf (a9): bins M + weights M-1 in, N out (det)
    Mb, N = 64, 16
    bins = torch.rand(R, Mb, **f).sort(-1).values
    w = torch.rand(R, Mb - 1, **f)
    report("rend_util.sample_pdf 64 -> 16 det", timeit(lambda: rend_util.sample_pdf(bins, w, N, det=True)), 4 * Mb + 4 * (Mb - 1) + 4 * N)
    uu = torch.rand(R, N, **f)
    report("rend_util.sample_pdf 64 -> 16 given u", timeit(lambda: rend_util.sample_pdf(bins, w, N, u=uu)), 4 * Mb + 4 * (Mb - 1) + 8 * N)
    report("  (warp-per-ray kernel, same call)", timeit(lambda: rend_util.sample_pdf(bins, w, N, det=True, return_details=True)),
           4 * Mb + 4 * (Mb - 1) + 4 * N + 8 * N + 4 * Mb)
    b2, w2 = torch.rand(R, 112, **f).sort(-1).values, torch.rand(R, 111, **f)
    report("rend_util.sample_pdf 112 -> 16 det", timeit(lambda: rend_util.sample_pdf(b2, w2, N, det=True)), 4 * 112 + 4 * 111 + 4 * N)
    del b2, w2
    Mb, N = 512, 64
    bins = torch.rand(R // 8, Mb, **f).sort(-1).values
    w = torch.rand(R // 8, Mb - 1, **f)
    R_keep, R = R, R // 8
    report("rend_util.sample_pdf 512 -> 64 det", timeit(lambda: rend_util.sample_pdf(bins, w, N, det=True)), 4 * Mb + 4 * (Mb - 1) + 4 * N)
    R = R_keep
    # VolSDF compositing (a17): 192 x (sdf, nabla, radiance, d) in
    Mv = 192
    R_keep, R = R, R // 2
    sdfv = torch.randn(R, Mv, **f) * 0.3
    nabv = torch.randn(R, Mv, 3, **f)
    radv = torch.rand(R, Mv, 3, **f)
    dv = torch.rand(R, Mv, **f).sort(-1).values
    lib = _lib.get_lib()
    al, be = torch.tensor([10.0], device=dev), torch.tensor([0.1], device=dev)
    rgb, depth, acc, nrm = torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, 3, device=dev)

    def fnv():
        _lib.check(lib.nr_volsdf_composite(_lib.ptr(sdfv), _lib.ptr(nabv), _lib.ptr(radv), _lib.ptr(dv), _lib.ptr(al), _lib.ptr(be),
                                           R, Mv, None, None, None, 0, 0, _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(nrm),
                                           None, None, None, _lib.stream_ptr(dev)), "volsdf_composite")
    report("nr_volsdf_composite (normals)", timeit(fnv), 6144 + 32)
    R = R_keep
    # UNISURF compositing (a18): 96 x (logit, nabla, radiance, d) in
    Mu = 96
    lgu = torch.randn(R, Mu, **f) * 3.0
    nabu = torch.randn(R, Mu, 3, **f)
    radu = torch.rand(R, Mu, 3, **f)
    du = torch.rand(R, Mu, **f).sort(-1).values
    rgbu, depu, accu, nrmu = torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, 3, device=dev)

    def fnu():
        _lib.check(lib.nr_unisurf_composite(_lib.ptr(lgu), _lib.ptr(nabu), _lib.ptr(radu), _lib.ptr(du), R, Mu, 0, _lib.ptr(rgbu),
                                            _lib.ptr(depu), _lib.ptr(accu), _lib.ptr(nrmu), None, None, _lib.stream_ptr(dev)), "unisurf_composite")
    report("nr_unisurf_composite (normals)", timeit(fnu), 3072 + 32)
