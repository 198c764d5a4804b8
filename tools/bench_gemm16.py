"""Micro-benchmark of the 16-bit training GEMMs (csrc/gemm16.cu) at the shapes of a 512-ray NeuS step:
n = 130 560 rows, 256 x 256 weights.  Usage: python tools/bench_gemm16.py [rows] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib  # noqa: E402
from neurecon_b200.models import autograd_rev as ar  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 130560
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
lib = _lib.get_lib()
g = torch.Generator(device=dev).manual_seed(0)
h16 = dict(dtype=torch.float16, device=dev)
A = torch.randn(n, 256, device=dev, generator=g).half()
S = torch.rand(n, 256, device=dev, generator=g).half()
Pm = torch.randn(n, 256, device=dev, generator=g).half()
W = torch.randn(256, 256, device=dev, generator=g) * 0.08
b = torch.zeros(256, device=dev)
Wp = ar._Packed(W, 256, 256)
Y, O2 = torch.empty(n, 256, **h16), torch.empty(n, 256, **h16)
dW = torch.zeros(256, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, mbytes):
    print("%-44s %7.1f us   %6.0f GB/s of %5.0f MB moved, %5.1f TFLOP/s" % (name, ms * 1e3, mbytes / ms, mbytes, 2.0 * n * 256 * 256 / ms / 1e9), flush=True)


mb = n * 256 * 2 / 1e6
st = _lib.stream_ptr(dev)
report("G_SOFTPLUS (A in; h, S out)", timeit(lambda: ar._gemm16(A, Wp, b, n, 256, 256, Y, 1, ar.G_SOFTPLUS, out2=O2)), 3 * mb)
report("G_SCALE (A, S in; p out)", timeit(lambda: ar._gemm16(A, Wp, None, n, 256, 256, Y, 1, ar.G_SCALE, aux_a=S)), 3 * mb)
report("G_SCALE + addend (A, S, Z2 in; z out)", timeit(lambda: ar._gemm16(A, Wp, None, n, 256, 256, Y, 1, ar.G_SCALE, aux_a=S, aux_b=Pm)), 4 * mb)
report("G_ADJ (A, S, p in; gb, zb2 out)", timeit(lambda: ar._gemm16(A, Wp, None, n, 256, 256, Y, 1, ar.G_ADJ, aux_a=S, aux_b=Pm, out2=O2)), 5 * mb)
report("G_LINEAR fp32 out", timeit(lambda: ar._gemm16(A, Wp, b, n, 256, 256, torch.empty(n, 256, device=dev), 0, ar.G_LINEAR)), 3 * mb)
report("gemm16_tn (G, X in)", timeit(lambda: _lib.check(lib.nr_gemm16_tn(_lib.ptr(A), 256, _lib.ptr(Pm), 256, n, 256, 256, _lib.ptr(dW), 256, 1.0, st), "tn")), 2 * mb)
c = torch.zeros(256, device=dev)
report("colsum16", timeit(lambda: _lib.check(lib.nr_colsum16(_lib.ptr(A), 256, n, 256, 1.0, _lib.ptr(c), st), "cs")), mb)
