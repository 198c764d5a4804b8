"""tcgen05 rate probe: cycles per 128 x N x 16 MMA from shared-memory operands, alone and under the shared-memory
traffic of the fused MLP kernel (epilogue stores, weight bulk copies).  Usage: python tools/bench_umma_rate.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.get_devtools()
src = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
for grid in (1, 148):
    out = torch.zeros(grid, 2, dtype=torch.int64, device=dev)
    for N in (32, 64, 128, 256):
        for sw, bc in ((0, 0), (0, 1), (8, 0), (16, 0), (16, 1)):
            n = 2048
            for _ in range(2):
                _lib.check(lib.nr_bench_umma(N, n, sw, bc, _lib.ptr(src), grid, _lib.ptr(out), _lib.stream_ptr(dev)), "bench_umma")
            torch.cuda.synchronize()
            o = out.cpu().float()
            print("grid=%3d N=%3d store_warps=%2d bulk=%d : %.1f cycles/MMA (issue loop %.1f)  ideal %.0f" % (
                grid, N, sw, bc, o[:, 0].mean() / n, o[:, 1].mean() / n, N / 2), flush=True)
