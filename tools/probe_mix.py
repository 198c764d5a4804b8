"""Which instruction classes of the epilogue overlap?  (csrc/devtools/probe_mix.cu)  Cycles per scheduler for 8 x A, 8 x B and
8 x A interleaved with 8 x B, 4 warps per scheduler (512 threads per SM)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.get_devtools()
grid, threads, iters = 148, 512, 2000
out = torch.zeros(4, dtype=torch.int32, device=dev); cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
names = ["none", "fma.f32", "fma.f32x2", "ex2", "lop3", "max.f32", "prmt", "cvt.f16x2", "fma.f16x2", "add.s32", "fma.f32x2 (3 operands)"]


def run(a, b):
    for _ in range(2):
        _lib.check(lib.nr_probe_mix(a, b, threads, iters, grid, _lib.ptr(out), _lib.ptr(cyc), _lib.stream_ptr(dev)), "probe_mix")
    torch.cuda.synchronize()
    return cyc.float().mean().item() / iters * 4 / (threads / 32)     # cycles per scheduler for one warp's 8 (+ 8) instructions


alone = {a: run(a, 0) for a in range(1, len(names))}
for a in range(1, len(names)):
    print("%-24s alone: %5.1f cycles per 8 warp instructions (%.2f per instruction)" % (names[a], alone[a], alone[a] / 8), flush=True)
for a, b in ((2, 3), (2, 4), (2, 5), (2, 6), (2, 7), (2, 1), (1, 3), (1, 4), (1, 5), (3, 4), (3, 5), (4, 5), (4, 6), (10, 4), (10, 3), (8, 4), (2, 9), (1, 9)):
    m = run(a, b)
    print("%-12s + %-12s mix %5.1f   (A %5.1f, B %5.1f, sum %5.1f, max %5.1f)" % (names[a], names[b], m, alone[a], alone[b],
                                                                                alone[a] + alone[b], max(alone[a], alone[b])), flush=True)
