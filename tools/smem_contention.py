"""Do the epilogue's shared-memory stores get bandwidth while tcgen05.mma streams its operands from shared memory?
Store throughput of one warp (of W store warps) with and without a concurrent back-to-back MMA stream."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.get_devtools()
src = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
grid = 148
out = torch.zeros(3 * grid, dtype=torch.int64, device=dev)
for N in (128, 256):
    for sw in (1, 4, 8, 16):
        res = {}
        for label, n in (("no MMA", -131072), ("MMA", 2048)):
            for _ in range(2):
                out.zero_()
                _lib.check(lib.nr_bench_umma(N, n, sw, 0, _lib.ptr(src), grid, _lib.ptr(out), _lib.stream_ptr(dev)), "bench_umma")
            torch.cuda.synchronize()
            o = out.cpu().double()
            cyc = o[0:2 * grid:2].mean().item()
            stores = o[2 * grid:].mean().item()
            res[label] = (cyc, stores)
        (c0, s0), (c1, s1) = res["no MMA"], res["MMA"]
        print("N=%3d store warps=%2d: alone %.1f B/clk/SM of stores; under MMA %.1f B/clk/SM (%.0f%%), MMA %.1f cycles each" % (
            N, sw, s0 * 512 * sw / c0, s1 * 512 * sw / c1, 100 * (s1 / c1) / (s0 / c0), c1 / 2048), flush=True)
