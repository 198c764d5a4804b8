"""Per-SM issue rate of the epilogue's special-function / conversion ops (lanes per clock per SM)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.get_devtools()
out = torch.zeros(4, device=dev); cyc = torch.zeros(148, dtype=torch.int64, device=dev)
names = ["ex2.approx", "rcp.approx", "cvt.f16x2.f32", "fma.f32", "cvt.bf16x2.f32", "lg2.approx", "fma.f32x2 (4 instr per iteration; lanes = instr lanes x2)",
         "ex2.f16x2 (2 values per lane)", "tanh.f16x2 (2 values per lane)", "tanh.f32", "fma.f16x2 (2 values per lane)", "max.f16x2"]
for threads in (512, 1024):
    for op, nm in enumerate(names):
        iters = 2000
        for _ in range(2):
            _lib.check(lib.nr_probe_alu(op, threads, iters, 148, _lib.ptr(out), _lib.ptr(cyc), _lib.stream_ptr(dev)), "probe")
        torch.cuda.synchronize()
        c = cyc.float().mean().item()
        print("threads=%4d %-16s %.2f lanes/clk/SM  (%.2f cycles per warp instruction per SM)" % (
            threads, nm, threads * iters * 8 / c, c / (threads / 32 * iters * 8)), flush=True)

for warps in (1, 4, 8, 16):
    iters = 2000
    for _ in range(2):
        _lib.check(lib.nr_bench_ldtm(warps, iters, 148, _lib.ptr(cyc), _lib.ptr(out), _lib.stream_ptr(dev)), "ldtm")
    torch.cuda.synchronize()
    c = cyc.float().mean().item()
    print("tcgen05.ld.32x32b.x16: %2d warps: %.1f B/clk/SM  (%.1f cycles per 2 KB load per warp)" % (
        warps, warps * iters * 4 * 2048 / c, c / (iters * 4)), flush=True)
