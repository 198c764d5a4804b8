"""Cycles per 16-value chunk of mlp_rev_kernel's forward epilogue in isolation (csrc/devtools/probe_epi.cu): 16 warps per SM,
four per scheduler, as in the kernel.  Inside the kernel a warp needs ~1450 cycles per chunk (tools/trace_rev.py)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neurecon_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.get_devtools()
grid = 148
bias = (torch.rand(256, device=dev) - 0.5) * 0.2
scratch = torch.zeros(32 * 8192 * grid + 4096, dtype=torch.uint8, device=dev)
cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
names = ["shipped: math + codes + stores", "no operand stores (st.shared)", "no code stores (st.global.cg)", "no stores", "scalar FFMA",
         "math only", "no activation math", "quadratic log1p, cubic 1/(1+u)", "no bias FMA", "code stores with default policy",
         "one transcendental (tanh): full", "tanh, relu on the FMA pipe"]
iters = 2000
for v, nm in enumerate(names):
    for _ in range(2):
        _lib.check(lib.nr_probe_epi(v, iters, grid, _lib.ptr(bias), _lib.ptr(scratch), _lib.ptr(cyc), _lib.stream_ptr(dev)), "probe_epi")
    torch.cuda.synchronize()
    c = cyc.float().mean().item() / (4 * iters)
    print("variant %d %-36s %.0f cycles per chunk and warp (4 warps per scheduler: %.0f per chunk-slot)" % (v, nm, c, c / 4), flush=True)
for pollers in (1, 2, 3, 4):
    for _ in range(2):
        _lib.check(lib.nr_probe_epi(0, iters | (pollers << 20), grid, _lib.ptr(bias), _lib.ptr(scratch), _lib.ptr(cyc), _lib.stream_ptr(dev)), "probe_epi")
    torch.cuda.synchronize()
    c = cyc.float().mean().item() / (4 * iters)
    print("shipped, %d warps polling an mbarrier (try_wait loop): %.0f cycles per chunk and warp" % (pollers, c), flush=True)
for v, nm in ((20, "8 fat warps (168 registers), 16-column chunks"), (21, "8 fat warps, two chunks in flight")):
    for _ in range(2):
        _lib.check(lib.nr_probe_epi(v, iters, grid, _lib.ptr(bias), _lib.ptr(scratch), _lib.ptr(cyc), _lib.stream_ptr(dev)), "probe_epi")
    torch.cuda.synchronize()
    c = cyc.float().mean().item() / (8 * iters)
    print("variant %d %-44s %.0f cycles per chunk and warp (2 warps per scheduler: %.0f per chunk-slot)" % (v, nm, c, c / 2), flush=True)
