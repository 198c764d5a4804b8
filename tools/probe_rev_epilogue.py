"""Where does the forward epilogue of mlp_rev_kernel spend its time?  Times the reverse-mode kernel of the TEST TWIN library
(libneurecon_b200_inject.so carries the probes: debug_flags 8 = no forward epilogue, 16 = no operand stores, 32 = no activation
math, 256 = no TMEM loads, 1 = no scratch stores, 4 = no backward epilogue).  Usage: python tools/probe_rev_epilogue.py [n]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from neurecon_b200 import build as nr_build  # noqa: E402
from soak_mlp import Runner, load_kernels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
r = Runner(n, torch.device("cuda:0"), load_kernels(nr_build.INJECT_LIB_PATH))
for flags in (0, 8, 16, 32, 48, 256, 272, 288, 304, 1, 4, 12, 0):
    for _ in range(2):
        r.launch("rev", flags)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r.launch("rev", flags)
    e1.record()
    torch.cuda.synchronize()
    print("flags %4d  %.3f ms" % (flags, e0.elapsed_time(e1) / 5), flush=True)
