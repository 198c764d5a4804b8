"""Time of the split-precision SDF kernel (csrc/mlp_rev_split.cu, precision 'fp16x2'): sdf + normals (+ radiance operand image)
of n points, one launch per timing.  Usage: python tools/bench_split.py [n] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from soak_mlp import Runner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
r = Runner(n, torch.device("cuda:0"))
for mode in ("split:rev_img", "split:rev_sdf"):
    for _ in range(2):
        r.launch(mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r.launch(mode)
    e1.record()
    torch.cuda.synchronize()
    print("%-14s %.3f ms per %d points" % (mode, e0.elapsed_time(e1) / reps, n), flush=True)
