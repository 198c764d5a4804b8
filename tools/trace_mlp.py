"""Timeline of the MMA <-> epilogue hand-offs inside the fused MLP kernel (CTA 0), from clock64 stamps."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from neurecon_b200 import _lib
from conftest import build_neus
from bench_mlp import run
mode = sys.argv[1] if len(sys.argv) > 1 else "nablas"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 148 * 64 * 6
dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
x = (torch.rand(n, 3, device=dev) - 0.5) * 1.5
v = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
net = m.implicit_surface._umma_net(m.radiance_net)
prog = net.program(mode)
prog.debug_flags = flags
img = torch.zeros((n + 127) // 128 * 65536, dtype=torch.uint8, device=dev) if ("img" in mode or mode.startswith("radiance")) else None
run(net, prog, x, v, n, 2, img)
tr = torch.zeros(3, 2048, 4, dtype=torch.int64, device=dev)
lib = _lib.get_lib()
lib.nr_mlp_umma_set_trace(_lib.ptr(tr))
ms = run(net, prog, x, v, n, 1, img)
lib.nr_mlp_umma_set_trace(None)
torch.cuda.synchronize()
t = tr.cpu().numpy()
ev = []
for r in range(3):
    for row in t[r]:
        if row[0] != 0:
            ev.append((int(row[2]), r, int(row[0]), int(row[1]), int(row[3])))
ev.sort()
# keep the last launch only (clock restarts are monotonic; take events of the final 'pair' sequence)
t0 = ev[0][0]
names = {10: "mma: wait in_ready", 11: "mma: got in_ready", 12: "mma: issued+commit", 20: "epi: wait acc", 21: "epi: got acc",
         22: "epi: math done", 23: "epi: published", 31: "epi: values half done", 33: "epi: tangents 0,1 half done"}
print("mode", mode, "flags", flags, "ms", ms)
for c, r, e, st, pair in ev[:int(os.environ.get('NR_TRACE_N', '140'))]:
    print("%9d  region %d  pair %d  step %2d tile %d  %s" % (c - t0, r, pair, st // 2, st % 2, names.get(e, e)))
