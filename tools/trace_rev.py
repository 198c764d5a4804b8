"""Timeline of the MMA <-> epilogue hand-offs inside mlp_rev_kernel (CTA 0, its second tile pair), from clock64 stamps of the
TEST TWIN library (libneurecon_b200_inject.so).  Usage: python tools/trace_rev.py [debug_flags] [pair_index]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from neurecon_b200 import _lib  # noqa: E402
from neurecon_b200._lib import C  # noqa: E402
from neurecon_b200 import build as nr_build  # noqa: E402
from soak_mlp import Runner, load_kernels  # noqa: E402

flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
sms = torch.cuda.get_device_properties(dev).multi_processor_count
n = sms * 256 * 4
lib = load_kernels(nr_build.INJECT_LIB_PATH)
lib.nr_mlp_rev_set_trace.restype, lib.nr_mlp_rev_set_trace.argtypes = C.c_int, [C.c_void_p]
r = Runner(n, dev, lib)
for _ in range(2):
    r.launch("rev", flags)
torch.cuda.synchronize()
tr = torch.zeros(8, 1024, 4, dtype=torch.int64, device=dev)
lib.nr_mlp_rev_set_trace(_lib.ptr(tr))
r.launch("rev", flags)
torch.cuda.synchronize()
lib.nr_mlp_rev_set_trace(None)
t = tr.cpu().numpy()
names = {11: "mma: got in_ready", 13: "mma: first chunk landed", 14: "mma: last chunk landed", 12: "mma: issued + commit",
         21: "epi: got acc", 24: "epi: first TMEM chunk in registers", 25: "epi: two chunks done", 22: "epi: done", 23: "epi: published"}
who = {0: "mma w1", 1: "mma w3", 2: "epi e0 ", 3: "epi e3 ", 4: "epi e8 ", 5: "epi e15", 6: "epi e5 "}
ev = []
for reg in range(8):
    for row in t[reg]:
        if row[0] != 0 and (row[3] == which * sms or (os.environ.get('NR_TRACE_TWO') and row[3] == (which + 1) * sms)):
            ev.append((int(row[2]), reg, int(row[0]), int(row[1])))
ev.sort()
t0 = ev[0][0]
print("flags", flags, "pair", which, "events", len(ev))
for c, reg, e, st in ev:
    print("%8d  %s  step %2d tile %d  %s" % (c - t0, who[reg], st // 2, st % 2, names.get(e, e)))
