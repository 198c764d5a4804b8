import os, sys, torch
sys.path.insert(0, "/root/repo")
from neurecon_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.get_devtools()
src = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)
out = torch.zeros(1, 2, dtype=torch.int64, device=dev)
for N in (128, 32):
    for n in (4, 8, 12, 16, 24, 32, 48, 64, 128):
        for _ in range(3):
            _lib.check(lib.nr_bench_umma(N, n, 0, 0, _lib.ptr(src), 1, _lib.ptr(out), _lib.stream_ptr(dev)), "b")
        torch.cuda.synchronize()
        o = out.cpu()
        print("N=%d n_mmas=%3d: issue loop %5d cycles, completion %5d cycles" % (N, n, o[0, 1], o[0, 0]))
