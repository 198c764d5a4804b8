"""Launch the fused MLP kernel a few times in one mode (for ncu).  Usage: prof_one.py MODE N [FLAGS]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from conftest import build_neus
from bench_mlp import run
mode, n = sys.argv[1], int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
m = build_neus(seed=1, device=dev)
x = (torch.rand(n, 3, device=dev) - 0.5) * 1.5
v = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
net = m.implicit_surface._umma_net(m.radiance_net)
prog = net.program(mode)
prog.debug_flags = flags
print(mode, n, flags, run(net, prog, x, v, n, 2), "ms")
