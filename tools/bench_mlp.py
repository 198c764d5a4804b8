"""Micro-benchmark of the fused tcgen05 MLP kernel (one launch per timing), with the profiling
switches of nr_umma_program_t.debug_flags.  Usage: python tools/bench_mlp.py [n_points] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neurecon_b200  # noqa: E402
from neurecon_b200 import _lib  # noqa: E402
from neurecon_b200._lib import C  # noqa: E402
from conftest import build_neus  # noqa: E402


def run(net, prog, x, v, n, reps, img=None):
    lib = _lib.get_lib()
    dev = x.device
    f = dict(dtype=torch.float32, device=dev)
    sdf, nab, rgb = torch.empty(n, **f), torch.empty(n, 3, **f), torch.empty(n, 3, **f)

    ws = torch.empty(max(int(lib.nr_mlp_umma_reverse_workspace(C.byref(prog), n)), 16), dtype=torch.uint8, device=dev) \
        if prog.reverse else None

    def go():
        if prog.reverse:
            _lib.check(lib.nr_mlp_umma_reverse(C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias),
                                               net.bias.numel(), _lib.ptr(x), n, _lib.ptr(sdf), _lib.ptr(nab), None, 256,
                                               _lib.ptr(img), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "rev")
            return
        if getattr(prog, "_pair", False):
            _lib.check(lib.nr_mlp_umma2_forward(C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias),
                                                net.bias.numel(), _lib.ptr(x), n, _lib.ptr(sdf), _lib.ptr(nab), None, 256,
                                                _lib.ptr(img), _lib.stream_ptr(dev)), "umma2")
            return
        _lib.check(lib.nr_mlp_umma_forward(C.byref(prog), _lib.ptr(net.image), net.image.numel() * 2, _lib.ptr(net.bias),
                                           net.bias.numel(), _lib.ptr(x), _lib.ptr(v), n, _lib.ptr(sdf), _lib.ptr(nab),
                                           None, 256, _lib.ptr(rgb), None, _lib.ptr(img), _lib.stream_ptr(dev)), "umma")
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    m = build_neus(seed=1, device=dev)
    x = (torch.rand(n, 3, device=dev) - 0.5) * 1.5
    v = torch.nn.functional.normalize(torch.randn(n, 3, device=dev), dim=-1)
    net = m.implicit_surface._umma_net(m.radiance_net)
    img = torch.zeros((n + 127) // 128 * 65536, dtype=torch.uint8, device=dev)
    modes = os.environ.get("NR_MODES")
    for mode, mflop in (("sdf", 1.049 * 459008 / 524544), ("nablas", 1.967), ("rev", 1.967), ("rev_img", 1.967),
                        ("fused", 1.967 + 0.543),
                        ("nablas_img", 1.967), ("radiance", 0.543),
                        ("nablas_imgf", 1.967 + 0.1316), ("radiancef", 0.543 - 0.1316),
                        ("pair:sdf", 1.049 * 459008 / 524544), ("pair:nablas", 1.967), ("pair:nablas_imgf", 1.967 + 0.1316)):
        if modes and mode not in modes.split(","):
            continue
        for flags in [int(f) for f in os.environ.get("NR_FLAGS", "0,2").split(",")]:
            pair = mode.startswith("pair:")
            prog = net.program(mode[5:] if pair else mode, pair=pair)
            prog._pair = pair
            prog.debug_flags = flags
            ms = run(net, prog, x, v, n, reps, img if "img" in mode or mode.startswith("radiance") else None)
            print("mode=%-15s flags=%d  %8.3f ms  %7.1f Mpts/s  %6.1f algorithmic TFLOP/s" % (
                mode, flags, ms, n / ms / 1e3, n * mflop * 1e6 / (ms * 1e-3) / 1e12), flush=True)


if __name__ == "__main__":
    main()
