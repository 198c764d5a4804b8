#!/usr/bin/env python
"""Benchmark of the NeuS ray-marched SDF volume-rendering hot path (BASELINE.json metric:
rays/sec, NeuS 64+64 samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16]

One "step" = one full 576x768 NeuS render (442 368 synthetic rays, configs/neus.yaml network,
64 coarse + 4x16 up-sampled samples, calc_normal) per GPU; with N GPUs every rank renders its own
view (independent rays, no data-path collective) => weak scaling.  Prints ONE JSON line (rank 0).

--impl reference times the reference's CPU path (the torch-CPU oracle port under oracle/; the
reference itself is pure Python and cannot travel to the GPU box) on a bounded sample of the
same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 576, 768
N_RAYS = H * W
N_SAMPLES, N_IMPORTANCE = 64, 64
CPU_SAMPLE_RAYS = 1024
TRAIN_BLOCK_TIMEOUT_S = 120
# algorithmic MFLOP per NeuS ray (inference) and per SDF query with nabla
MFLOP_PER_RAY = 704.9
MFLOP_PER_QUERY_NABLA = 1.967
NEUS_CFG = dict(multires=6, multires_view=4, rad_multires=-1, skips=[4], D=8, D_rad=4, speed_factor=10.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(seed, device):
    import torch
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import synthetic
    torch.manual_seed(0)
    m = neus.NeuS(**synthetic.NEUS_MODEL_KWARGS)
    synthetic.reseed_parameters(m, seed=seed)
    return m.to(device)


def train_step_bench(dev, world, rank, timed, R=512, steps=10):
    """The training half of the target: one NeuS iteration of configs/neus.yaml per GPU -- 512 rays, volume_render under
    autograd (perturb=True), L1 + 0.1 eikonal + 1.0 mask BCE (neus.py:443-478), backward incl. the second-order path,
    one flat gradient all-reduce (train.py:124), Adam -- replayed as ONE CUDA graph (train_util.CapturedStep)."""
    import torch
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import dist_util, synthetic, train_util
    model = build_model(1, dev)
    opt = train_util.FusedAdam(model.parameters(), lr=5e-4, capturable=True)
    o, d = synthetic.make_rays(R, shell_radius=2.5, jitter=0.1, seed=300 + rank)
    g = torch.Generator().manual_seed(400 + rank)
    o, d, target = o.to(dev), d.to(dev), torch.rand(R, 3, generator=g).to(dev)
    mask = (torch.rand(R, generator=g) < 0.7).to(dev)

    def iteration(o, d, target, mask):
        opt.zero_grad(set_to_none=False)
        rgb, _, ret = neus.volume_render(o, d, model, detailed_output=True, perturb=True)
        losses = train_util.neus_losses(rgb, target, ret["implicit_nablas"], mask_volume=ret["mask_volume"], target_mask=mask,
                                        w_eikonal=0.1, w_mask=1.0)
        losses["total"].backward()
        dist_util.allreduce_gradients(model.parameters())
        opt.step()
        return losses["total"].detach()

    step = train_util.CapturedStep(iteration, (o, d, target, mask), optimizer=opt, warmup=3)
    for _ in range(2):
        step(o, d, target, mask)
    ms = timed(lambda: step(o, d, target, mask), steps) / steps
    return {"ms_per_step": ms, "value": world * R / (ms * 1e-3), "unit": "rays/s", "rays_per_gpu": R,
            "workload": "NeuS training iteration (configs/neus.yaml: 512 rays per GPU, 64 + 4x16 samples, L1 + eikonal + mask "
                        "loss, backward with the second-order path, flat gradient all-reduce, Adam), one CUDA graph per iteration",
            "algorithmic_tflops_per_gpu": R * 1.85e9 / (ms * 1e-3) / 1e12}


def cpu_reference_rate(n_rays, reps, seed=1):
    """rays/s of the oracle port (torch CPU, fp32, all host threads) on `n_rays` rays of the workload."""
    import torch
    from oracle import neus as oneus
    from neurecon_b200.utils import synthetic
    # all host threads the process may use (torchrun pins OMP_NUM_THREADS=1 unless told otherwise)
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    m = build_model(seed, "cpu")
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    o, d = synthetic.make_rays(n_rays, shell_radius=2.5, jitter=0.1, seed=seed)
    cores = torch.get_num_threads()
    oneus.volume_render(o[:64], d[:64], sd, NEUS_CFG, calc_normal=True)  # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oneus.volume_render(o, d, sd, NEUS_CFG, calc_normal=True)
        ts.append(time.perf_counter() - t0)
    return n_rays / (sum(ts) / len(ts)), cores, ts


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    ts_all = []
    rate, cores, _ = cpu_reference_rate(CPU_SAMPLE_RAYS, 1)  # warm
    t0 = time.perf_counter()
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_rate(CPU_SAMPLE_RAYS, 1)
    rate, cores, ts = cpu_reference_rate(CPU_SAMPLE_RAYS, args.steps)
    ms = 1e3 * sum(ts) / len(ts)
    line = {
        "impl": "reference", "metric": "rays/sec (NeuS 64+64 samples)", "value": rate, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config("fp32"),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays of the 576x768 workload per step, oracle port (torch CPU fp32)" % CPU_SAMPLE_RAYS},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(precision):
    return {"workload": "NeuS volume_render, 576x768 = 442368 rays per GPU per step, configs/neus.yaml network "
                        "(8x256 SDF MLP + 4x256 radiance MLP), 64 coarse + 4x16 up-sampled samples, calc_normal, "
                        "perturb=False, random-init weights",
            "rays_per_step_per_gpu": N_RAYS, "rayschunk": 65536, "mlp_tier": precision,
            "l2_policy": "inputs and intermediates per step (>1 GB) exceed the 126 MB L2; no explicit flush"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel the roofline times, from one `ncu --set full`
# capture of that launch (profiles/): reverse-mode kernel of the fp16 tier / tangent-tile kernel of the bf16 tier
NCU_DRAM_BYTES_PER_LAUNCH = {"fp16": 46644736 + 787309056, "bf16": 13689344}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NEURECON_B200_PRECISION", None),
                    help="fp16 (default: fused tcgen05 MLP, fp16 operands, fp32 accumulate), bf16, or fp32 (SIMT tier)")
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import neurecon_b200
    from neurecon_b200 import _lib
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: with NCCL_DEBUG=VERSION in the environment NCCL printf()s its banner
        # ("NCCL version ...") to stdout, whatever NCCL_DEBUG_FILE says; other debug levels go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    if args.precision:
        neurecon_b200.set_precision(args.precision)
    precision = neurecon_b200.get_precision()
    lib = _lib.get_lib()
    n_rays = args.rays

    model = build_model(1, dev)
    o_host, d_host = synthetic.make_rays(n_rays, shell_radius=2.5, jitter=0.1, seed=100 + rank)
    o_pin, d_pin = o_host.pin_memory(), d_host.pin_memory()
    o_dev, d_dev = o_host.to(dev), d_host.to(dev)
    kw = dict(calc_normal=True, detailed_output=False, perturb=False, rayschunk=65536)

    def step_resident():
        with torch.no_grad():
            return neus.volume_render(o_dev, d_dev, model, **kw)

    out_pin = {k: torch.empty(s, dtype=torch.float32).pin_memory()
               for k, s in (("rgb", (n_rays, 3)), ("depth_volume", (n_rays,)), ("mask_volume", (n_rays,)),
                            ("normals_volume", (n_rays, 3)))}

    def step_e2e():
        with torch.no_grad():
            o = o_pin.to(dev, non_blocking=True)
            d = d_pin.to(dev, non_blocking=True)
            _, _, ret = neus.volume_render(o, d, model, **kw)
            for k, buf in out_pin.items():
                buf.copy_(ret[k], non_blocking=True)
        return ret

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.nr_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.nr_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * n_rays * args.steps / (ms * 1e-3)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = world * n_rays * args.steps / (ms_e2e * 1e-3)
    h2d = 2 * n_rays * 3 * 4
    d2h = sum(v.numel() * 4 for v in out_pin.values())

    # ---- roofline of the dominant kernel: the SDF MLP with analytic normals -------------------
    pk = peaks()
    n_pts = 65536 * 8
    pts = (torch.rand(n_pts, 3, device=dev) - 0.5) * 1.5
    with torch.no_grad():
        for _ in range(3):
            model.implicit_surface._run(pts, want_nablas=True, want_feat=False)
        torch.cuda.synchronize()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lk0 = lib.nr_launch_count()
        e0.record()
        for _ in range(reps):
            model.implicit_surface._run(pts, want_nablas=True, want_feat=False)
        e1.record()
        torch.cuda.synchronize()
        k_launches = (lib.nr_launch_count() - lk0) // reps
    k_ms = e0.elapsed_time(e1) / reps
    achieved = n_pts * MFLOP_PER_QUERY_NABLA * 1e6 / (k_ms * 1e-3) / 1e12
    if precision == "fp16":
        kname = ("mlp_rev_kernel (fused tcgen05, fp16 operands, reverse-mode normals): sdf + analytic nabla of %d points, "
                 "%d launch(es), %.3f ms; algorithmic 1.967 MFLOP/query = what the tensor pipe executes (forward sweep + "
                 "backward sweep)" % (n_pts, k_launches, k_ms))
        tsrc = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch (524288 points, "
                "profiles/mlp_rev_r1_ncu_summary.txt): 46.6 MB read (the points and the weights once), 787 MB written, of "
                "which 8.4 MB are outputs and the rest L2 write-backs of the 76 MB softplus' scratch (1.07 GB stored into "
                "it per launch, L2 hit rate 95 %); algorithmic 14.7 MB")
    elif precision == "bf16":
        kname = ("mlp_umma_kernel (fused tcgen05, bf16 operands, forward-mode tangent tiles): sdf + analytic nabla of %d "
                 "points, %d launch(es), %.3f ms; algorithmic 1.967 MFLOP/query (the tangents execute 4.2 MFLOP/query)"
                 % (n_pts, k_launches, k_ms))
        tsrc = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch (524288 points, "
                "profiles/mlp_umma_r1_ncu_524288.txt): 13.7 MB read, 0 written inside the kernel window; algorithmic 14.7 MB")
    else:
        kname = "gemm_kernel (fp32 SIMT tier): sdf + analytic nabla of %d points, %d launch(es), %.3f ms" % (
            n_pts, k_launches, k_ms)
        tsrc = None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16"], "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(precision),
                "traffic_source": tsrc, "peak_source": pk["src"] + " bf16 burst", "kernel": kname,
                "whole_step_frac": (value / world) * MFLOP_PER_RAY * 1e6 / 1e12 / pk["bf16"]}

    # ---- second half of BASELINE.json's metric: dense SDF-grid queries (extract_surface, mesh_util.py:82-111) --------
    # every rank evaluates the x-planes of its own 256^3 lattice (weak scaling, like the rays); sdf only, lattice
    # generated on the device, result left in HBM
    from neurecon_b200.utils import mesh_util
    GN = 256
    for _ in range(2):
        mesh_util.query_sdf_grid(model.implicit_surface, N=GN, plane_range=(0, GN))
    ms_grid = timed(lambda: mesh_util.query_sdf_grid(model.implicit_surface, N=GN, plane_range=(0, GN)), 3)
    sdf_qps = world * GN ** 3 * 3 / (ms_grid * 1e-3)

    line = None
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, ts = cpu_reference_rate(CPU_SAMPLE_RAYS, 3)
            cpu_baseline = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                            "sample": "%d rays of the same workload, oracle port (torch CPU fp32), 3 reps" % CPU_SAMPLE_RAYS}
        line = {
            "metric": "rays/sec (NeuS 64+64 samples)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16"}[precision], "data": "synthetic",
            "config": workload_config(precision),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "sdf_queries_per_s": {"value": sdf_qps, "unit": "queries/s", "workload": "%d^3 lattice per GPU, sdf only" % GN,
                                  "frac_of_bf16_peak": sdf_qps / world * 0.918 * 1e6 / 1e12 / pk["bf16"]},
        }

    # ---- secondary: the 512-ray training iteration.  The headline numbers above are complete; a watchdog prints them and
    # ends the process if this block (a CUDA graph with an NCCL all-reduce inside at N > 1) should ever fail to return ----
    finished = threading.Event()

    def bail():
        if finished.is_set():
            return
        if rank == 0:
            line["train_step"] = {"error": "the training block did not finish within %d s" % TRAIN_BLOCK_TIMEOUT_S}
            print(json.dumps(line), flush=True)
        os._exit(0)

    watchdog = threading.Timer(TRAIN_BLOCK_TIMEOUT_S, bail)
    watchdog.daemon = True
    watchdog.start()
    try:
        train = train_step_bench(dev, world, rank, timed)
    except Exception as e:  # reported in the line, the render numbers stand
        train = {"error": repr(e)[:300]}
    finished.set()
    watchdog.cancel()
    if rank == 0:
        line["train_step"] = train
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
