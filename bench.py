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
TRAIN_BLOCK_TIMEOUT_S = 240
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


def _event_ms(fn, reps=5, inner=4, flush=None):
    """best-of-`reps` ms per call, `inner` calls per timing, the L2 flushed (a > 126 MB write) before each timing"""
    import torch
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return min(ts)


def hbm_rooflines(dev, hbm_peak, R=N_RAYS):
    """K2 / K3 (sampling, compositing) against the HBM roofline: ALGORITHMIC bytes per ray / CUDA-event
    time of the kernel alone on R rays of synthetic per-sample buffers (hundreds of MB: larger than the L2, which is
    flushed before every timing anyway) / the measured copy bandwidth."""
    import torch
    from neurecon_b200 import _lib
    from neurecon_b200.models.frameworks import neus
    from neurecon_b200.utils import rend_util
    lib = _lib.get_lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    f = dict(device=dev, generator=g)
    out = {}

    def put(name, ms, bytes_per_ray, rays, kernel):
        gbs = rays * bytes_per_ray / (ms * 1e-3) / 1e9
        out[name] = {"kernel": kernel, "ms": ms, "rays": rays, "bytes_per_ray": bytes_per_ray, "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / hbm_peak}

    with torch.no_grad():
        M = 128                                            # NeuS: sdf[128] + nablas[128,3] + radiance[127,3] + d_mid[127] in, 32 B out
        sdf, nab = torch.randn(R, M, **f) * 0.3, torch.randn(R, M, 3, **f)
        rad, dmid = torch.rand(R, M - 1, 3, **f), torch.rand(R, M - 1, **f).sort(-1).values
        s = torch.tensor([20.0], device=dev)
        put("neus_composite", _event_ms(lambda: neus._composite(sdf, nab, rad, dmid, s, False, True, False), flush=flush),
            4084 + 32, R, "neus_composite_staged_kernel")
        del sdf, nab, rad, dmid
        Mb, N = 64, 16                                     # sample_pdf: bins M + weights M-1 in, N out (det)
        bins, w = torch.rand(R, Mb, **f).sort(-1).values, torch.rand(R, Mb - 1, **f)
        put("sample_pdf_64_to_16", _event_ms(lambda: rend_util.sample_pdf(bins, w, N, det=True), flush=flush),
            4 * Mb + 4 * (Mb - 1) + 4 * N, R, "sample_pdf_rows_kernel")
        del bins, w
        Rv, Mv = R // 2, 192                               # VolSDF: 192 x (sdf, nabla, radiance, d) in
        sdfv, nabv = torch.randn(Rv, Mv, **f) * 0.3, torch.randn(Rv, Mv, 3, **f)
        radv, dv = torch.rand(Rv, Mv, 3, **f), torch.rand(Rv, Mv, **f).sort(-1).values
        al, be = torch.tensor([10.0], device=dev), torch.tensor([0.1], device=dev)
        o3, o1a, o1b, o3n = (torch.empty(Rv, 3, device=dev), torch.empty(Rv, device=dev), torch.empty(Rv, device=dev),
                             torch.empty(Rv, 3, device=dev))

        def fnv():
            _lib.check(lib.nr_volsdf_composite(_lib.ptr(sdfv), _lib.ptr(nabv), _lib.ptr(radv), _lib.ptr(dv), _lib.ptr(al),
                                               _lib.ptr(be), Rv, Mv, None, None, None, 0, 0, _lib.ptr(o3), _lib.ptr(o1a),
                                               _lib.ptr(o1b), _lib.ptr(o3n), None, None, None, _lib.stream_ptr(dev)), "volsdf_composite")
        put("volsdf_composite", _event_ms(fnv, flush=flush), 6144 + 32, Rv, "volsdf_composite_staged_kernel")
        del sdfv, nabv, radv, dv
        Mu = 96                                            # UNISURF: 96 x (logit, nabla, radiance, d) in
        lgu, nabu = torch.randn(R, Mu, **f) * 3.0, torch.randn(R, Mu, 3, **f)
        radu, du = torch.rand(R, Mu, 3, **f), torch.rand(R, Mu, **f).sort(-1).values
        p3, p1a, p1b, p3n = (torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev),
                             torch.empty(R, 3, device=dev))

        def fnu():
            _lib.check(lib.nr_unisurf_composite(_lib.ptr(lgu), _lib.ptr(nabu), _lib.ptr(radu), _lib.ptr(du), R, Mu, 0,
                                                _lib.ptr(p3), _lib.ptr(p1a), _lib.ptr(p1b), _lib.ptr(p3n), None, None,
                                                _lib.stream_ptr(dev)), "unisurf_composite")
        put("unisurf_composite", _event_ms(fnu, flush=flush), 3072 + 32, R, "unisurf_composite_staged_kernel")
        del lgu, nabu, radu, du
        # NeuS up-sampler: the five nr_neus_upsample_step launches of one render (merge + slopes + logistic CDF + inverse-CDF
        # draw, then the final emit), timed one by one between the (untimed) sdf evaluations of an analytic sphere.
        # Bytes per ray and step: state read 8 (m_cur + n_new), state written 8 (m_cur + n_new), 16 new depths + points
        # written (256 B), the final step writes points, mid depths and mid points of all 128 samples instead.
        Ru, n0, nf = min(R, 1 << 18), 64, 16
        o = torch.nn.functional.normalize(torch.randn(Ru, 3, **f), dim=-1) * 2.5
        dr = torch.nn.functional.normalize(-o + 0.1 * torch.randn(Ru, 3, **f), dim=-1)
        ff = dict(dtype=torch.float32, device=dev)
        dirs, near, far = torch.empty(Ru, 3, **ff), torch.empty(Ru, **ff), torch.empty(Ru, **ff)
        st = _lib.stream_ptr(dev)
        cap = n0 + 4 * nf
        step_bytes = [8 * 64 * 2 + 256, 8 * 80 * 2 + 256, 8 * 96 * 2 + 256, 8 * 112 * 2 + 256, 8 * 128 * 2 + 12 * 128 + 4 * 127 + 12 * 127]

        def one_render():
            d_new, pts_new = torch.empty(Ru, n0, **ff), torch.empty(Ru, n0, 3, **ff)
            _lib.check(lib.nr_neus_ray_setup(_lib.ptr(o), _lib.ptr(dr), Ru, 1.0, float("nan"), float("nan"), n0, _lib.ptr(dirs),
                                             _lib.ptr(near), _lib.ptr(far), _lib.ptr(d_new), _lib.ptr(pts_new), st), "ray_setup")
            d_buf, sdf_buf = torch.empty(Ru, cap, **ff), torch.empty(Ru, cap, **ff)
            pts_all, d_mid, pts_mid = torch.empty(Ru, cap, 3, **ff), torch.empty(Ru, cap - 1, **ff), torch.empty(Ru, cap - 1, 3, **ff)
            m_cur, n_new, total = 0, n0, 0.0
            for it in range(5):
                sdf_new = (pts_new.norm(dim=-1) - 0.5).contiguous()
                n_next = 0 if it == 4 else nf
                d_next, pts_next = torch.empty(Ru, max(n_next, 1), **ff), torch.empty(Ru, max(n_next, 1), 3, **ff)
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.nr_neus_upsample_step(
                    _lib.ptr(o), _lib.ptr(dirs), Ru, _lib.ptr(d_buf), _lib.ptr(sdf_buf), cap, m_cur, _lib.ptr(d_new),
                    _lib.ptr(sdf_new), n_new, it, n_next, None, _lib.ptr(d_next), _lib.ptr(pts_next), _lib.ptr(pts_all),
                    _lib.ptr(d_mid), _lib.ptr(pts_mid), None, None, st), "neus_upsample_step")
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
                m_cur += n_new
                d_new, pts_new, n_new = d_next, pts_next, n_next
            return total
        one_render()
        put("neus_upsample_5_steps", min(one_render() for _ in range(3)), sum(step_bytes), Ru, "neus_upsample_kernel x5")
        # iso-surface extraction: 512^3 sdf lattice of a sphere in HBM -> indexed mesh in HBM; algorithmic
        # bytes = the volume once + the mesh once ("rays" = lattice points)
        from neurecon_b200.utils import mesh_util
        Ng = 512
        gg = torch.linspace(-1, 1, Ng, device=dev)
        vol = torch.sqrt(gg[:, None, None] ** 2 + gg[None, :, None] ** 2 + gg[None, None, :] ** 2) - 0.6
        v, fc = mesh_util.marching_cubes(vol, 0.0, (2.0 / Ng,) * 3)
        ms = _event_ms(lambda: mesh_util.marching_cubes(vol, 0.0, (2.0 / Ng,) * 3), flush=flush)
        mc_bytes = 4 * Ng ** 3 + 12 * v.shape[0] + 12 * fc.shape[0]
        put("marching_cubes_512", ms, mc_bytes / Ng ** 3, Ng ** 3, "mc_count_kernel + 2 cub scans + mc_generate_kernel (incl. the one host read of the totals)")
        out["marching_cubes_512"].update(vertices=int(v.shape[0]), triangles=int(fc.shape[0]))
        del vol, v, fc
    return out


def extra_configs(dev, rank, world, timed, precision):
    """BASELINE.json configs 2 - 5 and two informational legs, all under this run's clock.
    config 2 / 3: VolSDF 1024-ray render + training iteration, UNISURF 2048-ray render (per GPU);
    config 4: VolSDF + NeRF++ (configs/volsdf_nerfpp_blended.yaml) full 576x768 view, the rays of ONE image split over
      the ranks by contiguous ranges (`dist_util.shard_range`; strong scaling, no collective);
    config 5: extract_surface's 512^3 lattice WITH nablas, x-planes split over the ranks.
    fp32 tier: the same NeuS render on the fp32 SIMT tier (the <= 1e-4 path), 16 384 rays."""
    import torch
    import neurecon_b200
    from neurecon_b200.models.frameworks import neus, volsdf
    from neurecon_b200.utils import dist_util, mesh_util, rend_util, synthetic
    out = {}
    with torch.no_grad():
        # ---- config 4 ----
        torch.manual_seed(0)
        mv = volsdf.VolSDF(**dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=0.01, use_nerfplusplus=True))
        synthetic.reseed_parameters(mv, seed=3)
        mv = mv.to(dev)
        c2w = synthetic.look_at_pose([1.6, 1.2, 1.0])[None].to(dev)
        intr = synthetic.pinhole_intrinsics(H, W)[None].to(dev)
        ro, rd, _ = rend_util.get_rays(c2w, intr, H, W, N_rays=-1)
        lo, hi = dist_util.shard_range(N_RAYS, rank, world)
        ro, rd = ro[0, lo:hi].contiguous(), rd[0, lo:hi].contiguous()
        kw = dict(calc_normal=True, detailed_output=False, perturb=False, rayschunk=65536, near=0.0, far=6.0,
                  obj_bounding_radius=3.0, max_upsample_steps=5, use_nerfplusplus=True, N_outside=32)
        volsdf.volume_render(ro, rd, mv, **kw)
        ms = timed(lambda: volsdf.volume_render(ro, rd, mv, **kw), 2) / 2
        out["config4_volsdf_nerfpp_576x768"] = {
            "value": N_RAYS / (ms * 1e-3), "unit": "rays/s", "ms_per_image": ms, "scaling": "strong",
            "workload": "VolSDF + NeRF++ (configs/volsdf_nerfpp_blended.yaml, beta 0.01 random init), one 576x768 view, "
                        "128 + 64 samples + 32 outside, <= 5 up-sample iterations, its %d rays split over %d rank(s)" % (N_RAYS, world)}
        del mv
    # ---- configs 2 and 3 (BASELINE.json): VolSDF 1024-ray render and training iteration, UNISURF 2048-ray render; every
    # rank does its own batch (weak scaling), eager (the VolSDF sampler's early exit reads one flag per iteration) ----
    from neurecon_b200.models.frameworks import unisurf
    from neurecon_b200.utils import train_util
    torch.manual_seed(0)
    mv = volsdf.VolSDF(**dict(synthetic.VOLSDF_MODEL_KWARGS, beta_init=0.01))
    synthetic.reseed_parameters(mv, seed=3)
    mv = mv.to(dev)
    o2, d2 = synthetic.make_rays(1024, shell_radius=3.0 / 1.1, jitter=0.1, seed=500 + rank)
    o2, d2 = o2.to(dev), d2.to(dev)
    kw2 = dict(near=0.0, far=6.0, obj_bounding_radius=3.0, max_upsample_steps=6, perturb=False)
    with torch.no_grad():
        volsdf.volume_render(o2, d2, mv, calc_normal=True, detailed_output=False, **kw2)
        ms_r = timed(lambda: volsdf.volume_render(o2, d2, mv, calc_normal=True, detailed_output=False, **kw2), 5) / 5
    opt2 = train_util.FusedAdam(mv.parameters(), lr=5e-4)
    tgt2 = torch.rand(1024, 3, device=dev)

    def volsdf_iteration():
        opt2.zero_grad(set_to_none=False)
        rgb, _, ret = volsdf.volume_render(o2, d2, mv, detailed_output=True, **dict(kw2, perturb=True))
        nn_ = ret["implicit_nablas"].norm(dim=-1)
        loss = (rgb - tgt2).abs().mean() + 0.1 * ((nn_ - 1.0) ** 2).mean()            # volsdf.py:597-621 (L1 + eikonal)
        loss.backward()
        dist_util.allreduce_gradients(mv.parameters())
        opt2.step()
    for _ in range(3):
        volsdf_iteration()
    ms_t = timed(volsdf_iteration, 5) / 5
    out["config2_volsdf_1024"] = {
        "render_ms": ms_r, "render_rays_per_s": world * 1024 / (ms_r * 1e-3), "train_step_ms": ms_t,
        "train_rays_per_s": world * 1024 / (ms_t * 1e-3), "scaling": "weak",
        "workload": "VolSDF (configs/volsdf.yaml, beta 0.01 random init), 1024 rays per GPU: error-bounded beta up-sampling render "
                    "(128 + 64 samples, <= 6 iterations), and one training iteration (render under autograd, L1 + eikonal, backward "
                    "incl. ln_beta, gradient all-reduce, Adam), eager"}
    del opt2
    # the same iteration as ONE CUDA graph: the sampler then runs sync-free (every ray through all its iterations)
    try:
        opt2g = train_util.FusedAdam(mv.parameters(), lr=5e-4, capturable=True)

        def volsdf_iteration_g(o, d, tgt):
            opt2g.zero_grad(set_to_none=False)
            rgb, _, ret = volsdf.volume_render(o, d, mv, detailed_output=True, **dict(kw2, perturb=True))
            nn_ = ret["implicit_nablas"].norm(dim=-1)
            loss = (rgb - tgt).abs().mean() + 0.1 * ((nn_ - 1.0) ** 2).mean()
            loss.backward()
            dist_util.allreduce_gradients(mv.parameters())
            opt2g.step()
            return loss.detach()
        gstep = train_util.CapturedStep(volsdf_iteration_g, (o2, d2, tgt2), optimizer=opt2g, warmup=3)
        for _ in range(2):
            gstep(o2, d2, tgt2)
        ms_g = timed(lambda: gstep(o2, d2, tgt2), 5) / 5
        out["config2_volsdf_1024"].update(train_step_graph_ms=ms_g, train_graph_rays_per_s=world * 1024 / (ms_g * 1e-3))
        del gstep, opt2g
    except Exception as e:      # informational leg
        out["config2_volsdf_1024"]["train_step_graph_error"] = repr(e)[:200]
    del mv
    torch.manual_seed(0)
    mu = unisurf.UNISURF(**synthetic.UNISURF_MODEL_KWARGS)
    synthetic.reseed_parameters(mu, seed=4)
    mu = mu.to(dev)
    o3, d3 = synthetic.make_rays(2048, shell_radius=3.0, jitter=0.25, seed=600 + rank)
    o3, d3 = o3[None].to(dev), d3[None].to(dev)
    with torch.no_grad():
        kw3 = dict(batched=True, calc_normal=True, detailed_output=False, perturb=False)
        unisurf.volume_render(o3, d3, mu, **kw3)
        ms_u = timed(lambda: unisurf.volume_render(o3, d3, mu, **kw3), 5) / 5
    out["config3_unisurf_2048"] = {
        "render_ms": ms_u, "render_rays_per_s": world * 2048 / (ms_u * 1e-3), "scaling": "weak",
        "workload": "UNISURF (configs/unisurf.yaml), 2048 rays per GPU: 256-step root finding + 8 secant steps + interval sampling "
                    "(64 + 32 samples), render"}
    del mu
    with torch.no_grad():
        # ---- config 5 ----
        m = build_model(1, dev)
        GN = 512
        plo, phi = dist_util.shard_range(GN, rank, world)
        # one untimed pass over the same planes first: the result buffers of this size come out of the caching allocator once
        mesh_util.query_sdf_grid(m.implicit_surface, N=GN, plane_range=(plo, phi), with_nablas=True)
        ms = timed(lambda: mesh_util.query_sdf_grid(m.implicit_surface, N=GN, plane_range=(plo, phi), with_nablas=True), 1)
        out["config5_grid_512_nablas"] = {
            "value": GN ** 3 / (ms * 1e-3), "unit": "queries/s", "ms": ms, "scaling": "strong",
            "workload": "extract_surface lattice 512^3 (mesh_util.py:82-111) with analytic nablas, x-planes split over %d rank(s), "
                        "results left in HBM" % world,
            "algorithmic_tflops": GN ** 3 * MFLOP_PER_QUERY_NABLA * 1e6 / (ms * 1e-3) / 1e12}
        # ---- fp32 tier ----
        if precision != "fp32":
            o, d = synthetic.make_rays(16384, shell_radius=2.5, jitter=0.1, seed=100 + rank)
            o, d = o.to(dev), d.to(dev)
            neurecon_b200.set_precision("fp32")
            try:
                kwn = dict(calc_normal=True, detailed_output=False, perturb=False, rayschunk=65536)
                neus.volume_render(o[:2048], d[:2048], m, **kwn)
                ms = timed(lambda: neus.volume_render(o, d, m, **kwn), 1)
            finally:
                neurecon_b200.set_precision(precision)
            out["neus_fp32_tier"] = {"value": world * 16384 / (ms * 1e-3), "unit": "rays/s", "rays_per_gpu": 16384,
                                     "workload": "the headline NeuS render on the fp32 SIMT tier (<= 1e-4 parity), 16384 rays per GPU"}
        # ---- split-precision tensor tier: the <= 1e-4 contract on tcgen05 (csrc/mlp_rev_split.cu) ----
        if precision != "fp16x2":
            o, d = synthetic.make_rays(N_RAYS, shell_radius=2.5, jitter=0.1, seed=100 + rank)
            o, d = o.to(dev), d.to(dev)
            neurecon_b200.set_precision("fp16x2")
            try:
                kwn = dict(calc_normal=True, detailed_output=False, perturb=False, rayschunk=65536)
                neus.volume_render(o[:65536], d[:65536], m, **kwn)
                ms = timed(lambda: neus.volume_render(o, d, m, **kwn), 2) / 2
            finally:
                neurecon_b200.set_precision(precision)
            out["neus_fp16x2_tier"] = {
                "value": world * N_RAYS / (ms * 1e-3), "unit": "rays/s", "rays_per_gpu": N_RAYS, "ms_per_image": ms,
                "workload": "the headline NeuS render (576x768 view per GPU) on the split-precision tensor tier: SDF net on (hi, lo) "
                            "fp16 operand pairs (3 products as 2 MMAs per k-step), radiance net on fp16; <= 1e-4 parity"}
    return out


def reference_cuda_eager(dev, n_rays=4096):
    """Informational: the reference's algorithm as plain PyTorch ops ON THE GPU -- the oracle
    port's torch fp32 eager path on cuda:0, TF32 off -- i.e. what running the reference on a B200 amounts to."""
    import torch
    from oracle import neus as oneus
    from neurecon_b200.utils import synthetic
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = build_model(1, "cpu")
    sd = {k: v.detach().clone().to(dev) for k, v in m.state_dict().items()}
    o, d = synthetic.make_rays(n_rays, shell_radius=2.5, jitter=0.1, seed=1)
    o, d = o.to(dev), d.to(dev)
    oneus.volume_render(o[:512], d[:512], sd, NEUS_CFG, calc_normal=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        oneus.volume_render(o, d, sd, NEUS_CFG, calc_normal=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    return {"value": n_rays / (ms * 1e-3), "unit": "rays/s", "sample": "%d rays of the headline workload" % n_rays,
            "what": "oracle port of the reference's PyTorch ops, eager, fp32 (TF32 off), on cuda:0"}


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
        "config": reference_config(),
        "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays of the 576x768 workload per step, oracle port (torch CPU fp32)" % CPU_SAMPLE_RAYS},
        "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOAD = ("NeuS volume_render, 576x768 = 442368 rays per GPU per step, configs/neus.yaml network "
            "(8x256 SDF MLP + 4x256 radiance MLP), 64 coarse + 4x16 up-sampled samples, calc_normal, "
            "perturb=False, random-init weights")


def workload_config(precision):
    return {"workload": WORKLOAD, "rays_per_step_per_gpu": N_RAYS, "rayschunk": 65536, "mlp_tier": precision,
            "l2_policy": "inputs and intermediates per step (>1 GB) exceed the 126 MB L2; no explicit flush"}


def reference_config():
    """The reference arm times a BOUNDED SAMPLE of the workload: one of its steps is CPU_SAMPLE_RAYS rays of the same
    view, and ms_per_step is the time of that sample (value = rays/s is what compares with the GPU arm)."""
    return {"workload": WORKLOAD + " -- reference arm: each step renders a %d-ray sample of that view on the host CPU "
                        "(oracle port of the reference's torch-CPU path, fp32, all host threads)" % CPU_SAMPLE_RAYS,
            "rays_per_step_per_gpu": CPU_SAMPLE_RAYS, "sample_rays_per_step": CPU_SAMPLE_RAYS, "full_workload_rays": N_RAYS,
            "rayschunk": 65536, "mlp_tier": "fp32", "l2_policy": "n/a (CPU)"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel the roofline times, from one `ncu --set full`
# capture of that launch (profiles/): reverse-mode kernel of the fp16 tier / tangent-tile kernel of the bf16 tier
NCU_DRAM_BYTES_PER_LAUNCH = {"fp16": 17855232 + 212988160, "bf16": 13689344}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("NEURECON_B200_PRECISION", None),
                    help="fp16 (default: fused tcgen05 MLP, fp16 operands, fp32 accumulate), fp16x2 (split-precision tensor tier), bf16, or fp32 (SIMT tier)")
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
                "profiles/r2_ncu_mlp_rev_final_raw.csv.gz): 17.9 MB read, 213.0 MB written = 8.4 MB of outputs + the softplus' "
                "scratch lines evicted before the backward sweep read them back (the rest are dropped with "
                "discard.global.L2 after the read: round 1 wrote 787 MB; a --metrics-only capture of the same launch, "
                "profiles/r2_mlp_rev_dram.txt, saw 13.3 + 115.8 MB); algorithmic 14.7 MB; 0.2 TB/s, off the critical path")
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
                "whole_step_frac": (value / world) * MFLOP_PER_RAY * 1e6 / 1e12 / pk["bf16"],
                # the launches above run right after the timed render loop, at its power-capped clocks: next to the burst
                # peak (a GEMM timed alone on an idle GPU) the figure against the peak of a GEMM run back to back for seconds
                "peak_sustained": pk["bf16_sustained"],
                "frac_of_sustained_peak": (achieved / pk["bf16_sustained"]) if pk["bf16_sustained"] else None}

    # ---- second half of BASELINE.json's metric: dense SDF-grid queries (extract_surface, mesh_util.py:82-111) --------
    # every rank evaluates the x-planes of its own 256^3 lattice (weak scaling, like the rays); sdf only, lattice
    # generated on the device, result left in HBM
    from neurecon_b200.utils import mesh_util
    GN = 256
    for _ in range(2):
        mesh_util.query_sdf_grid(model.implicit_surface, N=GN, plane_range=(0, GN))
    ms_grid = timed(lambda: mesh_util.query_sdf_grid(model.implicit_surface, N=GN, plane_range=(0, GN)), 3)
    sdf_qps = world * GN ** 3 * 3 / (ms_grid * 1e-3)

    extras = extra_configs(dev, rank, world, timed, precision)
    hbm = hbm_rooflines(dev, pk["hbm"]) if rank == 0 else None
    eager = None
    if rank == 0 and world == 1:
        try:
            eager = reference_cuda_eager(dev)
        except Exception as e:                       # informational leg: never takes the line down
            eager = {"error": repr(e)[:200]}

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
            "dtype": {"fp32": "f32", "fp16": "f16", "bf16": "bf16", "fp16x2": "f16x2"}[precision], "data": "synthetic",
            "config": workload_config(precision),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "sdf_queries_per_s": {"value": sdf_qps, "unit": "queries/s", "workload": "%d^3 lattice per GPU, sdf only" % GN,
                                  "frac_of_bf16_peak": sdf_qps / world * 0.918 * 1e6 / 1e12 / pk["bf16"]},
            "roofline_hbm": {"peak_gbs": pk["hbm"], "peak_source": pk["src"] + " copy bandwidth", "kernels": hbm},
            "reference_cuda_eager": eager,
        }
        line.update(extras)

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
    train_x2 = train_plain = None
    if precision == "fp16" and "error" not in train:
        train["workload"] += ("; fp16 tier: forward / reverse sweeps on (hi, lo) fp16 operand pairs -- every parameter gradient "
                              "within 1e-2 of the reference's (tests/test_gpu_train_golden.py) --, up-sampler on the plain fp16 kernels")
        # the same iteration in the fp16x2 tier (its up-sampler runs on the split-precision inference kernel too) ...
        import neurecon_b200
        neurecon_b200.set_precision("fp16x2")
        try:
            train_x2 = train_step_bench(dev, world, rank, timed)
            train_x2["workload"] += "; fp16x2 tier"
        except Exception as e:
            train_x2 = {"error": repr(e)[:300]}
        finally:
            neurecon_b200.set_precision(precision)
        # ... and with plain fp16 sweeps (NEURECON_B200_TRAIN_SPLIT=0): faster, but single gradient tensors are up to 1e-1 off
        os.environ["NEURECON_B200_TRAIN_SPLIT"] = "0"
        try:
            train_plain = train_step_bench(dev, world, rank, timed)
            train_plain["workload"] += "; plain fp16 sweeps (NEURECON_B200_TRAIN_SPLIT=0): no gradient parity (worst tensor 1e-1 off)"
        except Exception as e:
            train_plain = {"error": repr(e)[:300]}
        finally:
            os.environ.pop("NEURECON_B200_TRAIN_SPLIT", None)
    finished.set()
    watchdog.cancel()
    if rank == 0:
        line["train_step"] = train
        if train_x2 is not None:
            line["train_step_fp16x2"] = train_x2
        if train_plain is not None:
            line["train_step_plain_fp16"] = train_plain
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
