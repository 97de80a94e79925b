#!/usr/bin/env python3
"""bench.py -- headline benchmark of the raw_ngp hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--repeats R] [--no-extras]

ours:       one JSON line.  The headline workload is BASELINE.json configs[1] -- synthetic NeRF training step (bound 1,
            cascade 1, grid 128^3, 4096 rays per GPU, max_steps 1024, fp16 hash table, 64-wide MLPs) through raw_ngp_b200
            (libngp_b200.so); under --gpus N every rank runs it on its own rays (weak scaling, peer-memory gradient
            exchange).  The timed region of EXACTLY K steps is repeated R times (each bracketed by barrier + synchronize,
            CUDA events, max over ranks) and the MEDIAN region is reported, so that one scheduling hiccup in a ~15 ms
            region does not move the number.
              value      training rays/s over all ranks, inputs resident in HBM
              e2e        the same step fed from pinned host buffers, loss read back each step
              roofline   the slowest kernel of the step, timed with CUDA events around its launches inside the step:
                         `frac` = compulsory HBM bytes / time / measured HBM peak, `l2` = corner-row gathers or reductions
                         / time / the ceiling measured in the same run with ngp_diag_l2_rate (csrc/diag.cu)
              grid_encode  configs[0]: GridEncoder fwd / bwd on 2^18 points (random and ray-coherent; fp16 / bf16 / fp32
                         tables; with input gradients), each against HBM and L2 ceilings, the reference's kernels beside
              configs4   the data-parallel BARF step (8192 rays per GPU, refined poses of 100 cameras, se3 all-reduce) at
                         the same N -- the configuration BASELINE names for 2 / 4 / 8 GPUs, also run at N = 1
              configs2   the light-stage step (N = 1 only); configs3: one 1920x1080 frame, ray tiles over the N ranks
              cpu_baseline  the PyTorch-on-CPU port of the same step (oracle/cpu_pipeline.py) on a bounded ray sample
reference:  the same step in the CPU port (the reference has no CPU path of its own; kind "port"), rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
CPU_SAMPLE_RAYS = 96
WORKLOAD = ("configs[1] NeRF training step (bound 1, cascade 1, grid 128^3, max_steps 1024, fp16 hash grid L16 F2 T2^19 + "
            "64-wide MLPs)")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed regions (B200_PROFILING.md recipe).  start() returns
    only after the first sample has arrived, so the sampler's start-up never falls into a timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self, wait_s=5.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < wait_s:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        return len(self.lines)

    def stop(self, first=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# scenes
# ----------------------------------------------------------------------------------------------------------------------
def build_model(device, **kw):
    """NeRFNetwork with the ball-shaped occupancy grid of SURVEY 8(d) config 2 (density 50 inside |x| < 0.5)."""
    from raw_ngp_b200 import raymarching, synthetic
    from raw_ngp_b200.nerf import NeRFNetwork, default_opt
    torch.manual_seed(0)
    cfg = dict(bound=1, grid_size=128, max_steps=1024, dt_gamma=0, T_thresh=1e-8, min_near=0.05, fp16=True, density_thresh=10,
               hashmap_size=19, hashgrid_resolution=2048)
    cfg.update(kw)
    model = NeRFNetwork(default_opt(**cfg)).to(device)
    grid = synthetic.ball_density_grid(H=cfg["grid_size"], cascade=model.cascade, bound=float(model.bound), radius=0.5, sigma=50.0).to(device)
    model.density_grid.copy_(grid)
    thresh = min(grid.clamp(min=0).mean().item(), 10.0)
    model.density_bitfield = raymarching.packbits(model.density_grid, thresh, model.density_bitfield)
    model.mean_density = grid.clamp(min=0).mean().item()
    return model


def build_scene(device, rank, seed=2, n_rays=RAYS_PER_GPU):
    from raw_ngp_b200 import synthetic
    model = build_model(device)
    o, d = synthetic.sphere_rays(n_rays, seed=seed + 1000 * rank)
    g = torch.Generator().manual_seed(7 + rank)
    target = torch.rand(n_rays, 3, generator=g)
    return model, o, d, target


def _median(v):
    v = sorted(v)
    return v[len(v) // 2]


class Timer:
    """R regions of exactly K calls of fn, each bracketed by barrier + synchronize and timed with CUDA events on the current
    stream; the per-region times are MAX-reduced over the ranks; the median region is the result."""

    def __init__(self, world, device):
        self.world, self.device = world, device

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def regions(self, fn, K, R, warmup=0):
        for _ in range(warmup):
            fn()
        ms = []
        for _ in range(R):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                fn()
            e1.record()
            self.barrier()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor(ms, device=self.device, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.tolist()
        return {"median": _median(ms), "min": min(ms), "max": max(ms), "regions": [round(v, 4) for v in ms]}


def time_kernel(fn, iters=20, warm=3, repeats=5):
    """median over `repeats` of the mean launch time (ms) of `iters` back-to-back calls (CUDA events, current stream)"""
    for _ in range(warm):
        fn()
    st = torch.cuda.current_stream()
    out = []
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(iters):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / iters)
    return _median(out)


# ----------------------------------------------------------------------------------------------------------------------
# L2 ceilings (measured in this run) and per-kernel byte tables
# ----------------------------------------------------------------------------------------------------------------------
def l2_ceilings(device):
    """csrc/diag.cu on this GPU: random 4-byte row gathers out of an L2-resident 32 MiB table, random red.add.f16x2 and
    red.add.v2.f16x2 into it (2048 threads per SM, 8 independent operations per thread in flight)."""
    from raw_ngp_b200 import _lib
    n_rows = 1 << 23
    table = torch.zeros(n_rows, dtype=torch.int32, device=device)
    sink = torch.zeros(1, dtype=torch.int32, device=device)
    blocks, rounds = 148 * 4, 64

    def rate(mode):
        ops = blocks * 512 * rounds * (8 if mode < 2 else 4)
        ms = time_kernel(lambda: _lib.call("ngp_diag_l2_rate", _lib.ptr(table), n_rows, blocks, rounds, mode, _lib.ptr(sink), _lib.stream()),
                         iters=5, warm=2, repeats=3)
        return ops / (ms * 1e-3) / 1e9
    g, r1, r2 = rate(0), rate(1), rate(2)
    return {"gather_rows_G_per_s": g, "red_f16x2_G_per_s": r1, "red_v2_f16x2_G_per_s": r2,
            "how": "ngp_diag_l2_rate (csrc/diag.cu): random 4-byte rows of an L2-resident 32 MiB table, 148 x 4 CTAs x 512 threads, 8 independent operations per thread in flight"}


# algorithmic bytes per sample of the step's kernels.  `hbm` = compulsory HBM traffic (inputs, saved
# activations, outputs); `l2_rows` = table rows gathered / reduced per sample (L2 traffic by design: the 23 MiB table and its
# gradient stay resident in the 126 MB L2)
STEP_KERNEL_BYTES = {
    # xyz + dirs in; enc 64 + hidden 4 x 128 + in2 64 saved (fp16); sigma 4 + rgb 12 out
    "ngp_field_forward_full": {"hbm": 12 + 12 + 64 + 4 * 128 + 64 + 4 + 12, "l2_rows": 128, "l2_kind": "gather"},
    # xyz, d sigma, sigma, d rgb, rgb in; saved enc 64 + in2 64 + hidden 4 x 128 in
    "ngp_field_backward_full": {"hbm": 12 + 4 + 4 + 12 + 12 + 64 + 4 * 128 + 64, "l2_rows": 128, "l2_kind": "reduce"},
    "ngp_march_rays_train_write": {"hbm": 4 + 32, "l2_rows": 0},
    "ngp_composite_train_mse": {"hbm": 2 * 24 + 16, "l2_rows": 0},
}


def kernel_roofline(name, ms, M, hbm_peak, ceil):
    b = STEP_KERNEL_BYTES[name]
    t = ms * 1e-3
    out = {"ms_per_launch": ms, "hbm_bytes_per_sample": b["hbm"], "hbm_GBps": b["hbm"] * M / t / 1e9,
           "hbm_frac": b["hbm"] * M / t / 1e9 / hbm_peak}
    if b["l2_rows"]:
        rows = b["l2_rows"] * M / t / 1e9
        c = ceil["gather_rows_G_per_s"] if b["l2_kind"] == "gather" else 2 * ceil["red_v2_f16x2_G_per_s"]
        out.update({"l2_rows_per_sample": b["l2_rows"], "l2_G_rows_per_s": rows, "l2_ceiling_G_rows_per_s": c, "l2_frac": rows / c,
                    "l2_kind": b["l2_kind"] + (" (ceiling: all-miss random gathers; L1 hits put the kernel above it)" if b["l2_kind"] == "gather"
                                               else " (ceiling: paired red.v2.f16x2, 2 rows per operation; rows counted before warp aggregation)")})
    return out


# ----------------------------------------------------------------------------------------------------------------------
# configs[0]: encoder micro-benchmark
# ----------------------------------------------------------------------------------------------------------------------
def encoder_micro(device, hbm_peak, ceil=None, full=True):
    """BASELINE configs[0] / SURVEY 8(d) config 1: GridEncoder fwd / bwd on 2^18 points (L = 16, F = 2, T = 2^19, 16 -> 2048).
    Inputs: "random" = uniform in the cube; "coherent" = 4096 rays x 64 consecutive samples of the configs[1] marcher.  Tables:
    fp16 (the model's), bf16, fp32 (the fork's default dtype).  L2-warm (the table is resident in the 126 MB L2 in the training
    steady state) and with an L2 flush between launches.  Per point: HBM-compulsory bytes = 12 (xyz) + 32 s (output / incoming
    gradient), L2 rows = 128 corner rows."""
    import numpy as np
    from raw_ngp_b200 import _lib, raymarching, synthetic
    from raw_ngp_b200.gridencoder import GridEncoder
    B = 2 ** 18
    enc = GridEncoder(desired_resolution=2048).to(device)
    S = float(np.log2(enc.per_level_scale))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    pts = {"random": ((synthetic.uniform_points(B, seed=0) + 1) / 2).to(device)}
    if full:
        model, o, d, _ = build_scene(device, 0, n_rays=3 * 4096)     # enough rays through the ball for 4096 x 64 samples
        o, d = o.to(device), d.to(device)
        nears, fars = synthetic.near_far_torch(o, d, model.aabb_train, 0.05)
        xyzs, _, _, rays, _ = raymarching.march_rays_train(o, d, None, 1.0, False, model.density_bitfield, 1, 128, nears, fars, False, 0.0, 1024)
        rays = rays.long()
        sel = rays[:, 1] >= 64
        idx = (rays[sel, 0][:4096, None] + torch.arange(64, device=device)[None, :]).reshape(-1)
        assert idx.numel() == B, "not enough rays with 64 samples"
        pts["coherent"] = ((xyzs[idx] + 1) / 2).contiguous()
        del model
    dts = {"f16": (torch.float16, _lib.NGP_F16), "bf16": (torch.bfloat16, _lib.NGP_BF16), "f32": (torch.float32, _lib.NGP_F32)}
    gen = torch.Generator().manual_seed(1)
    grad32 = (torch.randn(B, 32, generator=gen) * 1e-3).to(device)
    res = {"points": B, "L": 16, "F": 2, "log2T": 19, "table_entries": int(enc.embeddings.shape[0]), "variants": {}}

    def flushed(fn):
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return _median(ts)

    def entry(ms, s, bwd=False, extra_hbm=0):
        hbm = (12 + 32 * s + extra_hbm) * B
        e = {"ms": ms, "mpts_s": B / ms / 1e3, "hbm_bytes_per_point": 12 + 32 * s + extra_hbm, "hbm_frac": hbm / (ms * 1e-3) / 1e9 / hbm_peak,
             "algorithmic_bytes_per_point": 12 + 256 * s + 32 * s + extra_hbm,
             "algorithmic_frac_of_hbm_peak": (12 + 256 * s + 32 * s + extra_hbm) * B / (ms * 1e-3) / 1e9 / hbm_peak}
        if ceil is not None:
            rows = 128 * B / (ms * 1e-3) / 1e9
            c = 2 * ceil["red_v2_f16x2_G_per_s"] if bwd else ceil["gather_rows_G_per_s"]
            e.update({"l2_G_rows_per_s": rows, "l2_frac": rows / c})
        return e

    for dname, (tdt, did) in dts.items():
        if not full and dname != "f16":
            continue
        table = enc.embeddings.data.to(tdt).contiguous()
        out = torch.empty(B, 32, device=device, dtype=tdt)
        sink = torch.zeros_like(table)
        grad = grad32.to(tdt).contiguous()
        gin = torch.zeros(B, 3, device=device)
        s = table.element_size()
        flags = _lib.NGP_GRID_REF_ROUNDING if tdt == torch.float16 else 0
        for pname, x in pts.items():
            def fwd(fl=flags):
                _lib.call("ngp_grid_encode_forward", x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3, 2, 16, 16,
                          S, 16, None, 0, 0, 0, did, fl, _lib.stream())

            def bwd(g_in=None):
                _lib.call("ngp_grid_encode_backward", grad.data_ptr(), x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), sink.data_ptr(),
                          B, 3, 2, 16, 16, S, 16, g_in, 0, 0, 0, did, 0, _lib.stream())
            v = {"fwd": entry(time_kernel(fwd), s), "bwd": entry(time_kernel(bwd), s, bwd=True)}
            if dname == "f16":
                v["fwd_l2_flushed_ms"] = flushed(fwd)
                v["bwd_l2_flushed_ms"] = flushed(bwd)
                v["fwd_point_level_kernel"] = entry(time_kernel(lambda: fwd(flags | _lib.NGP_GRID_POINT_LEVEL_KERNELS)), s)
            if full and (dname == "f16" or pname == "random"):
                v["bwd_with_input_grads"] = entry(time_kernel(lambda: bwd(gin.data_ptr())), s, bwd=True, extra_hbm=12)
            res["variants"][f"{dname}/{pname}"] = v
        del table, out, sink, grad
    try:
        from oracle import ref_cuda
        if ref_cuda.available():
            args = (enc.per_level_scale, enc.base_resolution)
            x = pts["random"]
            ref = {}
            for dname, (tdt, _) in dts.items():
                if not full and dname != "f16":
                    continue
                if tdt == torch.bfloat16:
                    continue          # the reference's bf16 instantiation has no packed atomics; not what the model uses
                table = enc.embeddings.data.to(tdt).contiguous()
                grad = grad32.to(tdt).contiguous()
                t_rf = time_kernel(lambda: ref_cuda.grid_forward(x, table, enc.offsets, *args))
                t_rb = time_kernel(lambda: ref_cuda.grid_backward(grad, x, table, enc.offsets, *args))
                ref[f"{dname}/random"] = {"fwd_ms": t_rf, "bwd_ms": t_rb, "fwd_mpts_s": B / t_rf / 1e3, "bwd_mpts_s": B / t_rb / 1e3}
            ref["note"] = "reference extension (unmodified source, sm_100a) incl. its wrapper's [L,B,C] permute and gradient zero-fill"
            res["reference_cuda"] = ref
    except Exception as e:  # the reference build is optional
        res["reference_cuda"] = {"unavailable": str(e)[:120]}
    # back-compatible summary keys (fp16 table, random points)
    h = res["variants"]["f16/random"]
    res.update({"dtype": "f16", "fwd_mpts_s": h["fwd"]["mpts_s"], "bwd_mpts_s": h["bwd"]["mpts_s"], "fwd_ms_l2_warm": h["fwd"]["ms"],
                "bwd_ms_l2_warm": h["bwd"]["ms"], "fwd_ms_l2_flushed": h["fwd_l2_flushed_ms"], "bwd_ms_l2_flushed": h["bwd_l2_flushed_ms"]})
    return res


def encoder_cpu(points=1 << 14):
    """SURVEY 8(d) "CPU baseline": the same grid interpolation fwd + bwd in vectorised PyTorch on the host cores
    (oracle/cpu_pipeline.py: index_select gathers, autograd index_add scatter), bounded sample, best of 3 after 1 warm-up."""
    from oracle import cpu_pipeline
    from raw_ngp_b200 import synthetic
    from raw_ngp_b200.gridencoder import GridEncoder
    torch.set_num_threads(os.cpu_count() or 1)
    enc = GridEncoder(desired_resolution=2048)
    table = enc.embeddings.data.clone().requires_grad_(True)
    x = (synthetic.uniform_points(points, seed=0) + 1) / 2
    g = torch.randn(points, 32, generator=torch.Generator().manual_seed(1))
    offs = [int(v) for v in enc.offsets]
    tf, tb = [], []
    for _ in range(4):
        t0 = time.perf_counter()
        y = cpu_pipeline.grid_encode(x, table, offs, enc.per_level_scale, 16)
        t1 = time.perf_counter()
        y.backward(g)
        t2 = time.perf_counter()
        table.grad = None
        tf.append(t1 - t0)
        tb.append(t2 - t1)
    return {"points": points, "fwd_mpts_s": points / min(tf[1:]) / 1e6, "bwd_mpts_s": points / min(tb[1:]) / 1e6, "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"{points} of the 2^18 random points, fp32 table, best of 3 after 1 warm-up (bwd includes the 48.8 MB zero-fill of the table gradient, as in the reference wrapper)"}


# ----------------------------------------------------------------------------------------------------------------------
# reference arms
# ----------------------------------------------------------------------------------------------------------------------
def reference_cuda_step(device, model, o, d, tgt, steps=20, warm=3):
    """The SAME training step through the reference's own, unmodified CUDA extensions (oracle/_ref, compiled for sm_100a) +
    the PyTorch pieces the reference uses (nn.Linear under autocast, GradScaler, torch.optim.Adam) -- oracle/ref_gpu_step.py.
    Context for the headline: how fast the reference itself is on this GPU.  None if oracle/_ref is not built."""
    try:
        from oracle import ref_cuda, ref_gpu_step
        if not ref_cuda.available():
            return {"unavailable": "oracle/_ref not built"}
        enc = model.grid_encoder
        torch.manual_seed(0)
        ref = ref_gpu_step.RefNeRF(enc.offsets.cpu(), enc.per_level_scale, enc.base_resolution, 1.0).to(device)
        rs = ref_gpu_step.RefTrainStep(ref, model.density_bitfield, model.aabb_train)
        for _ in range(warm):
            rs.step(o, d, tgt)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            rs.step(o, d, tgt)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"ms_per_step": ms, "rays_per_s": RAYS_PER_GPU / (ms * 1e-3), "samples_per_step": rs.num_points,
                "note": "reference extensions (unmodified source, sm_100a) + nn.Linear/autocast + GradScaler + torch Adam; fp32 table as in the fork"}
    except Exception as e:  # the reference build is optional
        return {"unavailable": str(e)[:160]}


def _cpu_step_setup(rays):
    from oracle import cpu_pipeline
    from raw_ngp_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    grid = synthetic.ball_density_grid(H=128, cascade=1)
    bitfield = synthetic.packbits_torch(grid, min(grid.clamp(min=0).mean().item(), 10.0))
    o, d = synthetic.sphere_rays(rays, seed=2)
    nears, fars = synthetic.near_far_torch(o, d, torch.tensor([-1.0] * 3 + [1.0] * 3), 0.05)
    nears, fars = nears.view(-1).contiguous(), fars.view(-1).contiguous()
    target = torch.rand(rays, 3, generator=torch.Generator().manual_seed(7))
    model = cpu_pipeline.CpuNeRF()
    optim = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    g = torch.Generator().manual_seed(3)

    def step():
        return cpu_pipeline.train_step(model, optim, o, d, target, bitfield, nears, fars, torch.rand(rays, generator=g))[1]
    return step


def cpu_baseline(steps=3, rays=CPU_SAMPLE_RAYS):
    """The CPU port on a bounded sample of the same workload (same scene, same ray distribution): best of `steps`."""
    step = _cpu_step_setup(rays)
    times, M = [], 0
    for _ in range(steps + 1):
        t0 = time.perf_counter()
        M = step()
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {"value": rays / best, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{rays} rays ({M} samples) per step of the configs[1] scene, full step incl. Adam over the 12.2M-entry table; best of {steps} after 1 warm-up",
            "ms_per_step": best * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    rays = CPU_SAMPLE_RAYS
    step = _cpu_step_setup(rays)
    M = 0
    # The whole run has to end within a few minutes on whatever host cores the box has (a step is ~0.2 s on 16 cores, most of it
    # Adam over the 12.2 M-entry table, which no smaller ray sample would shrink): if the first step says that W + K steps do not
    # fit NGP_REF_BUDGET_S (default 240 s), fewer steps are run and the line says so (`steps_requested`).
    budget = float(os.environ.get("NGP_REF_BUDGET_S", "240"))
    step()                                   # (lazy initialisation: not representative)
    t0 = time.perf_counter()
    step()
    t_first = time.perf_counter() - t0
    requested = K
    if (W + K) * t_first > budget:
        W = min(W, 2)
        K = max(3, min(K, int(budget / t_first) - W))
    for _ in range(max(W - 2, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(K):
        M = step()
    dt = time.perf_counter() - t0
    v = rays * K / dt
    line = {
        "impl": "reference", "metric": "training rays/s", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": rays, "samples_per_step": M},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{rays} rays ({M} samples) per step; the reference has no CPU implementation, this is oracle/cpu_pipeline.py (PyTorch on the host cores)"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if requested != K:
        line["steps_requested"] = requested
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# sub-blocks: the other BASELINE configurations
# ----------------------------------------------------------------------------------------------------------------------
def config4_block(device, rank, world, timer, K, R):
    """BASELINE configs[4] / SURVEY 8(d) config 5: data-parallel training with BARF pose refinement -- 8192 rays per GPU
    generated on the device from the refined poses of 100 cameras (different pixels per rank), annealed feature window, ray
    gradients -> d se3, peer-memory exchange of the table / MLP / se3 gradients (csrc/optim.cu), identical optimizers
    on every rank.  Run at every N (N = 1: the single-GPU base of the scaling curve)."""
    import torch.distributed as dist
    from raw_ngp_b200 import _lib, pose
    from raw_ngp_b200.trainer import FusedTrainStep
    n_rays = 8192
    model = build_model(device, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
    model.update_annealing(0.25)
    C, HW, focal = 100, 800, 1000.0
    poses = pose.look_at_poses(C, radius=2.0).to(device)
    g = torch.Generator().manual_seed(100 + rank)
    idx = torch.randint(0, C, (n_rays,), generator=g).to(device)
    ij = torch.randint(0, HW, (n_rays, 2), generator=g).float() + 0.5
    dirs = pose.pixel_directions(ij[:, 0], ij[:, 1], (focal, focal, HW / 2, HW / 2)).to(device)
    tgt = torch.rand(n_rays, 3, generator=g).to(device)
    cam = pose.CameraOptimizer(C, device)
    fs = FusedTrainStep(model, n_rays, loss_scale=128.0, pose_optimizer=cam, poses=poses, pose_lr=1e-3,
                        process_group=dist.group.WORLD if world > 1 else None)
    fs.set_camera_rays(idx, dirs, tgt)
    progress = [0.25]

    def one():
        progress[0] += 1e-5                      # the annealing window moves every step, as in training
        model.update_annealing(progress[0])
        return fs.step(update_grid=False)
    r = timer.regions(one, K, R, warmup=5)
    M = fs.last_num_points
    fs.flush()
    torch.cuda.synchronize()
    ms = r["median"] / K
    out = {"workload": "configs[4] data-parallel BARF step: 8192 rays/GPU from refined poses of 100 cameras, ray gradients -> se3, "
                       "table/MLP and se3 gradient exchange over NVLink peer memory (no NCCL call in the step)",
           "n_gpus": world, "rays_per_gpu": n_rays, "samples_rank0": M, "steps": K, "repeats": R, "ms_per_step": ms,
           "ms_per_step_min": r["min"] / K, "ms_per_step_max": r["max"] / K, "value": world * n_rays / (ms * 1e-3), "unit": "rays/s",
           "dp_mode": "single" if world == 1 else ("peer-memory fused reduce-scatter + Adam + all-gather" if fs.peer is not None else "NCCL all-reduce"),
           "se3_moved": float(fs.se3.abs().max().item())}
    del fs
    return out


def config2_block(device, timer, K, R):
    """BASELINE configs[2] / SURVEY 8(d) config 3: light-stage relighting step -- SH of view AND light direction (view_mlp 47 ->
    80 -> 80 -> 3), contraction (renderer.py:171-176: grid bound 2, 2 cascades), clamped_exp, HDR loss with exposure, 8192 rays."""
    from raw_ngp_b200 import synthetic
    from raw_ngp_b200.trainer import FusedTrainStep
    n_rays = 8192
    model = build_model(device, bound=2, contract=True, rfield=True, color_activation="clamped_exp", density_activation="clamped_exp")
    o, d = synthetic.sphere_rays(n_rays, seed=2)
    ld = synthetic.unit_vectors(n_rays, seed=3)
    tgt = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(7))
    exposure = torch.tensor([1.0, 0.25, 1.0 / 16])[torch.arange(n_rays) % 3]
    o, d, ld, tgt, exposure = (t.to(device) for t in (o, d, ld, tgt, exposure))
    fs = FusedTrainStep(model, n_rays, loss_scale=128.0, loss="hdr")
    fs.set_rays(o, d, tgt, rays_ldir=ld, exposure=exposure)
    r = timer.regions(lambda: fs.step(update_grid=False), K, R, warmup=3)
    M = fs.last_num_points
    fs.flush()
    torch.cuda.synchronize()
    ms = r["median"] / K
    kt = fs.profile_kernels(iters=3)
    out = {"workload": "configs[2] light-stage step: 8192 rays, SH(view) + SH(light), view_mlp 47-80-80-3, contraction (2 cascades), HDR loss",
           "rays": n_rays, "samples": M, "steps": K, "repeats": R, "ms_per_step": ms, "value": n_rays / (ms * 1e-3), "unit": "rays/s",
           "samples_per_s": M / (ms * 1e-3), "ns_per_sample": ms * 1e6 / max(M, 1), "warp_specialised_kernels": bool(fs.ws),
           "step_kernels_ms": {k: round(v, 5) for k, v in kt.items()}}
    del fs
    return out


def config3_block(device, rank, world, timer, frames=5):
    """BASELINE configs[3] / SURVEY 8(d) config 4: one 1920x1080 pinhole frame (fx = fy = 1200, camera on r = 2 looking at the
    origin, configs[1] scene, perturb off) through the inference loop of run_cuda (renderer.py:573-616); rays are split over the
    N ranks (interleaved tiles of 4096 rays so that every rank sees the same mix of hitting / missing rays), no collective;
    frame time = max over ranks."""
    from raw_ngp_b200 import parallel
    model = build_model(device)
    model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
    model.eval()
    W, H, f = 1920, 1080, 1200.0
    j, i = torch.meshgrid(torch.arange(H, device=device), torch.arange(W, device=device), indexing="ij")
    dirs = torch.stack([(i - W / 2) / f, -(j - H / 2) / f, -torch.ones_like(i, dtype=torch.float32)], -1).reshape(-1, 3)
    n_total = dirs.shape[0]
    ids = parallel.interleaved_tiles(n_total, rank, world, tile=4096).to(device)
    dirs = dirs[ids].contiguous()
    rays_o = torch.tensor([0.0, 0.0, 2.0], device=device).expand_as(dirs).contiguous()

    def frame():
        with torch.no_grad():
            return model.render(rays_o, dirs, bg_color=1.0, perturb=False)["image"]
    img = frame()
    r = timer.regions(frame, 1, frames, warmup=1)
    ms = r["median"]
    return {"workload": "configs[3] inference: 1920x1080 frame via march_rays / composite_rays, interleaved 4096-ray tiles over the ranks, no collective",
            "n_gpus": world, "rays_total": n_total, "rays_rank0": int(ids.numel()), "ms_per_frame": ms, "ms_per_frame_min": r["min"],
            "frames_per_s": 1e3 / ms, "value": n_total / (ms * 1e-3), "unit": "rays/s", "mean_colour_rank0": float(img.mean().item())}


# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from raw_ngp_b200 import _lib
    from raw_ngp_b200.trainer import FusedTrainStep
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (ours) needs a CUDA device: the hot path has no CPU fallback")
    _lib.load()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    K, W = args.steps, args.warmup
    R = args.repeats if args.repeats else max(3, min(11, 20000 // max(K, 1)))
    hbm_peak, peak_src = _peaks()
    timer = Timer(world, device)

    # the clock sampler starts first and has delivered a sample before anything is timed
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    model, o_cpu, d_cpu, tgt_cpu = build_scene(device, rank)
    step = FusedTrainStep(model, RAYS_PER_GPU, lr=1e-2, loss_scale=128.0, update_extra_interval=16,
                          process_group=dist.group.WORLD if world > 1 else None)
    o, d, tgt = o_cpu.to(device), d_cpu.to(device), tgt_cpu.to(device)
    # the occupancy grid of the synthetic scene is fixed (random-init weights would empty it): update_extra_state is
    # exercised once per 16 steps on a scratch copy so its cost is inside the timed region without changing M
    scratch = dict(grid=model.density_grid.clone(), bits=model.density_bitfield.clone())
    model.iter_density = 16

    def one_step(ro, rd, tg):
        if step.global_step % step.update_extra_interval == 0:
            step.flush()                      # the occupancy update queries the density with the up-to-date weights
            model.update_extra_state()
            model.density_grid.copy_(scratch["grid"])
            model.density_bitfield.copy_(scratch["bits"])
            model.iter_density = 16           # steady state of training: the partial update (renderer.py:853-876); the first 16 are full
        return step.step(ro, rd, tg, update_grid=False)

    # ---------------- device-resident timing: R regions of exactly K steps, median ----------------
    for _ in range(W):
        one_step(o, d, tgt)
    timer.barrier()
    c0 = sampler.mark() if rank == 0 else 0
    l0, r0 = _lib.launch_count, step.kernels_replayed
    dev_r = timer.regions(lambda: one_step(o, d, tgt), K, R)
    launches = ((_lib.launch_count - l0) + (step.kernels_replayed - r0)) // R   # per K-step region: eager C-ABI calls + kernels replayed from the CUDA graphs
    ms = dev_r["median"]
    M = step.last_num_points
    final_loss = float(step.loss.item())

    # ---------------- end to end: pinned host inputs, loss read back ----------------
    o_pin, d_pin, t_pin = o_cpu.pin_memory(), d_cpu.pin_memory(), tgt_cpu.pin_memory()
    # Every step's loss is read back into pinned host memory; the host consumes it one step late (double buffer + event), the
    # way a training loop logs it, so that the read-back of step k does not stall the launch of step k + 1.
    loss_host = [torch.empty(1, pin_memory=True), torch.empty(1, pin_memory=True)]
    loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_ms, host_losses = [], []
    for _ in range(R):
        timer.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for k in range(K):
            loss = one_step(o_pin, d_pin, t_pin)            # pinned host -> static device buffers (async H2D) inside step()
            loss_host[k & 1].copy_(loss.reshape(1), non_blocking=True)
            loss_ready[k & 1].record()
            if k > 0:
                loss_ready[(k - 1) & 1].synchronize()
                host_losses.append(float(loss_host[(k - 1) & 1][0]))
        loss_ready[(K - 1) & 1].synchronize()
        host_losses.append(float(loss_host[(K - 1) & 1][0]))
        f1.record()
        timer.barrier()
        e2e_ms.append(f0.elapsed_time(f1))
    assert len(host_losses) == K * R and all(v == v for v in host_losses)
    clocks = sampler.stop(c0) if rank == 0 else None      # sampled over the device-resident and end-to-end timed regions
    t = torch.tensor(e2e_ms, device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = _median(t.tolist())
    step.flush()        # every rank applies its pending update here (collective with N > 1): nothing rank-0-only may flush later

    extras = {}
    if not args.no_extras:
        K4 = max(10, min(K, 100))
        extras["configs4"] = config4_block(device, rank, world, timer, K4, 5)
        extras["configs3"] = config3_block(device, rank, world, timer)
        torch.cuda.empty_cache()

    if rank == 0:
        ceil = l2_ceilings(device)
        # ---------------- roofline of the dominant kernel, timed live inside the step (CUDA events per launch) --------
        kt = step.profile_kernels(iters=10)
        per_kernel = {k: kernel_roofline(k, kt[k], M, hbm_peak, ceil) for k in kt if k in STEP_KERNEL_BYTES}
        dom = max(per_kernel, key=lambda k: kt[k])
        pk = per_kernel[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": pk["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": pk["hbm_frac"],
                    "traffic": None, "peak_source": peak_src,
                    "definition": "achieved = compulsory HBM bytes per launch (inputs + saved activations + outputs; SURVEY 8d terms) / CUDA-event launch time; "
                                  "the table-row gathers / reductions are L2 traffic by design and are rated separately in `l2`",
                    "algorithmic_bytes_per_sample": pk["hbm_bytes_per_sample"], "algorithmic_bytes_per_launch": pk["hbm_bytes_per_sample"] * M,
                    "ms_per_launch": kt[dom], "samples_per_launch": M,
                    "l2": {k: pk.get(k) for k in ("l2_rows_per_sample", "l2_G_rows_per_s", "l2_ceiling_G_rows_per_s", "l2_frac", "l2_kind")},
                    "l2_ceilings": ceil,
                    "kernels": per_kernel,
                    "step_hbm_frac": sum(STEP_KERNEL_BYTES[k]["hbm"] for k in per_kernel) * M / (ms / K * 1e-3) / 1e9 / hbm_peak,
                    "step_kernels_ms": {k: round(v, 5) for k, v in kt.items()}}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            tj = json.load(open(prof))
            if dom in tj:
                roofline["traffic"] = tj[dom].get("dram_bytes_per_launch")
                roofline["traffic_note"] = tj[dom].get("note")
        micro = encoder_micro(device, hbm_peak, ceil, full=not args.no_extras)
        ref_step = reference_cuda_step(device, model, o, d, tgt)
        if world == 1 and not args.no_extras:
            del step
            torch.cuda.empty_cache()
            extras["configs2"] = config2_block(device, timer, max(5, min(K, 30)), 3)
        total_rays = RAYS_PER_GPU * world
        line = {
            "metric": "training rays/s", "value": total_rays * K / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_gpu": RAYS_PER_GPU, "samples_per_step_per_gpu": M, "samples_per_ray": M / RAYS_PER_GPU,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "timing": f"median of {R} timed regions of exactly {K} steps each (barrier + synchronize on both sides, CUDA events, max over ranks)",
                       "l2": "inputs re-read each step; table 23 MiB + grads are L2 resident by design (steady state of training); no flush",
                       "occupancy_update": "update_extra_state every 16 steps inside the timed region (the partial update of the training steady state: H^3/4 uniform + H^3/4 occupied cells)"},
            "timing": {"repeats": R, "ms_per_step_min": dev_r["min"] / K, "ms_per_step_max": dev_r["max"] / K, "region_ms": dev_r["regions"]},
            "samples_per_s": M * world * K / (ms * 1e-3),
            "e2e": {"value": total_rays * K / (ms_e2e * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": int(o_pin.numel() * 4 + d_pin.numel() * 4 + t_pin.numel() * 4), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "grid_encode": micro,
            "reference_cuda_step": ref_step, "final_loss": final_loss,
        }
        line.update(extras)
        if world == 1:
            line["cpu_baseline"] = cpu_baseline()
            if not args.no_extras:
                line["grid_encode"]["cpu"] = encoder_cpu()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=0, help="timed regions of K steps (default: 11, fewer for very large K)")
    ap.add_argument("--no-extras", action="store_true", dest="no_extras", help="headline line only (profiling runs)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
