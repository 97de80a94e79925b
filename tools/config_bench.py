#!/usr/bin/env python3
"""Training-step throughput on BASELINE.json configs[2] (light stage) and configs[4] (data-parallel BARF), SURVEY.md 8(d).
bench.py measures configs[1]; this tool gives the other two training configurations the same treatment (synthetic scene,
CUDA events, max over ranks) and, with --autograd, times the op-by-op autograd path (TrainStep, the reference's structure on
this repo's operators) beside the captured step.

    python tools/config_bench.py --config 3 [--bound 2|8] [--steps 200] [--autograd] [--profile]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/config_bench.py --config 5

Prints one JSON line per measurement (rank 0).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_RAYS = 8192


def build_model(device, **kw):
    from raw_ngp_b200 import raymarching, synthetic
    from raw_ngp_b200.nerf import NeRFNetwork, default_opt
    torch.manual_seed(0)
    cfg = dict(grid_size=128, max_steps=1024, dt_gamma=0, T_thresh=1e-8, min_near=0.05, fp16=True, density_thresh=10,
               hashmap_size=19, hashgrid_resolution=2048)
    cfg.update(kw)
    model = NeRFNetwork(default_opt(**cfg)).to(device)
    grid = synthetic.ball_density_grid(H=128, cascade=model.cascade, bound=float(model.bound), radius=0.5, sigma=50.0).to(device)
    model.density_grid.copy_(grid)
    thresh = min(grid.clamp(min=0).mean().item(), 10.0)
    model.density_bitfield = raymarching.packbits(model.density_grid, thresh, model.density_bitfield)
    model.mean_density = grid.clamp(min=0).mean().item()
    return model


def timed(fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def config3(args, device, barrier, report):
    """light-stage relighting: SH of view AND light direction (view_mlp 47 -> 80 -> 80 -> 3), HDR loss with per-ray exposure,
    scene contraction (through the renderer contraction means a [-2, 2] grid with 2 cascades, renderer.py:171-176, while the
    marcher receives the real bound: --bound 2 or --bound 8)."""
    from raw_ngp_b200 import synthetic
    from raw_ngp_b200.trainer import FusedTrainStep, TrainStep
    kw = dict(bound=args.bound, contract=True, rfield=True, color_activation="clamped_exp", density_activation="clamped_exp")
    model = build_model(device, **kw)
    o, d = synthetic.sphere_rays(N_RAYS, seed=2)
    ld = synthetic.unit_vectors(N_RAYS, seed=3)
    tgt = torch.rand(N_RAYS, 3, generator=torch.Generator().manual_seed(7))
    exposure = torch.tensor([1.0, 0.25, 1.0 / 16])[torch.arange(N_RAYS) % 3]
    o, d, ld, tgt, exposure = (t.to(device) for t in (o, d, ld, tgt, exposure))
    base = dict(config="configs[2] light-stage step", rays=N_RAYS, cascades=int(model.cascade), bound=float(model.real_bound),
                contract=True, view_mlp="47-80-80-3", loss="hdr")
    if args.autograd:
        import copy
        ref_model = copy.deepcopy(model)
        ts = TrainStep(ref_model, loss_scale=128.0)

        def auto_step():
            ref_model.train()
            out = ref_model.render(o, d, rays_ldir=ld, bg_color=1.0, perturb=True)
            clip = torch.minimum(torch.ones((), device=device), out["image"] * exposure.unsqueeze(1))
            loss = ((clip - tgt) ** 2 * (1.0 / (1e-3 + clip.detach())) ** 2).sum() / (3 * N_RAYS)
            (loss * 128.0).backward()
            ts.found_inf.zero_()
            ts.inv_scale.fill_(1 / 128.0)
            ts.opt.step(ts.inv_scale, ts.found_inf, zero_grad=True)
            return out["num_points"]
        ms = timed(auto_step, max(args.steps // 10, 5), 3, barrier)
        report(dict(base, path="autograd op-by-op (TrainStep)", ms_per_step=round(ms, 4), rays_per_s=round(N_RAYS / ms * 1e3, 1)))
    fs = FusedTrainStep(model, N_RAYS, loss_scale=128.0, loss="hdr", max_samples=args.max_samples)
    fs.set_rays(o, d, tgt, rays_ldir=ld, exposure=exposure)
    ms = timed(lambda: fs.step(update_grid=False), args.steps, args.warmup, barrier)
    M = fs.last_num_points
    report(dict(base, path="FusedTrainStep (CUDA graph)", samples=M, ms_per_step=round(ms, 4), rays_per_s=round(N_RAYS / ms * 1e3, 1),
                samples_per_s=round(M / ms * 1e3, 1)))
    if args.profile:
        prof = fs.profile_kernels(10)
        report(dict(base, kernels_us={k: round(v * 1e3, 1) for k, v in prof.items()}))


def config5(args, device, barrier, report, world, rank):
    """data-parallel training with BARF pose refinement: 8192 rays per GPU from refined poses of 100 cameras, annealed feature
    window, all-reduce of the table / MLP / se3 gradients, identical optimizers on every rank."""
    import torch.distributed as dist
    from raw_ngp_b200 import pose
    from raw_ngp_b200.trainer import FusedTrainStep
    model = build_model(device, bound=1, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
    model.update_annealing(0.25)
    C, HW, focal = 100, 800, 1000.0
    poses = pose.look_at_poses(C, radius=2.0).to(device)
    g = torch.Generator().manual_seed(100 + rank)
    idx = torch.randint(0, C, (N_RAYS,), generator=g).to(device)
    ij = torch.randint(0, HW, (N_RAYS, 2), generator=g).float() + 0.5
    dirs = pose.pixel_directions(ij[:, 0], ij[:, 1], (focal, focal, HW / 2, HW / 2)).to(device)
    tgt = torch.rand(N_RAYS, 3, generator=g).to(device)
    cam = pose.CameraOptimizer(C, device)
    pg = dist.group.WORLD if world > 1 else None
    fs = FusedTrainStep(model, N_RAYS, loss_scale=128.0, pose_optimizer=cam, poses=poses, pose_lr=1e-3, process_group=pg,
                        max_samples=args.max_samples)
    fs.set_camera_rays(idx, dirs, tgt)
    progress = [0.25]

    def one():
        progress[0] += 1e-5                      # the annealing window moves every step, as in training
        model.update_annealing(progress[0])
        return fs.step(update_grid=False)
    ms = timed(one, args.steps, args.warmup, barrier)
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    M = fs.last_num_points
    fs.flush()
    torch.cuda.synchronize()
    report(dict(config="configs[4] data-parallel BARF step", n_gpus=world, rays_per_gpu=N_RAYS, cameras=C, samples_rank0=M,
                path="FusedTrainStep pose mode (CUDA graphs + NCCL)", ms_per_step=round(ms, 4),
                rays_per_s=round(world * N_RAYS / ms * 1e3, 1), se3_moved=float(fs.se3.abs().max().item())))
    if args.profile and world == 1:
        prof = fs.profile_kernels(10)
        report(dict(config="configs[4]", kernels_us={k: round(v * 1e3, 1) for k, v in prof.items()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[3, 5])
    ap.add_argument("--bound", type=int, default=2, choices=[2, 8])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--autograd", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--max-samples", type=int, default=None, dest="max_samples")
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def report(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    if args.config == 3:
        config3(args, device, barrier, report)
    else:
        config5(args, device, barrier, report, world, rank)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
