#!/usr/bin/env python3
"""Data-parallel check on N GPUs of one node (torchrun): the peer-memory update (reduce-scatter + Adam + all-gather in one
kernel, csrc/optim.cu) against the NCCL all-reduce + replicated Adam path, from the same initial model and the same rays:
  * after K steps the fp16 table and the MLP weights are BIT-IDENTICAL on all ranks (each mode),
  * the two modes agree up to the summation order of the fp16 gradients (NCCL reduces in fp16, the kernel in fp32),
  * device-timed ms/step of both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_check.py
"""
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from raw_ngp_b200.trainer import FusedTrainStep
    model0, o, d, tgt = bench.build_scene(dev, rank)
    o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
    K = int(os.environ.get("DP_STEPS", "30"))
    out = {}
    tables = {}
    for mode in ("peer", "nccl"):
        os.environ["NGP_DP_PEER"] = "1" if mode == "peer" else "0"
        model = copy.deepcopy(model0)
        fs = FusedTrainStep(model, o.shape[0], lr=1e-2, loss_scale=128.0, perturb=False)
        assert (fs.peer is not None) == (mode == "peer")
        losses = [fs.step(o, d, tgt, update_grid=False).item() for _ in range(K)]
        fs.flush()
        torch.cuda.synchronize()
        table = model.grid_encoder.embeddings.data.clone()
        w = fs.w_lp.clone()
        master = fs.gather_table_master().clone()
        # identical on all ranks?
        ref_t, ref_w = table.clone(), w.clone()
        dist.broadcast(ref_t, 0); dist.broadcast(ref_w, 0)
        same = torch.tensor([float(torch.equal(ref_t, table) and torch.equal(ref_w, w))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        # timing
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fs.step(update_grid=False)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        fs.flush(); torch.cuda.synchronize()
        tables[mode] = (table.float(), w.float())
        out[mode] = dict(loss_first=losses[0], loss_last=losses[-1], ranks_bit_identical=bool(same.item()), ms_per_step=round(float(ms.item()), 4),
                         rays_per_s=round(world * o.shape[0] / float(ms.item()) * 1e3, 1))
        out[mode]["master_vs_table_max"] = float((master.half().float() - table.float()).abs().max().item())
        del fs
    dt = (tables["peer"][0] - tables["nccl"][0]).abs()
    moved = (tables["nccl"][0] - model0.grid_encoder.embeddings.data.float()).abs()
    out["peer_vs_nccl"] = dict(table_max_abs_diff=float(dt.max().item()), table_mean_abs_diff=float(dt.mean().item()),
                               table_mean_abs_update=float(moved.mean().item()),
                               w_max_abs_diff=float((tables["peer"][1] - tables["nccl"][1]).abs().max().item()))
    # ---- BARF pose refinement (BASELINE configs[4]): the se3 gradient is summed through the peer mappings inside the pose Adam
    # kernel (peer mode) or all-reduced by NCCL; se3 must be bit-identical across ranks and agree between the modes
    from raw_ngp_b200 import pose
    C, n_rays = 100, 4096
    poses = pose.look_at_poses(C, radius=2.0).to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    idx = torch.randint(0, C, (n_rays,), generator=g).to(dev)
    ij = torch.randint(0, 800, (n_rays, 2), generator=g).float() + 0.5
    dirs = pose.pixel_directions(ij[:, 0], ij[:, 1], (1000.0, 1000.0, 400.0, 400.0)).to(dev)
    tg = torch.rand(n_rays, 3, generator=g).to(dev)
    se3s = {}
    for mode in ("peer", "nccl"):
        os.environ["NGP_DP_PEER"] = "1" if mode == "peer" else "0"
        model = bench.build_model(dev, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
        model.update_annealing(0.25)
        cam = pose.CameraOptimizer(C, dev)
        fs = FusedTrainStep(model, n_rays, loss_scale=128.0, pose_optimizer=cam, poses=poses, pose_lr=1e-3, perturb=False,
                            process_group=dist.group.WORLD)
        fs.set_camera_rays(idx, dirs, tg)
        for _ in range(K):
            fs.step(update_grid=False)
        fs.flush()
        torch.cuda.synchronize()
        se3 = fs.se3.clone()
        ref = se3.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([float(torch.equal(ref, se3))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fs.step(update_grid=False)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 200], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        fs.flush(); torch.cuda.synchronize()
        se3s[mode] = se3
        out["pose_" + mode] = dict(se3_bit_identical_across_ranks=bool(same.item()), se3_max_abs=float(se3.abs().max().item()),
                                   pose_steps=int(fs.pose_step_dev.item()), graph=getattr(fs, "_graph_um", None) is not None,
                                   ms_per_step=round(float(ms.item()), 4))
        del fs
    out["pose_peer_vs_nccl"] = dict(se3_max_abs_diff=float((se3s["peer"] - se3s["nccl"]).abs().max().item()))
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, steps=K, **out)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
