#!/usr/bin/env python3
"""Developer tool: a few eager (non-graph) FusedTrainStep steps on the bench scene, for `ncu` captures.
Usage: python tools/ncu_step.py [steps] [step|pose|infer]
  step  (default) configs[1] training step
  pose  configs[4] step on one GPU: rays from refined poses, BARF window, input gradients (the IG kernel variants)
  infer one 1080p-shaped inference pass on 2^18 rays of the frame"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
from raw_ngp_b200.trainer import FusedTrainStep  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    mode = sys.argv[2] if len(sys.argv) > 2 else "step"
    dev = torch.device("cuda:0")
    if mode == "pose":
        import config_bench
        from raw_ngp_b200 import pose
        model = config_bench.build_model(dev, bound=1, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
        model.update_annealing(0.25)
        C, HW, focal, N = 100, 800, 1000.0, config_bench.N_RAYS
        g = torch.Generator().manual_seed(100)
        idx = torch.randint(0, C, (N,), generator=g).to(dev)
        ij = torch.randint(0, HW, (N, 2), generator=g).float() + 0.5
        dirs = pose.pixel_directions(ij[:, 0], ij[:, 1], (focal, focal, HW / 2, HW / 2)).to(dev)
        tgt = torch.rand(N, 3, generator=g).to(dev)
        fs = FusedTrainStep(model, N, pose_optimizer=pose.CameraOptimizer(C, dev), poses=pose.look_at_poses(C).to(dev), use_graph=False)
        for _ in range(steps):
            loss = fs.step(cam_idx=idx, dirs_cam=dirs, target_rgb=tgt, update_grid=False)
        fs.flush()
        torch.cuda.synchronize()
        print("loss", float(loss.item()), "samples", fs.last_num_points)
        return
    if mode == "infer":
        model, _, _, _ = bench.build_scene(dev, 0)
        model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
        model.eval()
        W, H, f = 1920, 1080, 1200.0
        j, i = torch.meshgrid(torch.arange(H // 2 - 128, H // 2 + 128, device=dev), torch.arange(W // 2 - 512, W // 2 + 512, device=dev), indexing="ij")
        dirs = torch.stack([(i - W / 2) / f, -(j - H / 2) / f, -torch.ones_like(i, dtype=torch.float32)], -1).reshape(-1, 3).contiguous()
        rays_o = torch.tensor([0.0, 0.0, 2.0], device=dev).expand_as(dirs).contiguous()
        with torch.no_grad():
            img = model.render(rays_o, dirs, bg_color=1.0, perturb=False)["image"]
        torch.cuda.synchronize()
        print("rays", dirs.shape[0], "mean colour", float(img.mean().item()))
        return
    model, o, d, tgt = bench.build_scene(dev, 0)
    step = FusedTrainStep(model, bench.RAYS_PER_GPU, use_graph=False)
    o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
    for _ in range(steps):
        loss = step.step(o, d, tgt, update_grid=False)
    torch.cuda.synchronize()
    print("loss", float(loss.item()), "samples", step.last_num_points)


if __name__ == "__main__":
    main()
