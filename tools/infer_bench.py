#!/usr/bin/env python3
"""Developer tool: BASELINE configs[3] -- render one 1920x1080 frame through march_rays / composite_rays
(NeRFRenderer.run_cuda inference loop, nerf/renderer.py:573-616) on the configs[1] synthetic scene.  Rays are split into
contiguous tiles; with torchrun each rank renders its share (no collective).  Prints frames/s and rays/s."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raw_ngp_b200 import parallel  # noqa: E402


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    model, _, _, _ = bench.build_scene(dev, 0)
    model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
    model.eval()
    W, H, f = 1920, 1080, 1200.0
    j, i = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    dirs = torch.stack([(i - W / 2) / f, -(j - H / 2) / f, -torch.ones_like(i, dtype=torch.float32)], -1).reshape(-1, 3)
    rays_o = torch.tensor([0.0, 0.0, 2.0], device=dev).expand_as(dirs).contiguous()        # camera on r = 2 looking at the origin
    lo, hi = parallel.shard_range(dirs.shape[0], rank, world)
    chunk = 1 << int(os.environ.get('INFER_CHUNK_LOG2', '21'))
    def frame():
        outs = []
        with torch.no_grad():
            for s in range(lo, hi, chunk):
                e = min(s + chunk, hi)
                outs.append(model.render(rays_o[s:e], dirs[s:e].contiguous(), bg_color=1.0, perturb=False)["image"])
        return torch.cat(outs)
    img = frame()
    img = frame()
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        t0 = time.perf_counter()
        img = frame()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt, best = sum(times) / len(times), min(times)
    print(f"rank {rank}/{world}: {hi - lo} rays, {dt * 1e3:.1f} ms per frame share (best {best * 1e3:.1f}), "
          f"{(hi - lo) / dt / 1e6:.2f} Mrays/s, {1 / dt:.2f} frames/s (share), mean colour {img.mean().item():.4f}", flush=True)

if __name__ == "__main__":
    main()
