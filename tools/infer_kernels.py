#!/usr/bin/env python3
"""Developer tool: kernel-time breakdown of one 1920x1080 inference frame (BASELINE configs[3]) from torch.profiler, for the whole
frame on one GPU and for a 1/8 share of its ray tiles (what one rank of eight renders)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200 import parallel
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
model.eval()
W, H, f = 1920, 1080, 1200.0
j, i = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
dirs_all = torch.stack([(i - W / 2) / f, -(j - H / 2) / f, -torch.ones_like(i, dtype=torch.float32)], -1).reshape(-1, 3)
for world in (1, 8):
    ids = parallel.interleaved_tiles(dirs_all.shape[0], 0, world, tile=4096).to(dev)
    dirs = dirs_all[ids].contiguous()
    rays_o = torch.tensor([0.0, 0.0, 2.0], device=dev).expand_as(dirs).contiguous()
    def frame():
        with torch.no_grad():
            return model.render(rays_o, dirs, bg_color=1.0, perturb=False)["image"]
    for _ in range(3): frame()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); img = frame(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        frame(); torch.cuda.synchronize()
    rows = sorted(((e.key.replace("ngp::(anonymous namespace)::", "").replace("void ", "").split("(")[0][:40], e.count, round(e.device_time_total / 1e3, 3))
                   for e in prof.key_averages() if e.device_time_total > 0), key=lambda r: -r[2])[:6]
    print(f"1/{world} of the frame: {sorted(ts)[2]:.3f} ms, mean colour {img.mean().item():.6f};", rows)
