#!/usr/bin/env python3
"""Developer tool: per-step device time of the bench loop (which steps are slow: occupancy update, flush, ...)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from raw_ngp_b200.trainer import FusedTrainStep
dev = torch.device("cuda:0")
model, o, d, tgt = bench.build_scene(dev, 0)
step = FusedTrainStep(model, bench.RAYS_PER_GPU)
o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
scratch = dict(grid=model.density_grid.clone(), bits=model.density_bitfield.clone())
def one_step():
    if step.global_step % 16 == 0:
        step.flush(); model.update_extra_state()
        model.density_grid.copy_(scratch["grid"]); model.density_bitfield.copy_(scratch["bits"]); model.iter_density = 0
    return step.step(o, d, tgt, update_grid=False)
for _ in range(10): one_step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(65)]
ev[0].record()
for i in range(64):
    one_step(); ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(64)]
print(" ".join(f"{t:.2f}" for t in ts))
print("mean", sum(ts) / 64, "median", sorted(ts)[32])
