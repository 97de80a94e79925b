#!/usr/bin/env python3
"""Developer tool: kernel-time breakdown of one partial occupancy-grid update (update_extra_state, iter >= 16) from torch.profiler."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
model = bench.build_model(dev)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
for _ in range(3):
    model.iter_density = 16
    model.update_extra_state()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    model.iter_density = 16
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.update_extra_state(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("device time per update (events), us:", sorted(ts)[5])
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.iter_density = 16
    model.update_extra_state(); torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tot = 0
for e in ev:
    d = e.time_range.end - e.time_range.start; tot += d
    print(f"{e.time_range.start - t0:8.1f} {d:7.1f}  {e.name.replace('ngp::(anonymous namespace)::','').replace('void ','')[:80]}")
print("sum of kernel times us:", round(tot, 1), " span us:", round(ev[-1].time_range.end - t0, 1))
