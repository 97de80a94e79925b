#!/usr/bin/env python3
"""Developer tool: per-kernel breakdown of the training step with torch.profiler (not a benchmark)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raw_ngp_b200.trainer import FusedTrainStep, TrainStep  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    model, o, d, tgt = bench.build_scene(dev, 0)
    if "--autograd" in sys.argv:
        step = TrainStep(model, table_dtype=torch.float16)
    else:
        step = FusedTrainStep(model, bench.RAYS_PER_GPU)
    o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
    for _ in range(5):
        step.step(o, d, tgt, update_grid=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        step.step(o, d, tgt, update_grid=False)
    torch.cuda.synchronize()
    print(f"wall per step (no grid update): {(time.perf_counter() - t0) * 20:.3f} ms")
    t0 = time.perf_counter()
    model.update_extra_state()
    torch.cuda.synchronize()
    print(f"update_extra_state: {(time.perf_counter() - t0) * 1e3:.3f} ms")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            step.step(o, d, tgt, update_grid=False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
    if isinstance(step, FusedTrainStep):
        for k, v in step.profile_kernels(10).items():
            print(f"{k:40s} {v * 1e3:9.1f} us")


if __name__ == "__main__":
    main()
