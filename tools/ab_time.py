#!/usr/bin/env python3
"""Developer tool: forward / backward kernel times and the graph step time of BASELINE configs[1] for the loaded library
(NGP_B200_LIB selects a variant built by tools/ab_variant.sh).  Prints one JSON line."""
import json, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200.trainer import FusedTrainStep

dev = torch.device("cuda", 0)
model, o, d, tgt = bench.build_scene(dev, 0)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU)
fs.set_rays(o.to(dev), d.to(dev), tgt.to(dev))
for _ in range(20):
    fs.step(update_grid=False)
torch.cuda.synchronize()
k = fs.profile_kernels(iters=20)
ts = []
for rep in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        fs.step(update_grid=False)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50)
print(json.dumps({"lib": os.path.basename(os.environ.get("NGP_B200_LIB", "default")), "fwd_us": round(1e3 * k["ngp_field_forward_full"], 1),
                  "bwd_us": round(1e3 * k["ngp_field_backward_full"], 1), "adam_us": round(1e3 * k["ngp_fused_adam"], 1), "check_us": round(1e3 * k["ngp_check_finite_multi"], 1),
                  "march_us": round(1e3 * (k["ngp_march_rays_train_count_ex"] + k["ngp_march_rays_train_write"]), 1), "step_us": round(1e3 * statistics.median(ts), 1),
                  "M": fs.last_num_points}))
