#!/usr/bin/env python3
"""Developer tool: time of the field forward / backward kernels alone (BASELINE configs[1] samples), for kernel variants whose
outputs are not meaningful (role-isolation builds of tools/ab_variant.sh).  NGP_B200_LIB selects the library."""
import json, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200 import _lib
from raw_ngp_b200.trainer import FusedTrainStep

dev = torch.device("cuda", 0)
model, o, d, tgt = bench.build_scene(dev, 0)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU, use_graph=False)
fs.set_rays(o.to(dev), d.to(dev), tgt.to(dev))
fs._launch_march()
torch.cuda.synchronize()
names, ev = [], []
real = _lib.call
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); real(name, *a); e1.record()
    names.append(name); ev.append((e0, e1))
_lib.call = timed
for _ in range(30):
    fs._launch_field()
torch.cuda.synchronize()
_lib.call = real
out = {"lib": os.path.basename(os.environ.get("NGP_B200_LIB", "default")), "M": int(fs.counter[0].item())}
for k in ("ngp_field_forward_full", "ngp_field_backward_full"):
    ts = [e0.elapsed_time(e1) for n, (e0, e1) in zip(names, ev) if n == k][5:]
    out[k.replace("ngp_field_", "").replace("_full", "") + "_us"] = round(1e3 * statistics.median(ts), 1)
print(json.dumps(out))
