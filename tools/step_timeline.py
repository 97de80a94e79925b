#!/usr/bin/env python3
"""Developer tool: device timeline of the pipelined training step (BASELINE configs[1], one GPU) from torch.profiler's CUDA
activity records (CUPTI): start offset, duration and stream of every kernel of one replayed step, averaged over 20 steps.
    python tools/step_timeline.py [out.json]"""
import json, os, sys, statistics, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200.trainer import FusedTrainStep
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
model, o, d, tgt = bench.build_scene(dev, 0)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU)
fs.set_rays(o.to(dev), d.to(dev), tgt.to(dev))
for _ in range(30):
    fs.step(update_grid=False)
torch.cuda.synchronize()
N = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        fs.step(update_grid=False)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
# split into steps at the first kernel of the optimizer chain (check_finite) -- one per step
starts = [i for i, e in enumerate(ev) if "check_finite" in e.name]
rows = collections.OrderedDict()
for a, b in zip(starts[2:-1], starts[3:]):
    t0 = min(e.time_range.start for e in ev[a:b])
    for e in ev[a:b]:
        key = e.name.replace("void ", "").replace("ngp::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").split("(")[0][:56]
        rows.setdefault(key, []).append((e.time_range.start - t0, e.time_range.end - e.time_range.start))
    rows.setdefault("__step__", []).append((0.0, max(e.time_range.end for e in ev[a:b]) - t0))
out = []
for k, v in rows.items():
    out.append({"kernel": k, "n": len(v), "start_us": round(statistics.mean(x[0] for x in v), 1), "dur_us": round(statistics.mean(x[1] for x in v), 1)})
out.sort(key=lambda r: r["start_us"])
for r in out:
    print(f'{r["kernel"]:50s} start {r["start_us"]:8.1f}  dur {r["dur_us"]:7.1f}  (n={r["n"]})')
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
