#!/usr/bin/env python3
"""Developer tool: where a data-parallel step spends its time (CUDA events on both streams), under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 tools/dp_timeline.py
Prints, per rank, the mean over 60 steps of: march graph, update chain (side stream), exposed wait (main stream idle between the
end of the march and the start of the field graph), field graph, whole step."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200 import _lib, parallel
from raw_ngp_b200.trainer import FusedTrainStep

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model, o, d, tgt = bench.build_scene(dev, rank)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU, process_group=dist.group.WORLD if world > 1 else None)
o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
fs.set_rays(o, d, tgt)
for _ in range(10):
    fs.step(update_grid=False)
torch.cuda.synchronize()
if world == 1:
    print("single GPU: one pipelined graph per step; nothing to split"); sys.exit(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
main = torch.cuda.current_stream()
for it in range(60):
    t0, m1, u0, u1, f0, f1 = ev(), ev(), ev(), ev(), ev(), ev()
    t0.record(main)
    fs._side.wait_stream(main)
    with torch.cuda.stream(fs._side):
        u0.record(fs._side)
        if fs._graph_update is not None:
            _lib.weights_epoch += 1; fs.opt.step_count += 1
            fs._graph_update.replay()
        else:
            fs._reduce_and_update()
        u1.record(fs._side)
    fs._graph_march.replay()
    m1.record(main)
    main.wait_stream(fs._side)
    fs._launch_scaler_update()
    f0.record(main)
    fs._graph_field.replay()
    f1.record(main)
    fs._pending = True
    rows.append((t0, m1, u0, u1, f0, f1))
torch.cuda.synchronize()
import statistics as st
def mean(f): return st.mean(f(*r) for r in rows[5:])
out = {"rank": rank, "march_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: t0.elapsed_time(m1)),
       "update_chain_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: u0.elapsed_time(u1)),
       "update_start_after_step_start_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: t0.elapsed_time(u0)),
       "exposed_wait_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: m1.elapsed_time(f0)),
       "field_graph_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: f0.elapsed_time(f1)),
       "step_us": 1e3 * mean(lambda t0, m1, u0, u1, f0, f1: t0.elapsed_time(f1)), "graph_update": fs._graph_update is not None}
print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
fs.flush()
dist.barrier(); dist.destroy_process_group()
