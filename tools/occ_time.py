import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device("cuda:0")
model, o, d, tgt = bench.build_scene(dev, 0)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
grid0 = model.density_grid.clone(); bits0 = model.density_bitfield.clone()
for mode, it0 in (("full", 0), ("partial", 16)):
    for rep in range(3):
        model.density_grid.copy_(grid0); model.density_bitfield.copy_(bits0); model.iter_density = it0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model.update_extra_state()
        torch.cuda.synchronize(); t1 = time.perf_counter()
    print(mode, f"{(t1 - t0) * 1e3:.3f} ms wall")
from torch.profiler import profile, ProfilerActivity
model.iter_density = 0
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    model.update_extra_state(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
