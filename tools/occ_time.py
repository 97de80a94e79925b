import torch, time, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
model = bench.build_model(dev)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
for mode, it0 in (("full", 0), ("partial", 16)):
    model.iter_density = it0
    for _ in range(3):
        model.iter_density = it0
        model.update_extra_state()
    torch.cuda.synchronize()
    ts, tw = [], []
    for _ in range(10):
        model.iter_density = it0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); e0.record()
        model.update_extra_state()
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); tw.append((t1 - t0) * 1e3)
    print(mode, "gpu ms", sorted(ts)[5], "host enqueue ms", sorted(tw)[5])
    if hasattr(model, "update_extra_state_graph"):
        pass
