import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
from raw_ngp_b200 import _lib
dev = 'cuda'
g = torch.randn(12196240, device=dev).half()
w = torch.randn(14336, device=dev)
found = torch.zeros(1, device=dev); step = torch.zeros(1, dtype=torch.int32, device=dev); scratch = torch.zeros(2, dtype=torch.int32, device=dev)
args = ((ctypes.c_void_p * 2)(g.data_ptr(), w.data_ptr()), (ctypes.c_int * 2)(_lib.NGP_F16, _lib.NGP_F32), (ctypes.c_uint64 * 2)(g.numel(), w.numel()))
def run():
    _lib.call("ngp_check_finite_multi", args[0], args[1], args[2], 2, _lib.ptr(found), _lib.ptr(step), _lib.ptr(scratch), _lib.stream())
for _ in range(5): run()
torch.cuda.synchronize()
for label, flush in (("L2 warm", False), ("after 512 MB write", True)):
    big = torch.empty(128 * 1024 * 1024, device=dev) if flush else None
    ts = []
    for _ in range(20):
        if flush: big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    print(label, "single launch us (median):", sorted(ts)[10])
# device time per launch: 50 launches captured in a CUDA graph (eager back-to-back launches are bound by the host's ctypes calls)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(50): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("in a graph, us per launch:", e0.elapsed_time(e1) * 1e3 / 50, "found", found.item())
