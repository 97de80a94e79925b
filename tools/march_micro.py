"""Developer tool: device time of one inference march call (ngp_march_rays) in the steady state of the frame loop:
rays already inside the occupied ball, n_step samples each."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from raw_ngp_b200 import _lib
dev = torch.device("cuda:0")
model, _, _, _ = bench.build_scene(dev, 0)
W, H, f = 1920, 1080, 1200.0
j, i = torch.meshgrid(torch.arange(H // 2 - 128, H // 2 + 128, device=dev), torch.arange(W // 2 - 512, W // 2 + 512, device=dev), indexing="ij")
dirs = torch.stack([(i - W / 2) / f, -(j - H / 2) / f, -torch.ones_like(i, dtype=torch.float32)], -1).reshape(-1, 3).contiguous()
rays_o = torch.tensor([0.0, 0.0, 2.0], device=dev).expand_as(dirs).contiguous()
N = dirs.shape[0]
nears = torch.full((N,), 1.0, device=dev); fars = torch.full((N,), 3.0, device=dev)
alive = torch.arange(N, dtype=torch.int32, device=dev)
for t0 in (1.8, 1.0):
    for n_step in (1, 2, 8):
        rays_t = torch.full((N,), t0, device=dev)
        ts_ = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            xyzs, d_, ts = torch.empty(N * n_step, 3, device=dev), torch.empty(N * n_step, 3, device=dev), torch.empty(N * n_step, 2, device=dev)
            noises = torch.zeros(N, device=dev)
            torch.cuda.synchronize()
            e0.record()
            _lib.call("ngp_march_rays", N, n_step, _lib.ptr(alive), _lib.ptr(rays_t), _lib.ptr(rays_o), _lib.ptr(dirs), 1.0, 0, 0.0, 1024, 1, 128,
                      _lib.ptr(model.density_bitfield), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(xyzs), _lib.ptr(d_), _lib.ptr(ts), _lib.ptr(noises), _lib.stream())
            e1.record(); torch.cuda.synchronize()
            ts_.append(e0.elapsed_time(e1) * 1e3)
        kept = int((ts[:, 0] > 0).sum().item())
        print(f"t0={t0} n_step={n_step}: {min(ts_):.1f} us (rays {N}, kept samples {kept})")
