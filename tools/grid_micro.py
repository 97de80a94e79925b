#!/usr/bin/env python3
"""Developer tool: one variant of the BASELINE configs[0] micro-benchmark (GridEncoder fwd / bwd on 2^18 points), for ncu:

    python tools/grid_micro.py [fwd|fwd_point_level|bwd|bwd_ig] [f16|bf16|f32] [random|coherent] [iters]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raw_ngp_b200 import _lib, raymarching, synthetic  # noqa: E402
from raw_ngp_b200.gridencoder import GridEncoder  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dname = sys.argv[2] if len(sys.argv) > 2 else "f16"
pname = sys.argv[3] if len(sys.argv) > 3 else "random"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
B = 2 ** 18
enc = GridEncoder(desired_resolution=2048).to(dev)
S = float(np.log2(enc.per_level_scale))
tdt, did = {"f16": (torch.float16, _lib.NGP_F16), "bf16": (torch.bfloat16, _lib.NGP_BF16), "f32": (torch.float32, _lib.NGP_F32)}[dname]
if pname == "random":
    x = ((synthetic.uniform_points(B, seed=0) + 1) / 2).to(dev)
else:
    model, o, d, _ = bench.build_scene(dev, 0, n_rays=3 * 4096)
    o, d = o.to(dev), d.to(dev)
    nears, fars = synthetic.near_far_torch(o, d, model.aabb_train, 0.05)
    xyzs, _, _, rays, _ = raymarching.march_rays_train(o, d, None, 1.0, False, model.density_bitfield, 1, 128, nears, fars, False, 0.0, 1024)
    rays = rays.long()
    sel = rays[:, 1] >= 64
    idx = (rays[sel, 0][:4096, None] + torch.arange(64, device=dev)[None, :]).reshape(-1)
    x = ((xyzs[idx] + 1) / 2).contiguous()
table = enc.embeddings.data.to(tdt).contiguous()
out = torch.empty(B, 32, device=dev, dtype=tdt)
sink = torch.zeros_like(table)
grad = (torch.randn(B, 32, generator=torch.Generator().manual_seed(1)) * 1e-3).to(dev).to(tdt)
gin = torch.zeros(B, 3, device=dev)
flags = _lib.NGP_GRID_REF_ROUNDING if tdt == torch.float16 else 0
if what == "fwd_point_level":
    flags |= _lib.NGP_GRID_POINT_LEVEL_KERNELS


def run():
    if what.startswith("fwd"):
        _lib.call("ngp_grid_encode_forward", x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), out.data_ptr(), B, 3, 2, 16, 16, S, 16,
                  None, 0, 0, 0, did, flags, _lib.stream())
    else:
        _lib.call("ngp_grid_encode_backward", grad.data_ptr(), x.data_ptr(), table.data_ptr(), enc.offsets.data_ptr(), sink.data_ptr(), B, 3, 2,
                  16, 16, S, 16, gin.data_ptr() if what == "bwd_ig" else None, 0, 0, 0, did, 0, _lib.stream())


ms = bench.time_kernel(run, iters=iters, warm=2, repeats=3)
print(f"{what} {dname} {pname}: {ms * 1e3:.1f} us, {B / ms / 1e3:.0f} Mpts/s")
