"""Ceilings of the units that bound the two field kernels, measured on this GPU (csrc/diag.cu):
  gathers:    random 4-byte rows out of an L2-resident 32 MiB table, 8 independent loads per thread in flight
  reductions: random red.global.add.noftz.f16x2 / .v2.f16x2 into the same table
and, beside them, what the forward / backward kernels of the configs[1] step achieve (rows per second, from the sample
count of the step and the CUDA-event kernel times of FusedTrainStep.profile_kernels)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from raw_ngp_b200 import _lib
from raw_ngp_b200.trainer import FusedTrainStep

dev = torch.device("cuda:0")
n_rows = 1 << 23                        # 32 MiB of 4-byte rows: about the 23 MiB table + headroom, L2 resident (126 MB)
table = torch.zeros(n_rows, dtype=torch.int32, device=dev)
sink = torch.zeros(1, dtype=torch.int32, device=dev)
blocks, rounds = 148 * 4, 64

def rate(mode, threads_blocks=blocks):
    ops = threads_blocks * 512 * rounds * (8 if mode < 2 else 4)
    for _ in range(2):
        _lib.call("ngp_diag_l2_rate", _lib.ptr(table), n_rows, threads_blocks, rounds, mode, _lib.ptr(sink), _lib.stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.call("ngp_diag_l2_rate", _lib.ptr(table), n_rows, threads_blocks, rounds, mode, _lib.ptr(sink), _lib.stream())
    e1.record(); torch.cuda.synchronize()
    return ops * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9

g = rate(0); r1 = rate(1); r2 = rate(2)
print(f"ceiling: random 4-byte gathers      {g:7.1f} G rows/s")
print(f"ceiling: random red.add.f16x2       {r1:7.1f} G ops/s ({r1:7.1f} G rows/s)")
print(f"ceiling: random red.add.v2.f16x2    {r2:7.1f} G ops/s ({2 * r2:7.1f} G rows/s)")

model, o, d, tgt = bench.build_scene(dev, 0)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU, perturb=False)
fs.set_rays(o.to(dev), d.to(dev), tgt.to(dev))
for _ in range(3):
    fs.step(update_grid=False)
kt = fs.profile_kernels(10)
M = fs.last_num_points
rows = M * 16 * 8
tf, tb = kt["ngp_field_forward_full"] * 1e-3, kt["ngp_field_backward_full"] * 1e-3
print(f"forward : {M} samples x 128 corner rows in {tf * 1e6:.0f} us = {rows / tf / 1e9:6.1f} G rows/s gathered "
      f"= {rows / tf / 1e9 / g:.2f} of the gather ceiling")
print(f"backward: {M} samples x 128 corner rows in {tb * 1e6:.0f} us = {rows / tb / 1e9:6.1f} G rows/s reduced (before warp "
      f"aggregation / pairing) = {rows / tb / 1e9 / r1:.2f} of the scalar-reduction ceiling, {rows / tb / 1e9 / (2 * r2):.2f} of the paired one")
