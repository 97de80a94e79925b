#!/usr/bin/env python3
"""Developer experiment: time of the warp-specialised forward at configs[1] size with (a) all saves, (b) enc + in2 only, (c) none."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from raw_ngp_b200 import _lib
from raw_ngp_b200.trainer import FusedTrainStep
dev = torch.device("cuda:0")
model, o, d, tgt = bench.build_scene(dev, 0)
fs = FusedTrainStep(model, bench.RAYS_PER_GPU, use_graph=False, perturb=False)
fs.set_rays(o.to(dev), d.to(dev), tgt.to(dev))
fs._launch_forward_backward(); torch.cuda.synchronize()
m, opt, cap, ct = fs.model, fs.model.opt, fs.cap, ctypes
P = _lib.ptr
S, H, L, gt, ac, ip = fs._grid_scalars
enc = m.grid_encoder
c1, c2 = (ct.c_uint32 * 4)(*fs.p1), (ct.c_uint32 * 4)(*fs.p2)
w1, w2 = fs._ptrs(fs._w_lp_views[:3]), fs._ptrs(fs._w_lp_views[3:])
a1, a2 = fs._ptrs(fs.acts1), fs._ptrs(fs.acts2)
def run(enc_buf, acts1, in2, acts2):
    _lib.call("ngp_field_forward_full", P(fs.xyzs), P(fs.dirs), None, P(enc.embeddings), P(enc.offsets), None, float(m.bound), S, H, L, gt, ac, ip,
              w1, c1, w2, c2, cap, fs._m_dev, fs._density_act, float(opt.beta), fs._color_act, enc_buf, acts1, in2, acts2, P(fs.sigma), P(fs.rgb), None,
              _lib.stream())
for name, args in (("all saves", (P(fs.enc_buf), a1, P(fs.in2), a2)), ("enc + in2 only", (P(fs.enc_buf), None, P(fs.in2), None)), ("no saves", (None, None, None, None))):
    ms = bench.time_kernel(lambda: run(*args), iters=20, warm=3, repeats=5)
    print(f"{name:16s} {ms * 1e3:7.1f} us  (M = {fs.last_num_points})")
