import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from raw_ngp_b200 import raymarching
from raw_ngp_b200.nerf import near_far_from_aabb
dev = torch.device("cuda:0")
model, o, d, tgt = bench.build_scene(dev, 0)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
o, d = o.to(dev), d.to(dev)
nears, fars = near_far_from_aabb(o, d, model.aabb_train, 0.05)
xyzs, dirs, ts, rays, _ = raymarching.march_rays_train(o, d, None, 1.0, False, model.density_bitfield, 1, 128, nears, fars, True, 0, 1024)
print("M", xyzs.shape[0])
def run(grad):
    with torch.amp.autocast("cuda"):
        if grad:
            out = model(xyzs, dirs)
        else:
            with torch.no_grad():
                out = model(xyzs, dirs)
    return out
for grad in (False, True):
    if grad:
        for p in model.parameters(): p.requires_grad_(True)
    for _ in range(3): run(grad)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): run(grad)
    e1.record(); torch.cuda.synchronize()
    print("save activations" if grad else "no save", e0.elapsed_time(e1) / 10 * 1e3, "us per forward (incl. python wrapper)")
