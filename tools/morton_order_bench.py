"""Evidence for DESIGN.md's statement on Morton reordering (north_star: "input points are reordered by Morton code so that
neighbouring corners coalesce"): grid_encode forward / backward on BASELINE configs[0] (2^18 points, L16 F2 T2^19, fp16 table)
  (a) uniformly random points as they come, (b) the same points sorted by the Morton code of their 1024^3 cell,
  (c) ray-ordered samples (what the training path feeds), and the cost of producing the order (code + radix sort + gather,
      and the inverse permutation of the outputs that a drop-in op would owe its caller)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from raw_ngp_b200 import raymarching, synthetic
from raw_ngp_b200.gridencoder import GridEncoder

dev = torch.device("cuda:0")
B = 1 << 18
enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).to(dev)
enc.embeddings.data = enc.embeddings.data.uniform_(-1e-4, 1e-4).half()
torch.manual_seed(0)
x_rand = torch.rand(B, 3, device=dev) * 2 - 1
g = torch.randn(B, 32, device=dev).half()

def order(x):
    q = ((x + 1) * 0.5 * 1023).clamp(0, 1023).int().contiguous()
    code = raymarching.morton3D(q)
    perm = torch.sort(code.long()).indices          # radix sort on the device
    return perm

def t(fn, it=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

def fb(x):
    xr = x.clone()
    def fwd():
        with torch.no_grad():
            return enc(xr, bound=1)
    def fwd_bwd():
        enc.embeddings.grad = None
        out = enc(xr, bound=1)
        out.backward(g)
    f = t(fwd); fbw = t(fwd_bwd)
    return f, fbw - f

perm = order(x_rand)
x_sorted = x_rand[perm].contiguous()
# ray-coherent samples: 4096 rays x 64 consecutive samples of the configs[1] marcher
model, o, d, _ = bench.build_scene(dev, 0)
nears, fars = synthetic.near_far_torch(o.to(dev), d.to(dev), model.aabb_train, 0.05)
xyzs, _, _, rays, _ = raymarching.march_rays_train(o.to(dev), d.to(dev), None, 1.0, False, model.density_bitfield, 1, 128, nears, fars, False, 0.0, 1024)
x_ray = xyzs[:B].contiguous()
res = {}
for name, x in (("random", x_rand), ("morton-sorted", x_sorted), ("ray-ordered", x_ray)):
    f, b = fb(x)
    res[name] = (f, b)
    print(f"{name:14s}: forward {f:7.1f} us ({B / f:6.0f} Mpts/s)   backward {b:7.1f} us ({B / b:6.0f} Mpts/s)")
t_order = t(lambda: order(x_rand))
t_gather = t(lambda: x_rand[perm].contiguous())
out = torch.empty(B, 32, device=dev, dtype=torch.float16)
inv = torch.empty_like(perm); inv[perm] = torch.arange(B, device=dev)
t_unperm = t(lambda: out[inv])
print(f"cost of the order: code + sort {t_order:.1f} us, gather inputs {t_gather:.1f} us, un-permute outputs [B,32] fp16 {t_unperm:.1f} us "
      f"(+ the same for the incoming gradient in backward)")
gain_f = res['random'][0] - res['morton-sorted'][0]; gain_b = res['random'][1] - res['morton-sorted'][1]
print(f"gain of sorting: forward {gain_f:.1f} us, backward {gain_b:.1f} us; overhead forward {t_order + t_gather + t_unperm:.1f} us")
