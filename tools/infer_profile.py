import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
os.environ.setdefault('INFER_CHUNK_LOG2','21')
import bench
from raw_ngp_b200 import raymarching
dev = torch.device('cuda',0)
model,_,_,_ = bench.build_scene(dev,0)
model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
model.eval()
W,H,f = 1920,1080,1200.0
j,i = torch.meshgrid(torch.arange(H,device=dev), torch.arange(W,device=dev), indexing='ij')
dirs = torch.stack([(i-W/2)/f, -(j-H/2)/f, -torch.ones_like(i,dtype=torch.float32)],-1).reshape(-1,3).contiguous()
rays_o = torch.tensor([0.,0.,2.],device=dev).expand_as(dirs).contiguous()
calls=[]
orig = raymarching.march_rays
def spy(n_alive, n_step, *a, **k):
    calls.append((n_alive, n_step)); return orig(n_alive, n_step, *a, **k)
raymarching.march_rays = spy
import raw_ngp_b200.nerf.renderer as R
R.raymarching.march_rays = spy
with torch.no_grad():
    model.render(rays_o, dirs, bg_color=1.0, perturb=False)
torch.cuda.synchronize()
print('iterations', len(calls), 'samples', sum(a*b for a,b in calls), 'first', calls[:6], 'last', calls[-3:])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.no_grad():
        model.render(rays_o, dirs, bg_color=1.0, perturb=False)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=14, max_name_column_width=60))

# per-call device time of the marcher (first call = all rays of the frame, most of them crossing empty space)
evs = []
def timed(n_alive, n_step, *a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(n_alive, n_step, *a, **k); e1.record()
    evs.append((n_alive, n_step, e0, e1)); return r
R.raymarching.march_rays = timed
with torch.no_grad():
    model.render(rays_o, dirs, bg_color=1.0, perturb=False)
torch.cuda.synchronize()
print('march per call (n_alive, n_step, us):', [(a, b, round(e0.elapsed_time(e1) * 1e3)) for a, b, e0, e1 in evs][:12], '...')
