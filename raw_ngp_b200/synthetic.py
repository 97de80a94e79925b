"""Seeded synthetic inputs of the BASELINE.json configurations.

Everything here only *generates inputs* (rays, occupancy grids, tables); it is shared by tests/ and bench.py so
that both sides of a parity comparison see the same bytes.  Pure torch, device-agnostic.
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def morton3d_torch(coords):
    """10-bit Morton interleave with torch integer ops (input generator only; the operator is raymarching.morton3D)."""
    def spread(v):
        v = v.long() & 0x3FF
        v = (v | (v << 16)) & 0xFF0000FF
        v = (v | (v << 8)) & 0x0F00F00F
        v = (v | (v << 4)) & 0xC30C30C3
        v = (v | (v << 2)) & 0x49249249
        return v
    return (spread(coords[..., 0]) | (spread(coords[..., 1]) << 1) | (spread(coords[..., 2]) << 2)).to(torch.int32)


def ball_density_grid(H=128, cascade=1, bound=1.0, radius=0.5, sigma=50.0):
    """density_grid [cascade, H^3] (Morton order): sigma inside the ball |x| < radius, 0 outside (config 2)."""
    ar = torch.arange(H, dtype=torch.int32)
    xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
    coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
    idx = morton3d_torch(coords).long()
    unit = 2 * coords.float() / (H - 1) - 1
    grid = torch.zeros(cascade, H ** 3)
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        world = unit * (b - b / H)
        grid[cas, idx] = torch.where(world.norm(dim=-1) < radius, torch.tensor(float(sigma)), torch.tensor(0.0))
    return grid


def packbits_torch(grid, thresh):
    """Reference semantics of packbits in torch (input generator for CPU-side tests)."""
    bits = (grid.reshape(-1, 8) > thresh).to(torch.uint8)
    w = (2 ** torch.arange(8, dtype=torch.uint8)).to(torch.uint8)
    return (bits * w).sum(dim=-1).to(torch.uint8)


def sphere_rays(N, seed=2, origin_radius=2.0, target_radius=0.6):
    """Origins uniform on the sphere r = origin_radius, targets uniform in the ball r = target_radius,
    rays_d = normalised (target - origin).  Returns rays_o, rays_d [N, 3] fp32 (CPU)."""
    g = _gen(seed)
    o = torch.randn(N, 3, generator=g)
    o = o / o.norm(dim=-1, keepdim=True) * origin_radius
    t = torch.randn(N, 3, generator=g)
    t = t / t.norm(dim=-1, keepdim=True)
    r = torch.rand(N, 1, generator=g) ** (1.0 / 3.0) * target_radius
    t = t * r
    d = t - o
    d = d / d.norm(dim=-1, keepdim=True)
    return o.contiguous(), d.contiguous()


def unit_vectors(N, seed=3):
    g = _gen(seed)
    v = torch.randn(N, 3, generator=g)
    return (v / v.norm(dim=-1, keepdim=True)).contiguous()


def near_far_torch(rays_o, rays_d, aabb, min_near=0.05):
    """Differentiable slab test of the renderer (nerf/renderer.py:139-158); returns [N, 1] tensors."""
    tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
    tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
    far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
    mask = far < near
    near = torch.where(mask, torch.full_like(near, 1e9), near)
    far = torch.where(mask, torch.full_like(far, 1e9), far)
    near = torch.clamp(near, min=min_near)
    return near, far


def pinhole_rays(W=1920, H=1080, fx=1200.0, fy=1200.0, radius=2.0):
    """Full-frame rays of config 4: camera at (0, 0, radius) looking at the origin (-z forward)."""
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    x = (i + 0.5 - W / 2) / fx
    y = -(j + 0.5 - H / 2) / fy
    d = torch.stack([x, y, -torch.ones_like(x)], dim=-1).reshape(-1, 3)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor([0.0, 0.0, radius]).expand_as(d).contiguous()
    return o, d.contiguous()


def uniform_points(B, seed=0, lo=-1.0, hi=1.0):
    g = _gen(seed)
    return (torch.rand(B, 3, generator=g) * (hi - lo) + lo).contiguous()


def table_values(n_entries, C, seed=0, scale=1.0, dtype=torch.float32):
    g = _gen(seed + 1000)
    return ((torch.rand(n_entries, C, generator=g) * 2 - 1) * scale).to(dtype).contiguous()
