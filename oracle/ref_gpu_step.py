"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's own training step on the GPU, for timing next to ours.

One step of `Trainer.train_step` (nerf/train_utils.py:481-568, 863-930) with `cuda_ray=True`, restated around the
reference's *unmodified* CUDA extensions (oracle/_ref, via oracle/ref_cuda.py) and the PyTorch pieces the reference uses:
torch near_far_from_aabb (nerf/renderer.py:139-158), `nn.Linear` MLPs under `torch.autocast` (nerf/network.py:12-35,
renderer.py:546), trunc_exp (activation.py:9-21), GradScaler + `torch.optim.Adam(eps=1e-15)` (main.py:245,
train_utils.py:897-904).  The hash table is an fp32 parameter, as in this fork (gridencoder/grid.py:43-46).  Nothing here is
part of the product path.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ref_cuda


class _GridEncode(Function):            # gridencoder/grid.py:24-95
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution):
        out, _ = ref_cuda.grid_forward(inputs, embeddings, offsets, per_level_scale, base_resolution)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.meta = (per_level_scale, base_resolution)
        return out

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        _, g = ref_cuda.grid_backward(grad.contiguous().to(embeddings.dtype), inputs, embeddings, offsets, *ctx.meta)
        return None, g, None, None, None


class _TruncExp(Function):              # activation.py:9-21
    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-80, 80))


class _Composite(Function):             # raymarching/raymarching.py:333-390
    @staticmethod
    def forward(ctx, sigmas, rgbs, ts, rays, T_thresh):
        sigmas, rgbs = sigmas.float().contiguous(), rgbs.float().contiguous()
        weights, ws, depth, image = ref_cuda.composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh)
        ctx.save_for_backward(sigmas, rgbs, ts, rays, ws, depth, image)
        ctx.T = T_thresh
        return weights, ws, depth, image

    @staticmethod
    def backward(ctx, gw, gws, gd, gi):
        sigmas, rgbs, ts, rays, ws, depth, image = ctx.saved_tensors
        gs, gr = ref_cuda.composite_rays_train_backward(gw.contiguous(), gws.contiguous(), gd.contiguous(), gi.contiguous(),
                                                        sigmas, rgbs, ts, rays, ws, depth, image, ctx.T)
        return gs, gr, None, None, None


def _near_far(rays_o, rays_d, aabb, min_near):   # nerf/renderer.py:139-158
    tmin = (aabb[:3] - rays_o) / (rays_d + 1e-15)
    tmax = (aabb[3:] - rays_o) / (rays_d + 1e-15)
    near = torch.where(tmin < tmax, tmin, tmax).amax(dim=-1, keepdim=True)
    far = torch.where(tmin > tmax, tmin, tmax).amin(dim=-1, keepdim=True)
    mask = far < near
    near = torch.where(mask, torch.full_like(near, 1e9), near)
    far = torch.where(mask, torch.full_like(far, 1e9), far)
    return torch.clamp(near, min=min_near), far


class RefNeRF(nn.Module):               # nerf/network.py:37-143 (hashgrid 16x2, T=2^19, 2048; 64-wide bias-free MLPs)
    def __init__(self, offsets, per_level_scale, base_resolution=16, bound=1.0):
        super().__init__()
        self.register_buffer("offsets", offsets.int())
        self.per_level_scale, self.base_resolution, self.bound = per_level_scale, base_resolution, bound
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), 2).uniform_(-1e-4, 1e-4))
        self.grid_mlp = nn.ModuleList([nn.Linear(32, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 16, bias=False)])
        self.view_mlp = nn.ModuleList([nn.Linear(31, 64, bias=False), nn.Linear(64, 64, bias=False), nn.Linear(64, 3, bias=False)])

    @staticmethod
    def _mlp(net, x):
        for i, l in enumerate(net):
            x = l(x)
            if i + 1 < len(net):
                x = torch.relu_(x)
        return x

    def forward(self, x, d):
        x01 = (x + self.bound) / (2 * self.bound)
        f = _GridEncode.apply(x01.float().contiguous(), self.embeddings, self.offsets, self.per_level_scale, self.base_resolution)
        h = self._mlp(self.grid_mlp, f)
        sigma = _TruncExp.apply(h[..., 0])
        sh, _ = ref_cuda.sh_forward((d / torch.norm(d, dim=-1, keepdim=True)).float().contiguous(), 4)
        c = self._mlp(self.view_mlp, torch.cat([h[..., 1:], sh.to(h.dtype)], dim=-1))
        color = torch.clamp(torch.exp(c - 5.0), max=5.0)
        return sigma, color


class RefTrainStep:
    def __init__(self, model, bitfield, aabb, bound=1.0, cascade=1, grid_size=128, min_near=0.05, max_steps=1024, T_thresh=1e-8,
                 lr=1e-2, bg_color=1.0):
        self.m, self.bitfield, self.aabb = model, bitfield, aabb
        self.cfg = (bound, cascade, grid_size, min_near, max_steps, T_thresh, bg_color)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.99), eps=1e-15)
        self.scaler = torch.amp.GradScaler("cuda")
        self.num_points = 0

    def step(self, rays_o, rays_d, target):
        bound, C, H, min_near, max_steps, T_thresh, bg = self.cfg
        self.opt.zero_grad()
        nears, fars = _near_far(rays_o, rays_d, self.aabb, min_near)
        noises = torch.rand(rays_o.shape[0], device=rays_o.device)
        xyzs, dirs, ts, rays, _ = ref_cuda.march_rays_train(rays_o, rays_d, None, bound, False, self.bitfield, C, H,
                                                            nears.view(-1).contiguous(), fars.view(-1).contiguous(), noises, 0, max_steps)
        self.num_points = xyzs.shape[0]
        dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
        with torch.autocast("cuda", dtype=torch.float16):
            sigma, color = self.m(xyzs, dirs)
        _, ws, _, image = _Composite.apply(sigma, color, ts, rays, T_thresh)
        image = image + (1 - ws).unsqueeze(-1) * bg
        loss = torch.nn.functional.mse_loss(image, target, reduction="none").mean(-1).mean()
        self.scaler.scale(loss).backward()
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss.detach()
