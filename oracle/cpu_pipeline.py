"""TEST / BASELINE INFRASTRUCTURE ONLY -- the hot path restated in PyTorch on the CPU (vectorised), i.e. the
"PyTorch-on-CPU evaluation of the same grid-interpolation and compositing semantics" that BASELINE.json asks to
report next to the GPU numbers.  The reference itself has no CPU path (every op calls .cuda()), so this port is
what bench.py times as cpu_baseline / --impl reference (kind "port").  Never imported by raw_ngp_b200/.

Semantics: grid encoder gridencoder.cu:82-249 (autograd supplies :252-378), SH shencoder.cu:43-121, MLP
network.py:12-35, trunc_exp activation.py:9-21, march raymarching.cu:337-491 (C oracle), composite
raymarching.cu:519-597 (autograd supplies :623-712), Adam main.py:245.
"""
import math
from math import factorial, pi, sqrt

import numpy as np
import torch
import torch.nn.functional as F

from . import grid_oracle, raymarch_oracle

_PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]


def _index(gridtype, hs, res, pos):
    """pos [B, D] int64 -> entry row [B] (uint32 arithmetic emulated in int64).  gridencoder.cu:61-79"""
    D = pos.shape[1]
    stride, idx = 1, torch.zeros(pos.shape[0], dtype=torch.int64)
    for d in range(D):
        if stride <= hs:
            idx = (idx + pos[:, d] * stride) & 0xFFFFFFFF
            stride = (stride * res) & 0xFFFFFFFF
    if gridtype == 0 and stride > hs:
        idx = torch.zeros(pos.shape[0], dtype=torch.int64)
        for d in range(D):
            idx = idx ^ ((pos[:, d] * _PRIMES[d]) & 0xFFFFFFFF)
    return idx % hs


def grid_encode(x, table, offsets, per_level_scale, H, gridtype=0, align_corners=False, interp=0):
    """x [B, D] in [0,1] (fp32), table [sO, C] -> [B, L*C]; differentiable w.r.t. table and x."""
    B, D = x.shape
    L = len(offsets) - 1
    res_all = grid_oracle.level_resolutions(L, per_level_scale, H)
    inb = ((x >= 0) & (x <= 1)).all(dim=1, keepdim=True)
    outs = []
    for l in range(L):
        res = int(res_all[l])
        hs = int(offsets[l + 1] - offsets[l])
        lvl = table[int(offsets[l]):int(offsets[l + 1])]
        if align_corners:
            pos = x * (res - 1)
            base = torch.clamp(torch.floor(pos.detach()).long(), max=res - 2)
        else:
            pos = torch.clamp(x * res - 0.5, 0.0, float(res - 1))
            base = torch.floor(pos.detach()).long()
        frac = pos - base.float()
        if interp == 1:
            frac = frac * frac * (3.0 - 2.0 * frac)
        acc = 0
        for k in range(1 << D):
            p = torch.stack([torch.clamp(base[:, d] + 1, max=res - 1) if (k >> d) & 1 else base[:, d] for d in range(D)], 1)
            w = 1
            for d in range(D):
                w = w * (frac[:, d] if (k >> d) & 1 else 1 - frac[:, d])
            acc = acc + w.unsqueeze(1) * lvl.index_select(0, _index(gridtype, hs, res, p.clamp(min=0)))
        outs.append(torch.where(inb, acc, torch.zeros_like(acc)))
    return torch.cat(outs, dim=1)


def _legendre_derivs(lmax):
    from numpy.polynomial import legendre as npleg
    from numpy.polynomial import polynomial as nppoly
    return {(l, a): (nppoly.polyder(npleg.leg2poly([0] * l + [1]), a) if a else npleg.leg2poly([0] * l + [1]))
            for l in range(lmax) for a in range(l + 1)}


def sh_encode(d, degree=4):
    """Real SH basis [B, degree^2] with torch ops (differentiable).  Same polynomials as shencoder.cu:43-121."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    q = _legendre_derivs(degree)
    c, s = [torch.ones_like(x)], [torch.zeros_like(x)]
    for m in range(1, degree):
        c.append(x * c[-1] - y * s[-1])
        s.append(x * s[-2 + 1] + y * c[-2]) if False else s.append(x * s[m - 1] + y * c[m - 1])
    out = []
    for l in range(degree):
        for m in range(-l, l + 1):
            a = abs(m)
            K = sqrt((2 * l + 1) / (4 * pi) * factorial(l - a) / factorial(l + a))
            k = K if m == 0 else (-1) ** a * sqrt(2.0) * K
            poly = 0
            for i, co in enumerate(q[(l, a)]):
                if co != 0:
                    poly = poly + float(co) * z ** i
            if not torch.is_tensor(poly):
                poly = torch.full_like(z, float(poly))
            A = c[a] if m >= 0 else s[a]
            out.append(k * A * poly)
    return torch.stack(out, dim=1)


def composite_train(sigmas, rgbs, ts, rays, T_thresh=1e-4):
    """Differentiable compositing (raymarching.cu:519-597) on a padded [N, Lmax] layout."""
    N = rays.shape[0]
    counts = rays[:, 1].long()
    offs = rays[:, 0].long()
    Lmax = int(counts.max().item()) if N > 0 else 0
    if Lmax == 0:
        z = torch.zeros(N)
        return torch.zeros_like(sigmas), z, z.clone(), torch.zeros(N, 3)
    ar = torch.arange(Lmax).unsqueeze(0)
    valid = ar < counts.unsqueeze(1)
    idx = (offs.unsqueeze(1) + ar).clamp(max=max(sigmas.shape[0] - 1, 0))
    sg = torch.where(valid, sigmas[idx], torch.zeros(()))
    dt = ts[:, 1][idx]
    tt = ts[:, 0][idx]
    col = rgbs[idx]
    alpha = torch.where(valid, 1 - torch.exp(-sg * dt), torch.zeros(()))
    T_after = torch.cumprod(1 - alpha, dim=1)
    T_before = torch.cat([torch.ones(N, 1), T_after[:, :-1]], dim=1)
    # samples after the first one whose outgoing T < T_thresh are dropped
    dead_before = torch.cat([torch.zeros(N, 1, dtype=torch.bool), (T_after < T_thresh)[:, :-1]], dim=1)
    live = valid & ~(torch.cumsum(dead_before.long(), dim=1) > 0)
    w = torch.where(live, alpha * T_before, torch.zeros(()))
    weights = torch.zeros_like(sigmas).index_put((idx[live],), w[live])
    return weights, w.sum(1), (w * tt).sum(1), (w.unsqueeze(-1) * col).sum(1)


class CpuNeRF(torch.nn.Module):
    """Config-2 network on the CPU: hash grid (fp32) -> 32-64-64-16 -> sigma, SH(4)+15 -> 64-64-3 -> clamped exp."""

    def __init__(self, bound=1, log2T=19, desired=2048, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.bound = bound
        self.pls = float(np.exp2(np.log2(desired * bound / 16) / 15))
        self.offsets = grid_oracle.table_offsets(3, 16, self.pls, 16, log2T)
        self.table = torch.nn.Parameter((torch.rand(int(self.offsets[-1]), 2, generator=g) * 2 - 1) * 1e-4)

        def lin(i, o):
            w = torch.empty(o, i)
            torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5), generator=g)
            return torch.nn.Parameter(w)
        self.g = torch.nn.ParameterList([lin(32, 64), lin(64, 64), lin(64, 16)])
        self.v = torch.nn.ParameterList([lin(31, 64), lin(64, 64), lin(64, 3)])

    def forward(self, xyzs, dirs):
        f = grid_encode((xyzs + self.bound) / (2 * self.bound), self.table, self.offsets, self.pls, 16)
        h = F.linear(F.relu(F.linear(F.relu(F.linear(f, self.g[0])), self.g[1])), self.g[2])
        sigma = torch.exp(h[:, 0])
        e = sh_encode(dirs / dirs.norm(dim=-1, keepdim=True), 4)
        c = torch.cat([h[:, 1:], e], dim=-1)
        c = F.linear(F.relu(F.linear(F.relu(F.linear(c, self.v[0])), self.v[1])), self.v[2])
        return sigma, torch.clamp(torch.exp(c - 5.0), max=5.0)


def train_step(model, optimizer, rays_o, rays_d, target, bitfield, nears, fars, noises, bound=1.0, C=1, H=128,
               max_steps=1024, T_thresh=1e-8, bg=1.0):
    """One CPU training step on the given rays; returns (loss, M)."""
    xyzs, dirs, ts, rays, _ = raymarch_oracle.march_rays_train(rays_o.numpy(), rays_d.numpy(), None, bound, False,
                                                               bitfield.numpy(), C, H, nears.numpy(), fars.numpy(),
                                                               noises.numpy(), 0.0, max_steps)
    xyzs, dirs, ts, rays = map(torch.from_numpy, (xyzs, dirs, ts, rays))
    sigma, rgb = model(xyzs, dirs)
    _, ws, _, img = composite_train(sigma, rgb, ts, rays, T_thresh)
    img = img + (1 - ws).unsqueeze(-1) * bg
    loss = F.mse_loss(img, target, reduction="none").mean(-1).mean()
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    optimizer.step()
    return float(loss.item()), int(xyzs.shape[0])
