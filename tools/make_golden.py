#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own CUDA extensions (oracle/_ref, unmodified sources
compiled by oracle/build_ref.sh) on small seeded inputs.  Must run on a GPU box:

    gpurun -- python tools/make_golden.py gpurun_out/golden      # then copy gpurun_out/golden/*.npz to tests/golden/

The fixtures pin the CPU oracle (oracle/*.py, oracle/raymarch_oracle.c): tests/test_oracle_golden.py checks the
oracle against them without a GPU.  Inputs are stored next to outputs so the fixtures are self-contained.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_cuda  # noqa: E402
from raw_ngp_b200 import _lib, synthetic  # noqa: E402
from raw_ngp_b200.gridencoder.grid import level_table_offsets  # noqa: E402


def npy(t):
    return t.detach().cpu().numpy()


def grid_case(name, out_dir, D=3, C=2, L=16, log2T=19, base=16, desired=2048, dtype=torch.float32, B=160, gridtype=0,
              align_corners=False, interp=0, seed=0):
    pls = np.exp2(np.log2(desired / base) / (L - 1))
    offsets = torch.tensor(level_table_offsets(D, L, pls, base, log2T), dtype=torch.int32)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, D, generator=g)
    x[:8] = x[:8] * 1.5 - 0.25
    x[8:12] = 0.0
    x[12:16] = 1.0
    n_entries = int(offsets[-1])
    # a table that is a deterministic function of the entry index, so the fixture need not store 6M entries
    idx = torch.arange(n_entries * C, dtype=torch.float64)
    emb = (torch.frac(torch.sin(idx * 12.9898) * 43758.5453) * 2 - 1).reshape(n_entries, C).to(dtype)
    grad = (torch.randn(B, L * C, generator=g) * 1e-2).to(dtype)
    xc, ec, oc, gc = x.cuda(), emb.cuda(), offsets.cuda(), grad.cuda()
    out, dy_dx = ref_cuda.grid_forward(xc, ec, oc, pls, base, True, gridtype, align_corners, interp, None)
    gx, ge = ref_cuda.grid_backward(gc, xc, ec, oc, pls, base, dy_dx, gridtype, align_corners, interp, None)
    res = torch.zeros(L, dtype=torch.int32, device="cuda")
    _lib.call("ngp_grid_level_resolutions", L, float(np.log2(pls)), base, res.data_ptr(), _lib.stream())
    nz = torch.nonzero(ge.float().abs().sum(-1)).squeeze(-1)
    np.savez_compressed(os.path.join(out_dir, f"grid_{name}.npz"), D=D, C=C, L=L, log2T=log2T, base=base,
                        per_level_scale=pls, gridtype=gridtype, align_corners=align_corners, interp=interp,
                        dtype=str(dtype).split(".")[-1], offsets=offsets.numpy(), inputs=x.numpy(),
                        grad=npy(gc.float()), outputs=npy(out.float()), dy_dx=npy(dy_dx.float()),
                        grad_inputs=npy(gx.float()), grad_emb_rows=npy(nz.int()), grad_emb_vals=npy(ge[nz].float()),
                        level_resolutions=npy(res))


def sh_case(out_dir):
    d = synthetic.unit_vectors(128, seed=3)
    d[:3] = torch.eye(3)
    d[3:6] = -torch.eye(3)
    d[6:40] *= 0.8
    dc = d.cuda()
    out, jac = ref_cuda.sh_forward(dc, 8, True)
    grad = torch.randn(128, 64, generator=torch.Generator().manual_seed(7)).cuda()
    gin = ref_cuda.sh_backward(grad, dc, 8, jac)
    np.savez_compressed(os.path.join(out_dir, "sh_deg8.npz"), inputs=d.numpy(), outputs=npy(out), dy_dx=npy(jac),
                        grad=npy(grad), grad_inputs=npy(gin))


def march_case(name, out_dir, N=96, H=32, cascade=1, bound=1.0, contract=False, dt_gamma=0.0, max_steps=256, ldir=False,
               radius=0.5, seed=2):
    grid = synthetic.ball_density_grid(H=H, cascade=cascade, bound=bound, radius=radius)
    thresh = min(grid.clamp(min=0).mean().item(), 10.0)
    gridc = grid.cuda()
    bitfield = ref_cuda.packbits(gridc, thresh)
    o, d = synthetic.sphere_rays(N, seed=seed)
    d[: N // 4] *= 2.5
    oc, dc = o.cuda(), d.cuda()
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32).cuda()
    nears, fars = ref_cuda.near_far_from_aabb(oc, dc, aabb, 0.05)
    noises = torch.rand(N, generator=torch.Generator().manual_seed(seed + 1)).cuda()
    l = synthetic.unit_vectors(N, seed=3).cuda() if ldir else None
    xyzs, dirs, ts, rays, ldirs = ref_cuda.march_rays_train(oc, dc, l, bound, contract, bitfield, cascade, H, nears, fars,
                                                            noises, dt_gamma, max_steps)
    M = xyzs.shape[0]
    g = torch.Generator().manual_seed(seed + 2)
    sigmas = (torch.rand(M, generator=g) * 30).cuda()
    rgbs = torch.rand(M, 3, generator=g).cuda()
    T_thresh = 1e-4
    w, ws, dp, im = ref_cuda.composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh)
    gw = (torch.randn(M, generator=g) * 0.1).cuda()
    gws = torch.randn(N, generator=g).cuda()
    gdp = torch.randn(N, generator=g).cuda()
    gim = torch.randn(N, 3, generator=g).cuda()
    gs, gc = ref_cuda.composite_rays_train_backward(gw, gws, gdp, gim, sigmas, rgbs, ts, rays, ws, dp, im, T_thresh)

    # one inference iteration: n_step = 4 from rays_t = nears
    n_step = 4
    alive = torch.arange(N, dtype=torch.int32).cuda()
    rays_t = nears.clone()
    zn = torch.zeros(N).cuda()
    ix, idr, its = ref_cuda.march_rays(N, n_step, alive, rays_t, oc, dc, bound, contract, bitfield, cascade, H, nears, fars,
                                       zn, dt_gamma, max_steps)
    isig = (torch.rand(N * n_step, generator=g) * 30).cuda()
    irgb = torch.rand(N * n_step, 3, generator=g).cuda()
    iws, idp, iim = torch.zeros(N).cuda(), torch.zeros(N).cuda(), torch.zeros(N, 3).cuda()
    alive2, t2 = alive.clone(), rays_t.clone()
    ref_cuda.composite_rays(N, n_step, alive2, t2, isig, irgb, its, iws, idp, iim, 1e-2)

    np.savez_compressed(
        os.path.join(out_dir, f"march_{name}.npz"), N=N, H=H, cascade=cascade, bound=bound, contract=contract,
        dt_gamma=dt_gamma, max_steps=max_steps, thresh=thresh, T_thresh=T_thresh, radius=radius,
        bitfield=npy(bitfield), rays_o=o.numpy(), rays_d=d.numpy(), aabb=npy(aabb), nears=npy(nears), fars=npy(fars),
        noises=npy(noises), rays_ldir=npy(l) if ldir else np.zeros(0), xyzs=npy(xyzs), dirs=npy(dirs), ts=npy(ts), rays=npy(rays),
        ldirs=npy(ldirs) if ldir else np.zeros(0), sigmas=npy(sigmas), rgbs=npy(rgbs), weights=npy(w), weights_sum=npy(ws),
        depth=npy(dp), image=npy(im), grad_weights=npy(gw), grad_weights_sum=npy(gws), grad_depth=npy(gdp),
        grad_image=npy(gim), grad_sigmas=npy(gs), grad_rgbs=npy(gc), inf_n_step=n_step, inf_xyzs=npy(ix), inf_dirs=npy(idr),
        inf_ts=npy(its), inf_sigmas=npy(isig), inf_rgbs=npy(irgb), inf_weights_sum=npy(iws), inf_depth=npy(idp),
        inf_image=npy(iim), inf_alive=npy(alive2), inf_rays_t=npy(t2))


def util_case(out_dir):
    g = torch.Generator().manual_seed(11)
    coords = torch.randint(0, 1024, (4096, 3), generator=g, dtype=torch.int32)
    idx = ref_cuda.morton3D(coords.cuda())
    inv = ref_cuda.morton3D_invert(idx)
    o, d = synthetic.sphere_rays(512, seed=5)
    coordsph = ref_cuda.sph_from_ray((o * 0.2).cuda(), d.cuda(), 1.5)
    rays = torch.tensor([[0, 3], [3, 0], [3, 5], [8, 2]], dtype=torch.int32)
    flat = ref_cuda.flatten_rays(rays.cuda(), 10)
    np.savez_compressed(os.path.join(out_dir, "utils.npz"), coords=coords.numpy(), morton=npy(idx), morton_inv=npy(inv),
                        sph_o=(o * 0.2).numpy(), sph_d=d.numpy(), sph_coords=npy(coordsph), flat_rays=rays.numpy(),
                        flat=npy(flat))


def freq_case(out_dir):
    """freqencoder kernels (unmodified reference source): outputs + input gradients on 96 points, degree 6"""
    g = torch.Generator().manual_seed(5)
    x = torch.rand(96, 3, generator=g) * 2 - 1
    x[:3] = torch.tensor([[0.0, 1.0, -1.0], [0.5, -0.5, 0.25], [1e-3, -1e-3, 0.999]])
    deg = 6
    xc = x.cuda()
    out = ref_cuda.freq_forward(xc, deg)
    grad = torch.randn(96, out.shape[1], generator=g)
    gx = ref_cuda.freq_backward(grad.cuda(), out, 3, deg)
    # __sinf against the correctly rounded sine: the absolute error grows with |argument| (here <= 2^5 + pi/2)
    np.savez_compressed(os.path.join(out_dir, "freq.npz"), degree=deg, inputs=x.numpy(), outputs=npy(out), grad=grad.numpy(),
                        grad_inputs=npy(gx), atol=2e-5)


def main():
    if len(sys.argv) > 2 and sys.argv[2] == "freq":      # only the frequency-encoder fixture
        os.makedirs(sys.argv[1], exist_ok=True)
        freq_case(sys.argv[1])
        return
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    assert torch.cuda.is_available(), "make_golden.py runs the reference CUDA kernels and needs a GPU"
    grid_case("fp32_model", out_dir)
    grid_case("fp16_model", out_dir, dtype=torch.float16)
    grid_case("fp32_4096", out_dir, desired=4096, B=128)
    grid_case("fp32_smooth_align_tiled", out_dir, L=8, log2T=15, desired=256, gridtype=1, align_corners=True, interp=1)
    grid_case("fp32_d2_c4", out_dir, D=2, C=4, L=8, log2T=14, desired=512)
    sh_case(out_dir)
    freq_case(out_dir)
    march_case("ball_h32", out_dir)
    march_case("cone_h32", out_dir, dt_gamma=1 / 64)
    march_case("contract_c2", out_dir, N=32, cascade=2, bound=2.0, contract=True, ldir=True, max_steps=256)
    march_case("cascade3", out_dir, N=48, cascade=3, bound=4.0, ldir=True, max_steps=512)
    util_case(out_dir)
    print("golden fixtures written to", out_dir, sorted(os.listdir(out_dir)))


if __name__ == "__main__":
    main()
