"""CPU tests (no GPU): the parity oracle (oracle/*.py, oracle/raymarch_oracle.c) against the golden fixtures in
tests/golden/, which are outputs of the REFERENCE's own CUDA kernels (tools/make_golden.py, run on a B200) and of
the reference's SH source expressions evaluated verbatim (tools/eval_reference_sh_source.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import grid_oracle, raymarch_oracle, sh_oracle
from raw_ngp_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _table(n_entries, C, dtype):
    idx = torch.arange(n_entries * C, dtype=torch.float64)
    emb = (torch.frac(torch.sin(idx * 12.9898) * 43758.5453) * 2 - 1).reshape(n_entries, C)
    return emb.to(getattr(torch, dtype)).float().numpy()


GRID_FILES = sorted(glob.glob(os.path.join(GOLD, "grid_*.npz")))


@pytest.mark.parametrize("path", GRID_FILES, ids=[os.path.basename(p)[5:-4] for p in GRID_FILES])
def test_grid_oracle_matches_reference_kernels(path):
    z = np.load(path)
    D, C, L, H = int(z["D"]), int(z["C"]), int(z["L"]), int(z["base"])
    pls, gridtype, ac, interp = float(z["per_level_scale"]), int(z["gridtype"]), bool(z["align_corners"]), int(z["interp"])
    dtype = str(z["dtype"])
    half = dtype == "float16"
    offsets = z["offsets"]
    # host-side table layout (grid.py:124-134) and device-side resolutions (gridencoder.cu:133)
    assert np.array_equal(grid_oracle.table_offsets(D, L, pls, H, int(z["log2T"])), offsets)
    assert np.array_equal(grid_oracle.level_resolutions(L, pls, H), z["level_resolutions"])
    table = _table(int(offsets[-1]), C, dtype)
    out, dy_dx = grid_oracle.forward(z["inputs"], table, offsets, pls, H, gridtype, ac, interp, None, True, half)
    if half:
        np.testing.assert_allclose(out, z["outputs"], rtol=1e-3, atol=1e-3)
        assert (out == z["outputs"]).mean() > 0.99          # half accumulation restated bit for bit
        scale = np.abs(z["dy_dx"]).max()
        np.testing.assert_allclose(dy_dx / scale, z["dy_dx"] / scale, rtol=0, atol=2e-3)
    else:
        np.testing.assert_allclose(out, z["outputs"], rtol=1e-5, atol=1e-6)
        assert (out == z["outputs"]).mean() > 0.99
        scale = np.abs(z["dy_dx"]).max()
        np.testing.assert_allclose(dy_dx / scale, z["dy_dx"] / scale, rtol=0, atol=1e-6)
    B = z["inputs"].shape[0]
    gt = grid_oracle.backward(z["grad"], z["inputs"], table.shape, offsets, pls, H, gridtype, ac, interp, None, half)
    rows = z["grad_emb_rows"]
    ref_vals = z["grad_emb_vals"]
    scale = np.abs(ref_vals).max()
    np.testing.assert_allclose(gt[rows] / scale, ref_vals / scale, rtol=0, atol=5e-3 if half else 1e-5)
    untouched = np.ones(table.shape[0], bool)
    untouched[rows] = False
    assert np.abs(gt[untouched]).max() <= (1e-6 if half else 0.0) * scale + 1e-12
    gi = grid_oracle.input_backward(z["grad"], z["dy_dx"], B, D, C, L)
    scale = np.abs(z["grad_inputs"]).max()
    np.testing.assert_allclose(gi / scale, z["grad_inputs"] / scale, rtol=0, atol=5e-2 if half else 1e-5)


def test_sh_oracle_matches_reference_kernel_and_source():
    z = np.load(os.path.join(GOLD, "sh_deg8.npz"))
    Y, J = sh_oracle.sh_basis(z["inputs"], 8, jacobian=True)
    np.testing.assert_allclose(Y, z["outputs"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(J.reshape(Y.shape[0], -1), z["dy_dx"], rtol=1e-5, atol=1e-4)
    gi = sh_oracle.sh_backward(z["grad"], z["inputs"], 8)
    np.testing.assert_allclose(gi, z["grad_inputs"], rtol=1e-4, atol=1e-3)
    for deg in (1, 2, 4):   # lower degrees are prefixes of the same basis
        np.testing.assert_allclose(sh_oracle.sh_basis(z["inputs"], deg), z["outputs"][:, :deg * deg], rtol=1e-5, atol=1e-5)
    # the reference's source expressions, evaluated verbatim in float64
    s = np.load(os.path.join(GOLD, "sh_ref_source.npz"))
    Y, J = sh_oracle.sh_basis(s["inputs"], 8, jacobian=True)
    np.testing.assert_allclose(Y, s["outputs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(J[:, 0], s["dx"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(J[:, 1], s["dy"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(J[:, 2], s["dz"], rtol=1e-11, atol=1e-11)


def test_utils_oracle():
    z = np.load(os.path.join(GOLD, "utils.npz"))
    assert np.array_equal(raymarch_oracle.morton3D(z["coords"]), z["morton"])
    assert np.array_equal(raymarch_oracle.morton3D_invert(z["morton"]), z["morton_inv"])
    assert np.array_equal(z["morton_inv"], z["coords"])
    assert np.array_equal(synthetic.morton3d_torch(torch.from_numpy(z["coords"])).numpy(), z["morton"])
    np.testing.assert_allclose(raymarch_oracle.sph_from_ray(z["sph_o"], z["sph_d"], 1.5), z["sph_coords"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(raymarch_oracle.flatten_rays(z["flat_rays"], 10), z["flat"])


MARCH_FILES = sorted(glob.glob(os.path.join(GOLD, "march_*.npz")))


@pytest.mark.parametrize("path", MARCH_FILES, ids=[os.path.basename(p)[6:-4] for p in MARCH_FILES])
def test_raymarch_oracle_matches_reference_kernels(path):
    z = np.load(path)
    N, H, C = int(z["N"]), int(z["H"]), int(z["cascade"])
    bound, contract, dt_gamma, max_steps = float(z["bound"]), bool(z["contract"]), float(z["dt_gamma"]), int(z["max_steps"])
    radius = float(z["radius"]) if "radius" in z.files else 0.5
    grid = synthetic.ball_density_grid(H=H, cascade=C, bound=bound, radius=radius).numpy()
    # packbits + near/far: bit-exact
    assert np.array_equal(raymarch_oracle.packbits(grid, float(z["thresh"])), z["bitfield"])
    assert np.array_equal(synthetic.packbits_torch(torch.from_numpy(grid), float(z["thresh"])).numpy(), z["bitfield"])
    nears, fars = raymarch_oracle.near_far_from_aabb(z["rays_o"], z["rays_d"], z["aabb"], 0.05)
    assert np.array_equal(nears, z["nears"]) and np.array_equal(fars, z["fars"])
    # training march: counts bit-exact; samples per ray segment bit-exact (offsets of the reference are arrival-ordered)
    ldir = z["rays_ldir"] if z["rays_ldir"].size else None
    xyzs, dirs, ts, rays, ldirs = raymarch_oracle.march_rays_train(z["rays_o"], z["rays_d"], ldir, bound, contract, z["bitfield"],
                                                                   C, H, z["nears"], z["fars"], z["noises"], dt_gamma, max_steps)
    assert np.array_equal(rays[:, 1], z["rays"][:, 1])
    M = int(rays[:, 1].sum())
    assert M == z["xyzs"].shape[0]
    counts = rays[:, 1].astype(np.int64)
    ray_of = np.repeat(np.arange(N), counts)
    within = np.arange(M) - rays[:, 0].astype(np.int64)[ray_of]
    ref_pos = z["rays"][:, 0].astype(np.int64)[ray_of] + within
    assert np.array_equal(xyzs, z["xyzs"][ref_pos])
    assert np.array_equal(dirs, z["dirs"][ref_pos])
    assert np.array_equal(ts, z["ts"][ref_pos])
    if ldir is not None:
        assert np.array_equal(ldirs, z["ldirs"][ref_pos])
    # compositing on the reference's own sample layout (expf vs __expf: tolerance)
    T = float(z["T_thresh"])
    w, ws, dp, im = raymarch_oracle.composite_rays_train_forward(z["sigmas"], z["rgbs"], z["ts"], z["rays"], T)
    np.testing.assert_allclose(w, z["weights"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ws, z["weights_sum"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dp, z["depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(im, z["image"], rtol=1e-5, atol=1e-6)
    gs, gc = raymarch_oracle.composite_rays_train_backward(z["grad_weights"], z["grad_weights_sum"], z["grad_depth"],
                                                           z["grad_image"], z["sigmas"], z["rgbs"], z["ts"], z["rays"],
                                                           z["weights_sum"], z["depth"], z["image"], T)
    np.testing.assert_allclose(gc, z["grad_rgbs"], rtol=1e-5, atol=1e-6)
    scale = np.abs(z["grad_sigmas"]).max()
    np.testing.assert_allclose(gs / scale, z["grad_sigmas"] / scale, rtol=1e-4, atol=2e-6)
    # one inference iteration
    n_step = int(z["inf_n_step"])
    alive = np.arange(N, dtype=np.int32)
    ix, idr, its = raymarch_oracle.march_rays(N, n_step, alive, z["nears"], z["rays_o"], z["rays_d"], bound, contract,
                                              z["bitfield"], C, H, z["nears"], z["fars"], np.zeros(N, np.float32), dt_gamma, max_steps)
    assert np.array_equal(ix, z["inf_xyzs"]) and np.array_equal(idr, z["inf_dirs"]) and np.array_equal(its, z["inf_ts"])
    rays_t = z["nears"].copy()
    iws, idp, iim = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    raymarch_oracle.composite_rays(N, n_step, alive, rays_t, z["inf_sigmas"], z["inf_rgbs"], z["inf_ts"], iws, idp, iim, 1e-2)
    assert np.array_equal(alive, z["inf_alive"])
    np.testing.assert_allclose(rays_t, z["inf_rays_t"], rtol=0, atol=0)
    np.testing.assert_allclose(iws, z["inf_weights_sum"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(idp, z["inf_depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(iim, z["inf_image"], rtol=1e-5, atol=1e-6)


def test_freq_oracle_matches_reference_kernel_outputs():
    """tests/golden/freq.npz: outputs / input gradients of the reference's freqencoder kernels (tools/make_golden.py on a B200)."""
    from oracle import freq_oracle
    path = os.path.join(GOLD, "freq.npz")
    z = np.load(path)
    deg, D = int(z["degree"]), int(z["inputs"].shape[1])
    out = freq_oracle.forward(z["inputs"], deg)
    np.testing.assert_allclose(out, z["outputs"], rtol=0, atol=float(z["atol"]))          # __sinf vs np.sin
    assert np.array_equal(out[:, :D], z["outputs"][:, :D])                                   # the pass-through columns are exact
    np.testing.assert_allclose(freq_oracle.backward(z["grad"], z["outputs"], D, deg), z["grad_inputs"], rtol=1e-6, atol=1e-6)


def test_torch_near_far_matches_reference_renderer():
    """tests/golden/near_far_py.npz: nerf/renderer.py:139-158 run from the reference (tools/make_golden_state_dict.py), values
    and gradients; both restatements in this repo (the renderer's and the synthetic helper) must reproduce it exactly."""
    import torch
    from raw_ngp_b200 import synthetic
    from raw_ngp_b200.nerf.renderer import near_far_from_aabb
    z = np.load(os.path.join(GOLD, "near_far_py.npz"))
    for fn in (near_far_from_aabb, synthetic.near_far_torch):
        o = torch.from_numpy(z["rays_o"]).requires_grad_(True)
        d = torch.from_numpy(z["rays_d"]).requires_grad_(True)
        near, far = fn(o, d, torch.from_numpy(z["aabb"]), float(z["min_near"]))
        assert np.array_equal(near.detach().numpy(), z["nears"]) and np.array_equal(far.detach().numpy(), z["fars"])
        ((near * torch.from_numpy(z["g_near"])).sum() + (far * torch.from_numpy(z["g_far"])).sum()).backward()
        np.testing.assert_allclose(o.grad.numpy(), z["d_rays_o"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d.grad.numpy(), z["d_rays_d"], rtol=1e-6, atol=1e-7)


def _hdr_loss_torch(pred, gt, exposure):
    """the restatement of nerf/train_utils.py:529-536 used by the GPU tests (loss_weight 'none', lossmult 1)"""
    import torch
    clip = torch.minimum(torch.tensor(1.0), pred * exposure.unsqueeze(1))
    scaling = 1.0 / (1e-3 + clip.detach())
    return ((clip - gt) ** 2 * scaling ** 2).sum() / (3 * pred.shape[0])


def test_hdr_loss_restatement_matches_reference_lines():
    """tests/golden/hdr_loss.npz was produced by executing the reference's own lines (tools/make_golden_hdr.py)"""
    import torch
    z = np.load(os.path.join(GOLD, "hdr_loss.npz"))
    pred = torch.from_numpy(z["pred_rgb"]).requires_grad_(True)
    loss = _hdr_loss_torch(pred, torch.from_numpy(z["gt_rgb"]), torch.from_numpy(z["exposure"]))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(z["loss"]), rtol=1e-6)
    np.testing.assert_allclose(pred.grad.numpy(), z["d_pred"], rtol=1e-5, atol=1e-7)
