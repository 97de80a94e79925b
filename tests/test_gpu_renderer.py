"""GPU tests of the renderer-level rows of SURVEY 8 (a11, a12): update_extra_state, mark_untrained_grid, run_cuda.
The reference implements these as Python loops over torch ops (nerf/renderer.py:515-676, 716-897); the expected
values here are computed by restating those loops literally with torch + the reference's own morton3D / packbits
kernels (oracle/_ref), on the same model and the same torch RNG stream."""
import math

import numpy as np
import pytest
import torch

from raw_ngp_b200 import synthetic
from raw_ngp_b200.nerf import NeRFNetwork, default_opt

pytestmark = pytest.mark.gpu


def _model(**kw):
    torch.manual_seed(0)
    opt = default_opt(**kw)
    m = NeRFNetwork(opt).cuda()
    m.grid_encoder.embeddings.data.uniform_(-1.0, 1.0)     # make densities vary across space
    return m


def _reference_full_update(model, decay=0.95):
    """renderer.py:818-851 + :883-894, literally."""
    from oracle import ref_cuda
    H, dev = model.grid_size, model.density_grid.device
    density_grid = model.density_grid.clone()
    tmp_grid = -torch.ones_like(density_grid)
    ar = torch.arange(H, dtype=torch.int32, device=dev)
    xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
    indices = ref_cuda.morton3D(coords).long()
    xyzs = 2 * coords.float() / (H - 1) - 1
    for cas in range(model.cascade):
        bound = min(2 ** cas, model.bound)
        half_grid_size = bound / H
        cas_xyzs = xyzs * (bound - half_grid_size)
        cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
        with torch.amp.autocast("cuda", enabled=model.opt.fp16):
            sigmas = model.density(cas_xyzs)["sigma"].reshape(-1).detach()
        tmp_grid[cas, indices] = sigmas.float()
    valid = (density_grid >= 0) & (tmp_grid >= 0)
    density_grid[valid] = torch.maximum(density_grid[valid] * decay, tmp_grid[valid])
    mean = torch.mean(density_grid.clamp(min=0)).item()
    bitfield = ref_cuda.packbits(density_grid, min(mean, model.density_thresh))
    return density_grid, mean, bitfield


@pytest.mark.parametrize("bound,H", [(1, 64), (2, 32)])
def test_update_extra_state_full_matches_reference_loop(bound, H, ref_march):
    model = _model(bound=bound, grid_size=H, hashmap_size=15, hashgrid_resolution=256)
    with torch.no_grad():
        model.density_grid.uniform_(0, 5)
        model.density_grid[0, :100] = -1          # "untrained" cells stay -1 and never become occupied
    torch.manual_seed(42)
    exp_grid, exp_mean, exp_bits = _reference_full_update(model)
    torch.manual_seed(42)
    model.update_extra_state()
    assert model.iter_density == 1
    diff = (model.density_grid - exp_grid).abs()
    print("max |grid - expected| =", diff.max().item(), "mismatching cells =", (diff > 1e-6 + 1e-5 * exp_grid.abs()).sum().item())
    torch.testing.assert_close(model.density_grid, exp_grid, rtol=1e-5, atol=1e-6)
    assert (model.density_grid[0, :100] == -1).all()
    assert model.mean_density == pytest.approx(exp_mean, rel=1e-5)
    mism = (model.density_bitfield != exp_bits).sum().item()
    assert mism <= 2, f"{mism} bitfield bytes differ"     # a cell exactly at the threshold may flip


def test_update_extra_state_partial_invariants():
    model = _model(bound=1, grid_size=32, hashmap_size=14, hashgrid_resolution=128)
    model.iter_density = 16                                  # partial branch (renderer.py:854-880)
    with torch.no_grad():
        model.density_grid.uniform_(0, 5)
        model.density_grid[0, ::7] = -1
    before = model.density_grid.clone()
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("error")                  # the update must not synchronise with the host
    try:
        model.update_extra_state(decay=0.9)
    finally:
        torch.cuda.set_sync_debug_mode("default")
    after = model.density_grid
    assert (after[before < 0] == -1).all()
    changed = after != before
    # an updated cell is max(old*decay, sigma) >= old*decay; untouched cells are unchanged
    assert (after[changed] >= before[changed] * 0.9 - 1e-6).all()
    frac = changed.float().mean().item()
    assert 0.2 < frac < 0.6                                   # ~ H^3/2 samples with duplicates
    bits = model.density_bitfield
    th = min(model.mean_density, model.density_thresh)
    exp = synthetic.packbits_torch(after.cpu(), th)
    assert (bits.cpu() != exp).sum().item() <= 2


def test_update_extra_state_partial_with_empty_grid():
    """no occupied cell: the reference samples only the uniform cells (renderer.py:862-869); here the second half repeats them"""
    model = _model(bound=1, grid_size=32, hashmap_size=14, hashgrid_resolution=128)
    model.iter_density = 16
    with torch.no_grad():
        model.density_grid.zero_()
    model.update_extra_state(decay=0.9)
    after = model.density_grid
    assert torch.isfinite(after).all() and (after >= 0).all()
    frac = (after > 0).float().mean().item()
    assert 0.1 < frac < 0.3                                    # ~ H^3/4 distinct uniform cells were refreshed


def test_mark_untrained_grid_matches_reference_loop(ref_march):
    from oracle import ref_cuda
    model = _model(bound=2, grid_size=32, hashmap_size=14, hashgrid_resolution=128)
    g = torch.Generator().manual_seed(3)
    B = 6
    poses = torch.eye(4).repeat(B, 1, 1)
    for i in range(B):
        c = torch.randn(3, generator=g)
        c = c / c.norm() * 2.5
        fwd = -c / c.norm()
        up = torch.tensor([0.0, 1.0, 0.0])
        right = torch.linalg.cross(fwd, up); right = right / right.norm()
        up2 = torch.linalg.cross(right, fwd)
        poses[i, :3, 0], poses[i, :3, 1], poses[i, :3, 2], poses[i, :3, 3] = right, up2, -fwd, c

    class DS:
        pass
    ds = DS()
    ds.poses = poses.numpy()
    ds.intrinsics = np.array([400.0, 400.0, 200.0, 150.0])
    model.update_aabb(np.array([-1.5, -1.0, -2.0, 1.2, 2.0, 1.0], dtype=np.float32))
    model.mark_untrained_grid(ds, S=16)

    # literal restatement of renderer.py:716-809
    H, dev = model.grid_size, model.aabb_train.device
    fx, fy, cx, cy = ds.intrinsics
    mask_cam = torch.zeros_like(model.density_grid)
    mask_aabb = torch.zeros_like(model.density_grid)
    P = poses.to(dev)
    ar = torch.arange(H, dtype=torch.int32, device=dev)
    xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
    indices = ref_cuda.morton3D(coords).long()
    world = (2 * coords.float() / (H - 1) - 1).unsqueeze(0)
    for cas in range(model.cascade):
        bound = min(2 ** cas, model.bound)
        hgs = bound / H
        cw = world * (bound - hgs)
        mmin = (cw >= (model.aabb_train[:3] - hgs)).sum(-1) == 3
        mmax = (cw <= (model.aabb_train[3:] + hgs)).sum(-1) == 3
        mask_aabb[cas, indices] += (mmin & mmax).reshape(-1)
        cam = cw - P[:, :3, 3].unsqueeze(1)
        cam = cam @ P[:, :3, :3]
        cam[:, :, 2] *= -1
        mz = cam[:, :, 2] > model.opt.min_near
        mx = torch.abs(cam[:, :, 0]) < (cx / fx * cam[:, :, 2] + hgs * 2)
        my = torch.abs(cam[:, :, 1]) < (cy / fy * cam[:, :, 2] + hgs * 2)
        mask_cam[cas, indices] += (mz & mx & my).sum(0).bool().reshape(-1)
    expected = (mask_cam == 0) | (mask_aabb == 0)
    assert torch.equal(model.density_grid == -1, expected)
    assert 0.05 < expected.float().mean().item() < 0.99


def test_run_cuda_train_and_inference_agree():
    """The same rays rendered through the training path (march_rays_train + composite_rays_train) and through the
    inference loop (march_rays / composite_rays with compaction) give the same image."""
    model = _model(bound=1, grid_size=64, hashmap_size=15, hashgrid_resolution=256, T_thresh=1e-4)
    grid = synthetic.ball_density_grid(H=64, cascade=1).cuda()
    from raw_ngp_b200 import raymarching
    model.density_grid.copy_(grid)
    model.density_bitfield = raymarching.packbits(model.density_grid, 10.0, model.density_bitfield)
    o, d = synthetic.sphere_rays(3000, seed=5)
    o, d = o.cuda(), d.cuda()
    with torch.no_grad():
        model.train()
        tr = model.render(o, d, bg_color=1.0, perturb=False)
        model.eval()
        ev = model.render(o, d, bg_color=1.0, perturb=False)
    assert tr["num_points"] > 10000
    torch.testing.assert_close(tr["image"], ev["image"], rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(tr["depth"], ev["depth"], rtol=2e-3, atol=2e-3)
    assert ev["image"].shape == (3000, 3) and torch.isfinite(ev["image"]).all()


@pytest.mark.parametrize("kw", [dict(bound=1), dict(bound=2, contract=True), dict(bound=1, pose_opt="barf")],
                         ids=["bound1", "contract", "barf-window"])
def test_fast_inference_loop_matches_generic_loop(kw):
    """NeRFRenderer._march_composite_loop_fast (four direct launches per iteration on per-frame buffers) against the generic
    loop that mirrors renderer.py:588-616 op by op: same rays survive each iteration, same image / depth / weights."""
    from raw_ngp_b200 import raymarching
    model = _model(grid_size=64, hashmap_size=15, hashgrid_resolution=256, T_thresh=1e-3, **kw)
    model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.half()
    model.grid_encoder.embeddings.data.uniform_(-1.0, 1.0)          # densities large enough for early termination
    if kw.get("pose_opt") == "barf":
        model.update_annealing(0.3)
    grid = synthetic.ball_density_grid(H=64, cascade=model.cascade, bound=float(model.bound)).cuda()
    model.density_grid.copy_(grid)
    model.density_bitfield = raymarching.packbits(model.density_grid, 10.0, model.density_bitfield)
    o, d = synthetic.sphere_rays(5000, seed=6)
    o, d = o.cuda(), (d * 1.3).cuda()                              # un-normalised directions, like get_rays
    model.eval()
    with torch.no_grad():
        fast = model.render(o, d, bg_color=1.0, perturb=False)
        model.FAST_INFER = False
        slow = model.render(o, d, bg_color=1.0, perturb=False)
    assert model._fast_infer_args(None, "full") is not None
    for k in ("image", "depth"):
        torch.testing.assert_close(fast[k], slow[k], rtol=2e-3, atol=2e-3)
    # the iteration schedule (rows of the sample buffers, steps per ray and iteration) does not change a single bit: the
    # reference's schedule (N rows, <= 8 steps) against the default (4 N rows, <= 16 steps)
    model.FAST_INFER = True
    model.INFER_ROWS, model.INFER_MAX_STEP = 1, 8
    with torch.no_grad():
        ref_sched = model.render(o, d, bg_color=1.0, perturb=False)
    for k in ("image", "depth"):
        assert torch.equal(fast[k], ref_sched[k]), k
    assert (fast["image"] < 0.99).float().mean().item() > 0.1      # the ball is visible


def test_train_step_reduces_loss():
    """A few optimisation steps of the native TrainStep on a fixed batch lower the loss (end-to-end gradient check)."""
    from raw_ngp_b200 import raymarching
    from raw_ngp_b200.trainer import TrainStep
    torch.manual_seed(0)
    model = NeRFNetwork(default_opt(bound=1, grid_size=64, hashmap_size=15, hashgrid_resolution=256)).cuda()
    grid = synthetic.ball_density_grid(H=64, cascade=1).cuda()
    model.density_grid.copy_(grid)
    model.density_bitfield = raymarching.packbits(model.density_grid, 10.0, model.density_bitfield)
    step = TrainStep(model, lr=1e-2, table_dtype=torch.float16)
    o, d = synthetic.sphere_rays(2048, seed=2)
    o, d = o.cuda(), d.cuda()
    target = torch.rand(2048, 3, generator=torch.Generator().manual_seed(1)).cuda() * 0.2 + 0.4
    losses = [step.step(o, d, target, update_grid=False).item() for _ in range(30)]
    assert math.isfinite(losses[-1]) and losses[-1] < 0.8 * losses[0] and losses[-1] < losses[10] < losses[0], losses[::5]


@pytest.mark.parametrize("contract_on", [False, True], ids=["bound2", "contract"])
def test_proposal_path_matches_reference_run(contract_on):
    """NeRFRenderer.run (the proposal-network path, cuda_ray=False; nerf/proposal.py) against the REFERENCE's own renderer.py /
    network.py (unmodified, over its own kernels; oracle/ref_stack.py) on the same parameters and the same torch RNG stream:
    image, depth, proposal loss, gradients of all five parameter groups; staged inference."""
    import _refstep as R
    stack = R.stacks().get("ref")
    extra = dict(cuda_ray=False, num_steps=[64, 32, 16], background="white", lambda_proposal=1, lambda_distort=0, max_ray_batch=300,
                 bound=2, contract=contract_on, grid_size=32, hashmap_size=15, hashgrid_resolution=128)
    torch.manual_seed(1)
    ours = NeRFNetwork(default_opt(**extra)).cuda()
    with torch.no_grad():
        for n, p in ours.named_parameters():
            if n.endswith("embeddings"):
                p.uniform_(-0.5, 0.5)
    ref = stack.build_network(stack.make_opt(**extra)).cuda()
    res = ref.load_state_dict(ours.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    N = 700
    o, d = synthetic.sphere_rays(N, seed=5)
    o, d = o.cuda(), (d * 1.2).cuda()
    tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(9)).cuda()
    import contextlib
    for amp in (True, False):
        # fp16 autocast (how the trainer calls it): outputs and losses.  fp32: also the gradients -- the inter-level loss is a sum of
        # max(0, w - bound)^2 over a few violating intervals, so under fp16 its gradient flips with the last bit of a weight
        out = {}
        for name, m in (("ref", ref), ("ours", ours)):
            m.zero_grad(set_to_none=True)
            m.train()
            torch.manual_seed(21)
            ctx = torch.autocast("cuda", dtype=torch.float16) if amp else contextlib.nullcontext()
            with ctx:
                r = m.render(o, d, bg_color=1, perturb=True, update_proposal=True)
            loss = torch.nn.functional.mse_loss(r["image"].float(), tgt) + r["proposal_loss"]
            loss.backward()
            grads = {k: p.grad.detach().float().clone() for k, p in m.named_parameters() if p.grad is not None}
            m.eval()
            with torch.no_grad(), ctx:
                ev = m.render(o, d, bg_color=1, perturb=False)
            out[name] = (r, loss.detach(), grads, ev)
        (ra, la, ga, ea), (rb, lb, gb, eb) = out["ref"], out["ours"]
        tol = dict(rtol=2e-3, atol=2e-3) if amp else dict(rtol=1e-4, atol=1e-5)
        assert ra["num_points"] == rb["num_points"] == N * 16
        for k in ("image", "depth", "weights_sum"):
            torch.testing.assert_close(rb[k].float(), ra[k].float(), **tol)
            torch.testing.assert_close(eb[k].float(), ea[k].float(), **tol)
        torch.testing.assert_close(rb["proposal_loss"], ra["proposal_loss"], rtol=5e-3 if amp else 1e-4, atol=1e-6)
        torch.testing.assert_close(lb, la, rtol=2e-3 if amp else 1e-4, atol=1e-6)
        assert set(ga) == set(gb) and len(ga) == 13
        if not amp:
            for k in ga:
                scale = ga[k].abs().max().clamp(min=1e-12)
                err = (gb[k] - ga[k]).abs() / scale
                assert err.max().item() < 2e-3 and err.mean().item() < 1e-4, (k, err.max().item(), err.mean().item())


def test_distortion_loss_matches_pairwise_definition():
    """nerf/proposal.py: the O(T) prefix-sum form against the O(T^2) definition sum_ij w_i w_j |m_i - m_j| + 1/3 sum_i w_i^2 delta_i
    (what the reference's eff_distloss computes; that package is not installed, so the definition is the pin)."""
    from raw_ngp_b200.nerf import proposal as P
    torch.manual_seed(0)
    bins = torch.sort(torch.rand(64, 33, device="cuda"), dim=-1).values
    w = torch.rand(64, 32, device="cuda", requires_grad=True)
    m = (bins[..., 1:] + bins[..., :-1]) / 2
    dlt = bins[..., 1:] - bins[..., :-1]
    brute = ((w[..., :, None] * w[..., None, :] * (m[..., :, None] - m[..., None, :]).abs()).sum((-1, -2)) + (w ** 2 * dlt).sum(-1) / 3).mean()
    fast = P.distortion_loss(bins, w)
    torch.testing.assert_close(fast, brute, rtol=1e-5, atol=1e-7)
    g1, = torch.autograd.grad(fast, w, retain_graph=True)
    g2, = torch.autograd.grad(brute, w)
    torch.testing.assert_close(g1, g2, rtol=1e-4, atol=1e-6)
