"""GPU tests: FusedTrainStep (sync-free native step, CUDA graph) against the autograd path built from the individually
verified operators (march_rays_train -> NeRFNetwork -> composite_rays_train -> MSE -> backward)."""
import copy

import pytest
import torch

from raw_ngp_b200 import pose, raymarching, synthetic
from raw_ngp_b200.nerf import NeRFNetwork, default_opt
from raw_ngp_b200.trainer import FusedTrainStep, TrainStep

pytestmark = pytest.mark.gpu


def _scene(N, **kw):
    torch.manual_seed(0)
    cfg = dict(bound=1, grid_size=64, max_steps=256, hashmap_size=15, hashgrid_resolution=256)
    cfg.update(kw)
    opt = default_opt(**cfg)
    model = NeRFNetwork(opt).cuda()
    model.grid_encoder.embeddings.data.uniform_(-0.5, 0.5)
    grid = synthetic.ball_density_grid(H=64, cascade=model.cascade, bound=float(model.bound), radius=0.5, sigma=50.0).cuda()
    model.density_grid.copy_(grid)
    model.density_bitfield = raymarching.packbits(model.density_grid, min(grid.clamp(min=0).mean().item(), 10.0), model.density_bitfield)
    o, d = synthetic.sphere_rays(N, seed=5)
    tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(9))
    return model, o.cuda(), d.cuda(), tgt.cuda()


@pytest.mark.parametrize("kw", [dict(), dict(color_activation="sigmoid", density_activation="softplus")],
                         ids=["default", "sigmoid-softplus"])
def test_fused_step_gradients_match_autograd(kw):
    N = 1500
    model, o, d, tgt = _scene(N, **kw)
    ref_model = copy.deepcopy(model)

    # autograd path: persistent sink so both sides accumulate fp16 table gradients the same way
    ref = TrainStep(ref_model, loss_scale=128.0)
    ref_model.train()
    out = ref_model.render(o, d, bg_color=1.0, perturb=False)
    loss_ref = torch.nn.functional.mse_loss(out["image"], tgt, reduction="none").mean(-1).mean()
    (loss_ref * 128.0).backward()
    g_table_ref = ref.table_grad.float().clone()
    g_mlp_ref = [p.grad.float().clone() for p in ref.mlp_params]

    fs = FusedTrainStep(model, N, loss_scale=128.0, perturb=False, use_graph=False)
    fs.set_rays(o, d, tgt)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    assert fs.last_num_points == out["num_points"]
    torch.testing.assert_close(fs.image, out["image"].float(), rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(fs.loss[0], loss_ref.float(), rtol=2e-3, atol=1e-6)

    def close(a, b, name):
        scale = b.abs().max().clamp(min=1e-8)
        err = (a - b).abs() / scale
        assert err.max().item() < 3e-2 and err.mean().item() < 1e-3, (name, err.max().item(), err.mean().item())

    close(fs.table_grad.float(), g_table_ref, "table")
    layers = list(model.grid_mlp.net) + list(model.view_mlp.net)
    for i, (lin, gref) in enumerate(zip(layers, g_mlp_ref)):
        n, k = lin.weight.shape
        close(fs._w_grad_views[i][:n, :k], gref, f"w{i}")
        pad = fs._w_grad_views[i].clone()
        pad[:n, :k] = 0
        assert pad.abs().max().item() == 0.0      # padding rows / columns never receive gradient


def test_fused_step_light_stage_config_matches_autograd():
    """configs[2]: light-direction SH conditioning (view_mlp 47 -> 80 -> 80 -> 3), scene contraction (bound 2, 2 cascades through
    the renderer, renderer.py:171-176), clamped_exp colour, HDR loss -- the widths outside the warp-specialised kernels take the
    density-field + view-MLP kernel pairs inside the same sync-free step."""
    N = 1100
    model, o, d, tgt = _scene(N, bound=2, contract=True, rfield=True, color_activation="clamped_exp")
    assert model.cascade == 2 and model.view_mlp.net[0].weight.shape == (80, 47)
    ld = synthetic.unit_vectors(N, seed=3).cuda()
    exposure = torch.tensor([1.0, 0.25, 1 / 16]).cuda()[torch.arange(N).cuda() % 3]
    ref_model = copy.deepcopy(model)
    ref = TrainStep(ref_model, loss_scale=128.0)
    ref_model.train()
    out = ref_model.render(o, d, rays_ldir=ld, bg_color=1.0, perturb=False)
    pred = out["image"]
    clip = torch.minimum(torch.tensor(1.0, device=pred.device), pred * exposure.unsqueeze(1))
    loss_ref = ((clip - tgt) ** 2 * (1.0 / (1e-3 + clip.detach())) ** 2).sum() / (3 * N)
    (loss_ref * 128.0).backward()
    g_table_ref = ref.table_grad.float().clone()
    g_mlp_ref = [p.grad.float().clone() for p in ref.mlp_params]

    fs = FusedTrainStep(model, N, loss_scale=128.0, perturb=False, use_graph=False, loss="hdr")
    assert not fs.ws
    fs.set_rays(o, d, tgt, rays_ldir=ld, exposure=exposure)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    assert fs.last_num_points == out["num_points"] and fs.last_num_points > 10 * N
    torch.testing.assert_close(fs.image, pred.float(), rtol=2e-3, atol=2e-4)
    torch.testing.assert_close(fs.loss[0], loss_ref.float(), rtol=5e-3, atol=1e-6)

    def close(a, b, name):
        scale = b.abs().max().clamp(min=1e-8)
        err = (a - b).abs() / scale
        assert err.max().item() < 3e-2 and err.mean().item() < 1e-3, (name, err.max().item(), err.mean().item())

    close(fs.table_grad.float(), g_table_ref, "table")
    layers = list(model.grid_mlp.net) + list(model.view_mlp.net)
    for i, (lin, gref) in enumerate(zip(layers, g_mlp_ref)):
        n, k = lin.weight.shape
        close(fs._w_grad_views[i][:n, :k], gref, f"w{i}")

    # and the graph-replayed step trains it
    losses = [fs.step(o, d, tgt, rays_ldir=ld, update_grid=False, exposure=exposure).item() for _ in range(2)]
    g = FusedTrainStep(copy.deepcopy(ref_model), N, perturb=False, use_graph=True, loss="hdr")
    lg = [g.step(o, d, tgt, rays_ldir=ld, update_grid=False, exposure=exposure).item() for _ in range(15)]
    assert lg[-1] < lg[0] and all(l == l for l in lg + losses)


@pytest.mark.parametrize("kw", [dict(pose_opt="barf", start_annealing=0.0, end_annealing=0.5), dict(interpolation="smoothstep")],
                         ids=["barf-window", "smoothstep"])
def test_fused_step_ray_gradients_match_autograd(kw):
    """BARF: rays_o / rays_d require grad.  The autograd path runs op by op (grid_encode with input gradients, SH backward,
    march_rays_train.backward); the fused step produces the same dL/d rays in its backward kernel + one segment-sum kernel."""
    N = 1300
    kw = dict(kw)
    smooth = kw.pop("interpolation", None) == "smoothstep"
    model, o, d, tgt = _scene(N, **kw)
    if smooth:
        model.grid_encoder.interpolation, model.grid_encoder.interp_id = "smoothstep", 1
    if kw.get("pose_opt") == "barf":
        model.update_annealing(0.3)
    d = d * (1.0 + 0.5 * torch.rand(N, 1, device=d.device))       # get_rays does not normalise (train_utils.py:157-160)
    ref_model = copy.deepcopy(model)
    ref = TrainStep(ref_model, loss_scale=128.0)
    ref_model.train()
    ro, rd = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    out = ref_model.render(ro, rd, bg_color=1.0, perturb=False)
    loss_ref = torch.nn.functional.mse_loss(out["image"], tgt, reduction="none").mean(-1).mean()
    (loss_ref * 128.0).backward()
    g_table_ref = ref.table_grad.float().clone()

    fs = FusedTrainStep(model, N, loss_scale=128.0, perturb=False, use_graph=False, ray_grads=True)
    if kw.get("pose_opt") == "barf":
        fs.feat_weights.copy_(model._feat_weights(o.device))
    fs.set_rays(o, d, tgt)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    assert fs.last_num_points == out["num_points"]
    torch.testing.assert_close(fs.loss[0], loss_ref.float(), rtol=2e-3, atol=1e-6)

    def close(a, b, name, mx=3e-2, mean=2e-3):
        scale = b.abs().max().clamp(min=1e-8)
        err = (a - b).abs() / scale
        assert err.max().item() < mx and err.mean().item() < mean, (name, err.max().item(), err.mean().item())

    close(fs.table_grad.float(), g_table_ref, "table")
    assert ro.grad.abs().max().item() > 0 and rd.grad.abs().max().item() > 0
    close(fs.d_rays_o, ro.grad.float(), "rays_o")
    close(fs.d_rays_d, rd.grad.float(), "rays_d")


def _camera_batch(N, C, seed=7):
    g = torch.Generator().manual_seed(seed)
    poses = pose.look_at_poses(C, radius=2.0)
    idx = torch.randint(0, C, (N,), generator=g)
    # pixels of a 64 x 64 image, focal 80: the ball of radius 0.5 at distance 2 fills most of the frame
    ij = torch.randint(0, 64, (N, 2), generator=g).float() + 0.5
    dirs = pose.pixel_directions(ij[:, 0], ij[:, 1], (80.0, 80.0, 32.0, 32.0))
    tgt = torch.rand(N, 3, generator=g)
    return poses.cuda(), idx.cuda(), dirs.cuda(), tgt.cuda()


def test_fused_step_pose_gradients_match_autograd():
    """configs[4] ingredients: BARF window + rays generated from refined poses; d loss / d se3_refine of the captured step
    (pose kernel -> march -> field -> composite -> backward -> segment sum -> pose backward) == autograd through the torch
    formulation of barf/camera.py + get_rays and the op-by-op renderer."""
    N, C = 1400, 10
    model, _, _, _ = _scene(N, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
    model.update_annealing(0.35)
    poses, idx, dirs, tgt = _camera_batch(N, C)
    cam = pose.CameraOptimizer(C, "cuda")
    cam.se3_refine.weight.data.normal_(0, 0.02, generator=torch.Generator(device="cuda").manual_seed(1))

    ref_model, ref_cam = copy.deepcopy(model), copy.deepcopy(cam)
    ref = TrainStep(ref_model, loss_scale=128.0)
    ref_model.train()
    refined = ref_cam(poses[idx], idx)
    rd = (dirs.unsqueeze(1) @ refined[:, :3, :3].transpose(-1, -2)).squeeze(1)
    ro = refined[:, :3, 3]
    out = ref_model.render(ro, rd, bg_color=1.0, perturb=False)
    loss_ref = torch.nn.functional.mse_loss(out["image"], tgt, reduction="none").mean(-1).mean()
    (loss_ref * 128.0).backward()
    g_ref = ref_cam.se3_refine.weight.grad.float()
    assert out["num_points"] > 10 * N and g_ref.abs().max().item() > 0

    fs = FusedTrainStep(model, N, loss_scale=128.0, perturb=False, use_graph=False, pose_optimizer=cam, poses=poses)
    fs.feat_weights.copy_(model._feat_weights("cuda"))
    fs.set_camera_rays(idx, dirs, tgt)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    assert fs.last_num_points == out["num_points"]
    torch.testing.assert_close(fs.rays_o, ro.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(fs.rays_d, rd.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(fs.loss[0], loss_ref.float(), rtol=2e-3, atol=1e-6)
    scale = g_ref.abs().max()
    err = (fs.se3_grad - g_ref).abs() / scale
    assert err.max().item() < 3e-2 and err.mean().item() < 5e-3, (err.max().item(), err.mean().item())


def test_fused_step_pose_refinement_trains_under_graph():
    """Graph-replayed steps with the pose optimizer: se3_refine moves, the loss stays finite and decreases, a learning rate of
    zero (set_lr, read from device memory by the captured Adam) freezes the poses."""
    N, C = 1024, 8
    model, _, _, _ = _scene(N, pose_opt="barf", start_annealing=0.0, end_annealing=0.5)
    model.update_annealing(1.0)
    poses, idx, dirs, tgt = _camera_batch(N, C, seed=3)
    cam = pose.CameraOptimizer(C, "cuda")
    fs = FusedTrainStep(model, N, perturb=False, use_graph=True, pose_optimizer=cam, poses=poses, pose_lr=1e-3)
    losses = [fs.step(cam_idx=idx, dirs_cam=dirs, target_rgb=tgt, update_grid=False).item() for _ in range(12)]
    fs.flush()
    torch.cuda.synchronize()
    assert all(l == l for l in losses) and losses[-1] < losses[0]
    moved = cam.se3_refine.weight.data.abs().max().item()
    assert 1e-4 < moved < 0.1                    # ~ lr * steps
    assert cam.se3_refine.weight.data_ptr() == fs.se3.data_ptr()
    fs.set_lr(pose_lr=0.0)
    before = fs.se3.clone()
    for _ in range(3):
        fs.step(cam_idx=idx, dirs_cam=dirs, target_rgb=tgt, update_grid=False)
    fs.flush()
    torch.cuda.synchronize()
    assert torch.equal(fs.se3, before)


def test_fused_step_graph_trains_and_matches_eager():
    N = 1024
    model_a, o, d, tgt = _scene(N)
    model_b = copy.deepcopy(model_a)
    a = FusedTrainStep(model_a, N, perturb=False, use_graph=True)
    b = FusedTrainStep(model_b, N, perturb=False, use_graph=False)
    la, lb = [], []
    for _ in range(20):
        la.append(a.step(o, d, tgt, update_grid=False).item())
        lb.append(b.step(o, d, tgt, update_grid=False).item())
    assert la[-1] < 0.85 * la[0]                  # it optimises
    # graph replay == eager launches (differences: atomic order only)
    assert abs(la[0] - lb[0]) < 1e-6
    assert abs(la[-1] - lb[-1]) < 0.15 * max(la[-1], 1e-6)
    # module parameters are views of the trained master weights
    w = model_a.view_mlp.net[2].weight
    assert w.shape == (3, 64) and torch.equal(w, a._w_master_views[5][:3, :64])
    # trained table is visible to the encoder
    assert torch.equal(model_a.grid_encoder.embeddings.data, a.table_master.half())


def test_fused_step_capacity_overflow_drops_whole_rays():
    N = 512
    model, o, d, tgt = _scene(N)
    full = FusedTrainStep(copy.deepcopy(model), N, perturb=False, use_graph=False)
    full.set_rays(o, d, tgt)
    full._launch_forward_backward()
    M = full.last_num_points
    cap = M // 2
    part = FusedTrainStep(model, N, perturb=False, use_graph=False, max_samples=cap)
    part.set_rays(o, d, tgt)
    part._launch_forward_backward()
    torch.cuda.synchronize()
    fit = int(part.counter[2].item())
    rays = part.rays.cpu()
    ends = rays[:, 0] + rays[:, 1]
    assert fit <= cap and fit == int(ends[ends <= cap].max())
    # rays that fit render identically, the others see only the background
    ok = (ends <= cap) & (rays[:, 1] > 0)
    torch.testing.assert_close(part.image[ok.cuda()], full.image[ok.cuda()], rtol=1e-5, atol=1e-6)
    dropped = (ends > cap).cuda()
    assert torch.all(part.image[dropped] == 1.0)
    assert torch.isfinite(part.table_grad.float()).all()


def test_fused_step_hdr_loss_matches_torch():
    """loss_mode 1 = the clipped, tone-curve weighted MSE of the raw/HDR path (nerf/train_utils.py:529-536)."""
    N = 1200
    model, o, d, tgt = _scene(N, color_activation="exp")
    ref_model = copy.deepcopy(model)
    exposure = (torch.rand(N, generator=torch.Generator().manual_seed(3)) * 3 + 0.25).cuda()
    ref = TrainStep(ref_model, loss_scale=128.0)
    ref_model.train()
    out = ref_model.render(o, d, bg_color=1.0, perturb=False)
    pred = out["image"]
    clip = torch.minimum(torch.tensor(1.0, device=pred.device), pred * exposure.unsqueeze(1))
    scaling = 1.0 / (1e-3 + clip.detach())
    loss_ref = ((clip - tgt) ** 2 * scaling ** 2).sum() / (3 * N)
    (loss_ref * 128.0).backward()
    g_table_ref = ref.table_grad.float().clone()

    fs = FusedTrainStep(model, N, loss_scale=128.0, perturb=False, use_graph=False, loss="hdr")
    fs.set_rays(o, d, tgt, exposure=exposure)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    torch.testing.assert_close(fs.loss[0], loss_ref.float(), rtol=5e-3, atol=1e-6)
    scale = g_table_ref.abs().max().clamp(min=1e-8)
    err = (fs.table_grad.float() - g_table_ref).abs() / scale
    assert err.max().item() < 3e-2 and err.mean().item() < 1e-3, (err.max().item(), err.mean().item())


def test_padded_weight_cache_sees_native_updates():
    """field._pad_weight caches the padded fp16 weights per tensor version; the native optimizers write parameters through raw
    pointers (also from replayed graphs), so the cache key carries _lib.weights_epoch."""
    N = 1024
    model, o, d, tgt = _scene(N)
    fs = FusedTrainStep(model, N, perturb=False, use_graph=True)

    def render():
        model.eval()
        with torch.no_grad():
            return model.render(o, d, bg_color=1.0, perturb=False)["image"].clone()

    img0 = render()                       # fills the cache
    for _ in range(4):
        fs.step(o, d, tgt, update_grid=False)
    fs.flush()
    img1 = render()
    for lin in list(model.grid_mlp.net) + list(model.view_mlp.net):
        lin.weight._ngp_padded = None         # drop the cached copies
    img2 = render()
    assert torch.equal(img1, img2)
    assert not torch.equal(img0, img1)


def test_composite_hdr_loss_kernel_matches_reference_lines():
    """The fused composite + loss kernel in HDR mode against tests/golden/hdr_loss.npz (the reference's train_utils.py:512-541
    executed on seeded inputs): every ray has one opaque sample, so the composited colour IS that sample's colour and the
    kernel's loss / d loss / d rgb must be the fixture's."""
    import os
    import numpy as np
    from raw_ngp_b200 import _lib
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hdr_loss.npz"))
    dev = "cuda"
    pred = torch.from_numpy(z["pred_rgb"]).to(dev).contiguous()
    gt = torch.from_numpy(z["gt_rgb"]).to(dev).contiguous()
    exposure = torch.from_numpy(z["exposure"]).to(dev).contiguous()
    N = pred.shape[0]
    sigma = torch.full((N,), 1e6, device=dev)                     # alpha = 1 - exp(-sigma dt) == 1: the sample is opaque
    ts = torch.stack([torch.full((N,), 1.0, device=dev), torch.full((N,), 0.01, device=dev)], dim=-1).contiguous()
    rays = torch.stack([torch.arange(N, device=dev), torch.ones(N, device=dev)], dim=-1).int().contiguous()
    image, ray_loss, loss = torch.zeros(N, 3, device=dev), torch.zeros(N, device=dev), torch.zeros(1, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    d_sigma, d_rgb = torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
    scale = 128.0
    P = _lib.ptr
    _lib.call("ngp_composite_train_mse", P(sigma), P(pred), P(ts), P(rays), N, None, N, 1e-4, 1.0, P(gt), scale, P(image), P(ray_loss),
              P(loss), P(ticket), P(d_sigma), P(d_rgb), 1, P(exposure), _lib.stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(image, pred, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(loss.item(), float(z["loss"]), rtol=1e-5)
    torch.testing.assert_close(d_rgb / scale, torch.from_numpy(z["d_pred"]).to(dev), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", ["hdr_random_rgba_bayer_gaussian_entropy", "hdr_white_planck", "hdr_black_hanning_bayer", "mse_random_rgba_entropy"])
def test_composite_train_loss_kernel_matches_reference_train_step_lines(case):
    """ngp_composite_train_loss (per-ray background, RGBA ground truth, lossmult / loss_weight, entropy regulariser) against
    tests/golden/train_step_loss.npz = the reference's own train_step lines (nerf/train_utils.py:494-557) executed on seeded inputs
    by tools/make_golden_trainstep.py.  Every ray has ONE sample with alpha = the fixture's weights_sum and the fixture's colour, so
    image, loss and the gradients with respect to the composited colour / opacity map one to one onto the kernel's outputs."""
    import ctypes
    import os
    import numpy as np
    from raw_ngp_b200 import _lib
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_step_loss.npz"))
    f = lambda k: torch.from_numpy(np.ascontiguousarray(z[f"{case}/{k}"])).cuda()
    colour, ws, images, bg, lossmult, loss_weight = f("colour"), f("ws"), f("images"), f("bg"), f("lossmult"), f("loss_weight")
    exposure, d_comp, d_ws = f("exposure"), f("d_comp"), f("d_ws")
    hdr, lam = int(z[f"{case}/hdr"]), float(z[f"{case}/lambda_entropy"])
    N, dev, dt = ws.shape[0], "cuda", 0.01
    sigma = (-torch.log1p(-ws.double().clamp(max=1 - 1e-12)) / dt).float()
    sigma[ws >= 1.0] = 1e8
    ts = torch.stack([torch.ones(N, device=dev), torch.full((N,), dt, device=dev)], dim=-1).contiguous()
    rays = torch.stack([torch.arange(N, device=dev), torch.ones(N, device=dev)], dim=-1).int().contiguous()
    image, ray_loss, loss = torch.zeros(N, 3, device=dev), torch.zeros(N, device=dev), torch.zeros(1, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    d_sigma, d_rgb = torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
    ent, wso, parts = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(2, device=dev)
    inv_norm = lossmult.sum().reciprocal().reshape(1)
    target = images[:, :3].contiguous()
    alpha_t = images[:, 3].contiguous() if images.shape[1] == 4 else None
    P = _lib.ptr
    opts = _lib.LossOpts(P(bg), P(alpha_t), P(lossmult), P(loss_weight), P(inv_norm), None, lam, P(ent), P(wso), None, P(parts))
    scale = 64.0
    _lib.call("ngp_composite_train_loss", P(sigma), P(colour), P(ts), P(rays), N, None, N, 1e-8, 0.0, P(target), scale, P(image), P(ray_loss),
              P(loss), P(ticket), P(d_sigma), P(d_rgb), hdr, P(exposure), ctypes.byref(opts), _lib.stream())
    torch.cuda.synchronize()
    alpha = 1 - torch.exp(-sigma * dt)
    torch.testing.assert_close(wso, ws, rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(loss.item(), float(z[f"{case}/loss"]), rtol=2e-5)
    np.testing.assert_allclose(parts.sum().item(), loss.item(), rtol=1e-6)
    # d loss / d (composited colour) = grad_rgbs / w ;  d loss / d sigma = dt (1 - alpha) (sum_c dL/dcomp_c colour_c + dL/dws)
    g_scale = d_comp.abs().max()
    torch.testing.assert_close(d_rgb / scale, d_comp * alpha[:, None], rtol=1e-4, atol=1e-6 * g_scale.item())
    exp_sigma = dt * (1 - alpha) * ((d_comp * colour).sum(-1) + d_ws)
    torch.testing.assert_close(d_sigma / scale, exp_sigma, rtol=2e-4, atol=1e-6 * exp_sigma.abs().max().item())


def test_fused_step_extras_cam_near_far_adaptive_rays_and_random_bg():
    """cam_near_far clamps the marched interval like renderer.py:529-533; the adaptive ray count follows
    int(round(num_points / samples * num_rays)) (train_utils.py:563-564) on the device; bg_color='random' draws a per-ray background;
    rays beyond the live count get no samples, no loss, no gradient."""
    N = 1024
    model, o, d, tgt = _scene(N)
    fs = FusedTrainStep(model, N, perturb=False, use_graph=False, cam_near_far=True, adaptive_num_rays=True, num_points=20000,
                        bg_color="random", lambda_entropy=1e-3)
    cnf = torch.tensor([1.7, 2.2]).repeat(N, 1).cuda()
    fs.set_rays(o, d, tgt, cam_near_far=cnf)
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    M0 = fs.last_num_points
    t = fs.ts[:M0, 0]
    assert M0 > 0 and t.min().item() >= 1.7 and (t - fs.ts[:M0, 1]).max().item() < 2.2
    ref = model
    ref.train()
    out = ref.render(o, d, bg_color=0, perturb=False, cam_near_far=cnf)
    assert out["num_points"] == M0
    torch.testing.assert_close(fs.weights_sum, out["weights_sum"].float(), rtol=2e-3, atol=2e-4)
    assert fs.bg_rays.min().item() >= 0 and 0.3 < fs.bg_rays.mean().item() < 0.7
    n1 = int(fs.n_rays_dev.item())
    assert n1 == max(1, min(N, int(round((20000 / M0) * N))))
    fs.table_grad.zero_()
    fs._launch_forward_backward()
    torch.cuda.synchronize()
    counts = fs.rays[:, 1]
    assert (counts[n1:] == 0).all() and counts[:n1].sum().item() == fs.last_num_points
    assert (fs.image[n1:] == 0).all() and torch.isfinite(fs.loss).all()
    # graph replay trains with all the options on
    g = FusedTrainStep(_scene(N)[0], N, use_graph=True, cam_near_far=True, adaptive_num_rays=True, num_points=20000, bg_color="random",
                       lambda_entropy=1e-3, rgba_targets=True, lossmult=True, loss_weight=True, loss="hdr")
    rgba = torch.cat([tgt, torch.rand(N, 1, device="cuda")], dim=-1)
    ls = [g.step(o, d, rgba, update_grid=False, cam_near_far=cnf, lossmult=torch.ones(N, 3, device="cuda"), loss_weight=1.0).item() for _ in range(8)]
    g.flush()
    assert all(l == l for l in ls) and 1 <= int(g.n_rays_dev.item()) <= N


def test_fused_step_checkpoint_round_trip_and_reference_pth(tmp_path):
    """Resume: state_dict() / load_state_dict() carry the fp32 masters, Adam moments, device step counters and the loss scale, so a
    trainer rebuilt from a checkpoint continues like the original.  model_state_dict() is a checkpoint in the REFERENCE's layout
    (fp32 hash table): saved with torch.save, it loads with strict=True into the reference's own NeRFNetwork (oracle/ref_stack.py)
    and renders the same image there; loading it back through load_model_state_dict() restores the fp32 master exactly."""
    import _refstep as R
    N = 1024
    model, o, d, tgt = _scene(N)
    a = FusedTrainStep(model, N, perturb=False, use_graph=True, growth_interval=3)
    for _ in range(5):
        a.step(o, d, tgt, update_grid=False)
    ckpt, msd = a.state_dict(), a.model_state_dict()
    assert msd["grid_encoder.embeddings"].dtype == torch.float32 and float(ckpt["loss_scale"]) == 256.0      # grew once (3 clean steps)
    path = tmp_path / "ngp.pth"
    torch.save({"model": msd, "trainer": ckpt}, path)
    blob = torch.load(path, map_location="cuda")

    model_b, _, _, _ = _scene(N)
    model_b.grid_encoder.embeddings.data.zero_()
    b = FusedTrainStep(model_b, N, perturb=False, use_graph=True, growth_interval=3)
    b.load_model_state_dict(blob["model"])
    b.load_state_dict(blob["trainer"])
    assert torch.equal(b.table_master, a.table_master) and torch.equal(b.w_master, a.w_master)
    assert torch.equal(model_b.grid_encoder.embeddings.data, a.table_master.half()) and b.opt.step_count == 5
    la = [a.step(o, d, tgt, update_grid=False).item() for _ in range(4)]
    lb = [b.step(o, d, tgt, update_grid=False).item() for _ in range(4)]
    a.flush(); b.flush()
    assert abs(la[0] - lb[0]) < 1e-6 * max(1.0, abs(la[0])) and abs(la[-1] - lb[-1]) < 2e-3 * abs(la[-1])
    assert float(b.scale_dev) == float(a.scale_dev) and int(b.opt_step_dev) == int(a.opt_step_dev) == 9

    # a forced overflow: the step is skipped, the scale halves, the skipped-step counter is visible without being polled per step
    before = b.table_master.clone()
    b.step(o, d, tgt * float("inf"), update_grid=False)
    b.flush()
    torch.cuda.synchronize()
    assert torch.equal(b.table_master, before) and b.skipped_steps() == 1 and float(b.scale_dev) == 0.5 * float(a.scale_dev)

    # the reference's own network loads the checkpoint (strict) and renders the same image
    stack = R.stacks().get("ref")
    kw = {k: getattr(model.opt, k) for k in R.OPT_KEYS if hasattr(model.opt, k)}
    ref = stack.build_network(stack.make_opt(**kw)).cuda()
    res = ref.load_state_dict(torch.load(path, map_location="cuda")["model"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys and ref.grid_encoder.embeddings.dtype == torch.float32
    model_c, _, _, _ = _scene(N)                 # this repository's network from the same file (the trainers above moved on)
    model_c.load_state_dict(torch.load(path, map_location="cuda")["model"], strict=True)
    ref.eval(); model_c.eval()
    oo, dd = synthetic.sphere_rays(2000, seed=8)
    with torch.no_grad():
        img_ref = ref.render(oo.cuda(), dd.cuda(), bg_color=1, perturb=False)["image"]
        img = model_c.render(oo.cuda(), dd.cuda(), bg_color=1, perturb=False)["image"]
    torch.testing.assert_close(img.float(), img_ref.float(), rtol=2e-3, atol=2e-3)


def test_library_uniform_generator():
    """ngp_uniform (the fused step's jitter / background stream, drawn inside the captured graph): values in [0, 1), the moments of a
    uniform distribution, a new stream per call (device-side counter), the same stream for the same (seed, counter), and the two launch
    shapes (one block with the counter advance inside; many blocks + a separate advance) agreeing on the common prefix."""
    from raw_ngp_b200 import _lib
    dev = "cuda"
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    a, b = torch.empty(65536, device=dev), torch.empty(65536, device=dev)
    _lib.call("ngp_uniform", _lib.ptr(a), a.numel(), 1234, _lib.ptr(counter), _lib.stream())
    _lib.call("ngp_uniform", _lib.ptr(b), b.numel(), 1234, _lib.ptr(counter), _lib.stream())
    assert counter.item() == 2
    for t in (a, b):
        assert t.min().item() >= 0.0 and t.max().item() < 1.0
        assert abs(t.mean().item() - 0.5) < 5e-3 and abs(t.var().item() - 1 / 12) < 2e-3
    assert (a == b).float().mean().item() < 1e-3                      # a different stream every call
    corr = ((a - 0.5) * (b - 0.5)).mean().item() * 12
    lag = ((a[1:] - 0.5) * (a[:-1] - 0.5)).mean().item() * 12
    assert abs(corr) < 2e-2 and abs(lag) < 2e-2
    counter.zero_()
    c = torch.empty(65536, device=dev)
    _lib.call("ngp_uniform", _lib.ptr(c), c.numel(), 1234, _lib.ptr(counter), _lib.stream())
    assert torch.equal(a, c)                                          # same (seed, counter) -> same stream
    counter.zero_()
    big = torch.empty(3 * 65536 + 5, device=dev)                      # multi-block launch shape
    _lib.call("ngp_uniform", _lib.ptr(big), big.numel(), 1234, _lib.ptr(counter), _lib.stream())
    assert counter.item() == 1 and torch.equal(big[:65536], a) and big.max().item() < 1.0
    d = torch.empty(65536, device=dev)
    counter.zero_()
    _lib.call("ngp_uniform", _lib.ptr(d), d.numel(), 99, _lib.ptr(counter), _lib.stream())
    assert (a == d).float().mean().item() < 1e-3                      # another seed, another stream
