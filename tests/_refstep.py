"""Shared helpers of the -m gpu parity tests that run the REFERENCE's own nerf/network.py + nerf/renderer.py (oracle/ref_stack.py)
beside this repository's path on identical weights, rays and noises."""
import json
import os

import torch

from raw_ngp_b200 import raymarching, synthetic
from raw_ngp_b200.nerf import NeRFNetwork, default_opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPT_KEYS = ("bound", "contract", "grid_size", "min_near", "density_thresh", "dt_gamma", "max_steps", "T_thresh", "fp16", "hashmap_size",
            "hashgrid_resolution", "rfield", "pose_opt", "internal_activation", "beta", "density_activation", "color_activation",
            "start_annealing", "end_annealing", "num_cameras")


def stacks():
    import pytest
    from oracle import ref_cuda, ref_stack
    if not (ref_cuda.available() and ref_stack.available()):
        pytest.skip("oracle/_ref (reference kernels + staged reference python) not built: run oracle/build_ref.sh where /root/reference exists")
    return ref_stack


def build_scene(N, seed=5, table_scale=0.5, ball_radius=0.5, **cfg):
    """Our NeRFNetwork with a ball-shaped occupancy grid, random table / MLP weights, N sphere rays and random targets."""
    torch.manual_seed(0)
    opt = default_opt(**cfg)
    model = NeRFNetwork(opt).cuda()
    model.grid_encoder.embeddings.data.uniform_(-table_scale, table_scale)
    H = opt.grid_size
    grid = synthetic.ball_density_grid(H=H, cascade=model.cascade, bound=float(model.bound), radius=ball_radius, sigma=50.0).cuda()
    model.density_grid.copy_(grid)
    model.density_bitfield = raymarching.packbits(model.density_grid, min(grid.clamp(min=0).mean().item(), 10.0), model.density_bitfield)
    o, d = synthetic.sphere_rays(N, seed=seed)
    tgt = torch.rand(N, 3, generator=torch.Generator().manual_seed(9))
    return model, o.cuda(), d.cuda(), tgt.cuda()


def reference_model(stack, model, table_dtype=torch.float16, fp16=True):
    """The reference's NeRFNetwork (unmodified source) holding the same parameters and occupancy state as `model`.  The hash table
    is rounded to fp16 on both sides (BASELINE configs[1]: "fp16 hash grid"); table_dtype=float32 keeps those rounded values in an
    fp32 tensor (the fork's default dtype, grid.py:43-46) for the fp32 'truth' runs."""
    kw = {k: getattr(model.opt, k) for k in OPT_KEYS if hasattr(model.opt, k)}
    kw["fp16"] = fp16
    ropt = stack.make_opt(**kw)
    ref = stack.build_network(ropt).cuda()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sd["grid_encoder.embeddings"] = sd["grid_encoder.embeddings"].float()
    missing = ref.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    t16 = ref.grid_encoder.embeddings.data.half()
    ref.grid_encoder.embeddings.data = t16 if table_dtype == torch.float16 else t16.float()
    ref.mean_density, ref.iter_density = float(model.mean_density), int(model.iter_density)
    ref.annealing = model.annealing
    return ref


def hdr_loss(pred_rgb, gt_rgb, exposure):
    """nerf/train_utils.py:529-536 with lossmult = loss_weight = 1 (their defaults)."""
    rgb_render_clip = torch.minimum(torch.tensor(1.0, device=pred_rgb.device), pred_rgb * exposure.unsqueeze(1))
    resid_sq_clip = (rgb_render_clip - gt_rgb) ** 2
    scaling_grad = 1.0 / (1e-3 + rgb_render_clip.detach())
    data_loss = resid_sq_clip * scaling_grad ** 2
    lossmult = torch.ones_like(gt_rgb)
    return (data_loss * lossmult).sum() / lossmult.sum()


def mse_loss(pred_rgb, gt_rgb):
    """nerf/train_utils.py:538-541 with criterion = MSELoss(reduction='none') (main.py:236)."""
    return torch.nn.functional.mse_loss(pred_rgb, gt_rgb, reduction="none").mean(-1).mean()


def err_stats(a, b):
    """max and mean of |a - b| normalised by max |b|."""
    a, b = a.float(), b.float()
    scale = b.abs().max().clamp(min=1e-12)
    e = (a - b).abs() / scale
    return e.max().item(), e.mean().item()


def record(name, metrics):
    """Appends the measured parity figures to gpurun_out/parity_metrics.jsonl (copied into profiles/ for DESIGN.md)."""
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_metrics.jsonl"), "a") as fh:
            fh.write(json.dumps({"test": name, **metrics}) + "\n")
    except OSError:
        pass
