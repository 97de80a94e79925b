#!/usr/bin/env python3
"""Generates tests/golden/near_far_py.npz (the renderer's torch near_far_from_aabb, values and gradients) and
tests/golden/state_dict.json: names, shapes and dtypes of the state_dict of the REFERENCE's NeRFNetwork
(nerf/network.py) for three option sets, by importing the reference's Python modules from /root/reference in the build
container (CPU; the compiled extensions of oracle/_ref are put on the path under the names the reference imports, absent
third-party packages are stubbed -- none is touched by the constructors).  tests/test_abi_and_host.py checks that
raw_ngp_b200.nerf.NeRFNetwork produces the same entries, i.e. that reference checkpoints ('model' of train_utils.py:1141-1180)
load with strict=True.

    python tools/make_golden_state_dict.py tests/golden
"""
import glob
import importlib.util
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    return m


def main(out_dir):
    # compiled reference extensions under the names grid.py / raymarching.py / sphere_harmonics.py try first
    for mod in ("_gridencoder", "_raymarching_mob", "_shencoder", "_freqencoder"):
        hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", mod + ".*.so"))
        spec = importlib.util.spec_from_file_location(mod, hits[0])
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        sys.modules[mod] = m
    stub("torch_scatter", segment_csr=lambda src, indptr: torch.segment_reduce(src, "sum", offsets=indptr, axis=0))
    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return self
        def __getattr__(self, k): return _Any()
    for name in ("mcubes", "trimesh", "tensorboardX", "torch_efficient_distloss", "pymeshlab", "imageio", "lpips", "torch_ema",
                 "torchmetrics", "torchmetrics.functional", "rawpy", "easydict", "matplotlib", "matplotlib.pyplot", "packaging"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                stub(name, EasyDict=dict, eff_distloss=_Any(), SummaryWriter=_Any, ExponentialMovingAverage=_Any,
                     structural_similarity_index_measure=_Any(), LPIPS=_Any)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "barf"))          # barf/pose_analysis.py does `import camera`
    from nerf.network import NeRFNetwork
    from types import SimpleNamespace
    base = dict(bound=2, contract=False, grid_size=128, min_near=0.05, density_thresh=10, cuda_ray=True, dt_gamma=0, max_steps=1024,
                T_thresh=1e-8, fp16=True, hashmap_size=19, hashgrid_resolution=2048, rfield=False, pose_opt="none",
                internal_activation="relu", beta=1.0, density_activation="clamped_exp", color_activation="clamped_exp",
                start_annealing=0.0, end_annealing=0.5, lambda_orientation=0, compute_normals=False, device="cpu", num_cameras=0,
                update_extra_interval=16, noise=0.0, identity=False, c_lr=1e-3, iters=1000, scale=1.0)
    out = {}
    for tag, kw in (("default", {}), ("lightstage", dict(rfield=True, contract=True, bound=8)), ("bound1", dict(bound=1)),
                    ("barf", dict(bound=1, pose_opt="barf", num_cameras=7))):
        opt = SimpleNamespace(**dict(base, **kw))
        model = NeRFNetwork(opt)
        enc = model.grid_encoder
        out[tag] = dict(opt=kw, entries={k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()},
                        grid=dict(offsets=[int(v) for v in enc.offsets.tolist()], per_level_scale=float(enc.per_level_scale),
                                  base_resolution=int(enc.base_resolution), n_params=int(enc.n_params), output_dim=int(enc.output_dim)))
    # the differentiable slab test the renderer actually uses (nerf/renderer.py:139-158), outputs and gradients
    import numpy as np
    from nerf.renderer import near_far_from_aabb
    g = torch.Generator().manual_seed(21)
    o = torch.randn(257, 3, generator=g) * 1.5
    d = torch.randn(257, 3, generator=g)
    d[:8] = torch.tensor([0.0, 0.0, 1.0])                   # axis-parallel rays (division by d + 1e-15)
    o[8:16] = 0.0                                            # origins inside the box
    o, d = o.requires_grad_(True), d.requires_grad_(True)
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    near, far = near_far_from_aabb(o, d, aabb, 0.05)
    gn, gf = torch.randn(near.shape, generator=g), torch.randn(far.shape, generator=g)
    ((near * gn).sum() + (far * gf).sum()).backward()
    np.savez_compressed(os.path.join(out_dir, "near_far_py.npz"), rays_o=o.detach().numpy(), rays_d=d.detach().numpy(), aabb=aabb.numpy(),
                        min_near=0.05, nears=near.detach().numpy(), fars=far.detach().numpy(), g_near=gn.numpy(), g_far=gf.numpy(),
                        d_rays_o=o.grad.numpy(), d_rays_d=d.grad.numpy())
    # BARF / BAA-NGP annealing windows (network.py:77-109): run the reference's common_forward with a fixed ramp as the
    # encoder output and an identity grid_mlp, so that its outputs expose the per-feature weighting / blending
    class _Identity(torch.nn.Module):
        dim_out = 16
        def forward(self, x, **kw):
            return x
    class _Ones(torch.nn.Module):
        def forward(self, x, bound=1):
            return (torch.arange(1, 33, dtype=torch.float32) / 8).repeat(x.shape[0], 1)
    windows = {}
    for mode in ("barf", "baangp"):
        opt = SimpleNamespace(**dict(base, bound=1, pose_opt=mode, num_cameras=3, start_annealing=0.1, end_annealing=0.6))
        model = NeRFNetwork(opt)
        model.grid_encoder = _Ones()
        model.grid_mlp = _Identity()
        for a in (0.0, 0.1, 0.17, 0.3, 0.45, 0.6, 0.9):
            model.annealing = a
            sigma, feat = model.common_forward(torch.zeros(2, 3))
            windows[f"{mode}_{a}"] = torch.cat([torch.log(sigma[:1]), feat[0]]).numpy()      # = weighted features 0..31
    np.savez_compressed(os.path.join(out_dir, "annealing_py.npz"), **windows)
    with open(os.path.join(out_dir, "state_dict.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v["entries"]) for k, v in out.items()})
    print(json.dumps(out["default"]["entries"], indent=0)[:1500])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
