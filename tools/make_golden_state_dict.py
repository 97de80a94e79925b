#!/usr/bin/env python3
"""Generates tests/golden/state_dict.json: names, shapes and dtypes of the state_dict of the REFERENCE's NeRFNetwork
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
        out[tag] = dict(opt=kw, entries={k: [list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()})
    with open(os.path.join(out_dir, "state_dict.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v["entries"]) for k, v in out.items()})
    print(json.dumps(out["default"]["entries"], indent=0)[:1500])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden"))
