#!/usr/bin/env python3
"""Generates tests/golden/pose.npz by importing the REFERENCE's Python pose code (barf/camera.py, barf/camera_optimizers.py,
nerf/train_utils.get_rays) in the build container (CPU, fp32, autocast off) on seeded inputs:

    python tools/make_golden_pose.py tests/golden

The reference is imported from /root/reference with empty stub modules for its third-party imports that are not
installed here (easydict, the trainer's logging / metric packages); none of them is touched by the functions called.
tests/test_pose.py checks raw_ngp_b200/pose.py against the fixture without a GPU and without the reference.
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    _stub("easydict", EasyDict=dict)
    sys.path.insert(0, REF)
    camera = importlib.import_module("barf.camera")
    # get_rays lives in nerf/train_utils.py whose import chain needs a dozen absent packages: execute only that function
    src = open(os.path.join(REF, "nerf", "train_utils.py")).read()
    start = src.index("@torch.cuda.amp.autocast(enabled=False)\ndef get_rays")
    end = src.index("def visualize_rays")
    ns = {"torch": torch, "np": np}
    exec("def custom_meshgrid(*args):\n    return torch.meshgrid(*args, indexing='ij')\n", ns)
    exec(compile(src[start:end], "train_utils.get_rays", "exec"), ns)
    return camera, ns["get_rays"]


def main(out_dir):
    camera, get_rays = load_reference()
    g = torch.Generator().manual_seed(11)
    C, N, H, W = 12, 64, 48, 64
    se3 = torch.randn(C, 6, generator=g) * torch.tensor([0.3, 0.3, 0.3, 0.2, 0.2, 0.2])
    se3[0] = 0.0                       # the initial state of CameraOptimizer
    se3[1, :3] = 0.0                   # pure translation
    se3[2, :3] *= 8.0                  # large rotation (theta ~ 3)
    se3 = se3.requires_grad_(True)
    SE3 = camera.lie.se3_to_SE3(se3)
    # dataset poses: random rotations (from another se3) + translations
    poses = torch.eye(4).repeat(C, 1, 1)
    poses[:, :3, :] = camera.lie.se3_to_SE3(torch.randn(C, 6, generator=g)).detach()
    idx = torch.randint(0, C, (N,), generator=g)
    refined = camera.pose.compose([camera.lie.se3_to_SE3(se3[idx]), poses[idx][:, :3, :]])
    intr = np.array([55.0, 57.0, 31.5, 23.5], dtype=np.float32)
    coords = torch.stack([torch.randint(0, H, (N,), generator=g), torch.randint(0, W, (N,), generator=g)], dim=-1)
    rays = get_rays(refined, intr, H, W, N, coords=coords)
    # a fixed cotangent: gradients with respect to se3 through autograd of the reference code
    go, gd = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g)
    (rays["rays_o"] * go).sum().add((rays["rays_d"] * gd).sum()).backward()
    np.savez_compressed(os.path.join(out_dir, "pose.npz"), se3=se3.detach().numpy(), SE3=SE3.detach().numpy(), poses=poses.numpy(),
                        idx=idx.numpy(), refined=refined.detach().numpy(), intrinsics=intr, H=H, W=W, coords=coords.numpy(),
                        rays_o=rays["rays_o"].detach().numpy(), rays_d=rays["rays_d"].detach().numpy(), i=rays["i"].numpy(),
                        j=rays["j"].numpy(), g_rays_o=go.numpy(), g_rays_d=gd.numpy(), d_se3=se3.grad.numpy())
    print("wrote", os.path.join(out_dir, "pose.npz"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden")
