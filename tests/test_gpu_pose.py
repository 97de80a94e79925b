"""GPU tests: csrc/pose.cu (refined pose + ray generation, forward and backward) against the golden vectors of the reference's
Python code and against the torch formulation on a larger random case."""
import os

import numpy as np
import pytest
import torch

from raw_ngp_b200 import pose

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose.npz")


def test_pose_rays_kernels_match_reference_golden():
    z = np.load(GOLD)
    t = lambda k: torch.from_numpy(z[k]).cuda()
    se3 = t("se3").requires_grad_(True)
    dirs = pose.pixel_directions(t("i").float() + 0.5, t("j").float() + 0.5, tuple(z["intrinsics"].tolist()))
    o, d = pose.pose_rays(se3, t("poses"), t("idx").int(), dirs)
    torch.testing.assert_close(o, t("rays_o"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d, t("rays_d"), rtol=1e-5, atol=1e-6)
    ((o * t("g_rays_o")).sum() + (d * t("g_rays_d")).sum()).backward()
    torch.testing.assert_close(se3.grad, t("d_se3"), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("zero_init", [True, False])
def test_pose_rays_kernels_match_torch(zero_init):
    C, N = 100, 8192
    g = torch.Generator().manual_seed(1)
    se3 = (torch.zeros(C, 6) if zero_init else torch.randn(C, 6, generator=g) * 0.2).cuda()
    poses = pose.look_at_poses(C).cuda()
    idx = torch.randint(0, C, (N,), generator=g).cuda()
    dirs = torch.cat([torch.randn(N, 2, generator=g) * 0.4, -torch.ones(N, 1)], dim=-1).cuda()
    go, gd = torch.randn(N, 3, generator=g).cuda(), torch.randn(N, 3, generator=g).cuda()

    a = se3.clone().requires_grad_(True)
    refined = pose.compose_pair(pose.se3_to_SE3(a[idx]), poses[idx][:, :3, :])
    rd = (dirs.unsqueeze(1) @ refined[:, :3, :3].transpose(-1, -2)).squeeze(1)
    ro = refined[:, :3, 3]
    ((ro * go).sum() + (rd * gd).sum()).backward()

    b = se3.clone().requires_grad_(True)
    o, d = pose.pose_rays(b, poses, idx.int(), dirs)
    torch.testing.assert_close(o, ro, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(d, rd, rtol=1e-5, atol=1e-5)
    ((o * go).sum() + (d * gd).sum()).backward()
    scale = a.grad.abs().max()
    assert ((b.grad - a.grad).abs().max() / scale).item() < 1e-4


def test_pose_rays_edge_cases():
    """no rays; rays whose gradient is exactly zero contribute nothing; identity refinement passes the dataset pose through"""
    C = 5
    poses = pose.look_at_poses(C).cuda()
    se3 = torch.zeros(C, 6, device="cuda", requires_grad=True)
    o, d = pose.pose_rays(se3, poses, torch.zeros(0, dtype=torch.int32, device="cuda"), torch.zeros(0, 3, device="cuda"))
    assert o.shape == (0, 3) and d.shape == (0, 3)
    idx = torch.tensor([0, 4, 4, 2], dtype=torch.int32, device="cuda")
    dirs = torch.tensor([[0.0, 0.0, -1.0], [0.1, 0.2, -1.0], [0.0, 0.0, -1.0], [-0.3, 0.0, -1.0]], device="cuda")
    o, d = pose.pose_rays(se3, poses, idx, dirs)
    torch.testing.assert_close(o, poses[idx.long(), :3, 3], rtol=0, atol=1e-6)
    torch.testing.assert_close(d, (poses[idx.long(), :3, :3] @ dirs.unsqueeze(-1)).squeeze(-1), rtol=1e-6, atol=1e-6)
    (o * 0).sum().backward()
    assert se3.grad.abs().max().item() == 0.0
    cam1 = se3.grad[1].clone()
    assert cam1.abs().max().item() == 0.0          # a camera without rays never receives gradient
