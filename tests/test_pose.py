"""CPU tests: raw_ngp_b200/pose.py (torch formulation of barf/camera.py + camera_optimizers.py + train_utils.get_rays) against
tests/golden/pose.npz, which tools/make_golden_pose.py produced by running the reference's own Python code."""
import os

import numpy as np
import torch

from raw_ngp_b200 import pose

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose.npz")


def _load():
    z = np.load(GOLD)
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def test_se3_to_SE3_matches_reference():
    z = _load()
    torch.testing.assert_close(pose.se3_to_SE3(z["se3"]), z["SE3"], rtol=1e-6, atol=1e-7)
    # zero refinement is the identity, rotations are orthonormal
    assert torch.equal(pose.se3_to_SE3(torch.zeros(1, 6))[0], torch.eye(4)[:3])
    R = pose.se3_to_SE3(z["se3"])[:, :, :3]
    torch.testing.assert_close(R @ R.transpose(-1, -2), torch.eye(3).expand_as(R), rtol=0, atol=1e-5)


def test_refined_rays_and_gradients_match_reference():
    z = _load()
    se3 = z["se3"].clone().requires_grad_(True)
    idx = z["idx"].long()
    cam = pose.CameraOptimizer(se3.shape[0], "cpu")
    cam.se3_refine.weight.data.copy_(se3.data)
    refined = cam(z["poses"][idx], idx)
    torch.testing.assert_close(refined, z["refined"], rtol=1e-6, atol=1e-6)
    rays = pose.get_rays(refined, tuple(z["intrinsics"].tolist()), z["H"], z["W"], idx.shape[0], coords=z["coords"].long())
    torch.testing.assert_close(rays["rays_o"], z["rays_o"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rays["rays_d"], z["rays_d"], rtol=1e-6, atol=1e-6)
    assert torch.equal(rays["i"], z["i"]) and torch.equal(rays["j"], z["j"])
    ((rays["rays_o"] * z["g_rays_o"]).sum() + (rays["rays_d"] * z["g_rays_d"]).sum()).backward()
    torch.testing.assert_close(cam.se3_refine.weight.grad, z["d_se3"], rtol=1e-4, atol=1e-5)


def test_look_at_poses_are_rigid_and_face_the_origin():
    P = pose.look_at_poses(16)
    R = P[:, :3, :3]
    torch.testing.assert_close(R @ R.transpose(-1, -2), torch.eye(3).expand_as(R), rtol=0, atol=1e-5)
    fwd = -R[:, :, 2]                                    # the camera looks down -z
    c = P[:, :3, 3]
    torch.testing.assert_close(fwd, -c / c.norm(dim=-1, keepdim=True), rtol=0, atol=1e-5)


def test_numpy_oracle_matches_reference_fixture():
    """oracle/pose_oracle.py (independent numpy / float64 restatement) against the reference's outputs AND its autograd
    gradients (by central differences)"""
    from oracle import pose_oracle
    z = np.load(GOLD)
    np.testing.assert_allclose(pose_oracle.se3_to_SE3(z["se3"]), z["SE3"], rtol=0, atol=2e-6)
    H, W = int(z["H"]), int(z["W"])
    fx, fy, cx, cy = z["intrinsics"].tolist()
    i, j = z["i"].astype(np.float64) + 0.5, z["j"].astype(np.float64) + 0.5
    dirs = np.stack([(i - cx) / fx, -(j - cy) / fy, -np.ones_like(i)], axis=-1)
    o, d = pose_oracle.refined_rays(z["se3"], z["poses"], z["idx"], dirs)
    np.testing.assert_allclose(o, z["rays_o"], rtol=0, atol=3e-6)
    np.testing.assert_allclose(d, z["rays_d"], rtol=0, atol=3e-6)
    g = pose_oracle.d_se3_numeric(z["se3"], z["poses"], z["idx"], dirs, z["g_rays_o"].astype(np.float64), z["g_rays_d"].astype(np.float64))
    np.testing.assert_allclose(g, z["d_se3"], rtol=2e-4, atol=2e-4)
