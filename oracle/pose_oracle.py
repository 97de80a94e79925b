"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64 accumulate) of the pose path:
se3 -> SE3 (barf/camera.py:93-153, Taylor series of order 10), composition with the dataset pose (camera.py:56-63),
ray generation (nerf/train_utils.py:150-165).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import
this.  Pinned by tests/golden/pose.npz (outputs and gradients of the reference's own Python code)."""
import math

import numpy as np


def _series(s, first, order=10):
    """sum_i (-1)^i s^i / (2i + first)!  with s = theta^2"""
    return sum((-1) ** i * s ** i / math.factorial(2 * i + first) for i in range(order + 1))


def se3_to_SE3(wu):
    """[C, 6] -> [C, 3, 4]"""
    wu = np.asarray(wu, dtype=np.float64)
    out = np.zeros((wu.shape[0], 3, 4))
    for c, (w0, w1, w2, u0, u1, u2) in enumerate(wu):
        wx = np.array([[0, -w2, w1], [w2, 0, -w0], [-w1, w0, 0]])
        s = w0 * w0 + w1 * w1 + w2 * w2
        A, B, C = _series(s, 1), _series(s, 2), _series(s, 3)
        out[c, :, :3] = np.eye(3) + A * wx + B * wx @ wx
        out[c, :, 3] = (np.eye(3) + B * wx + C * wx @ wx) @ np.array([u0, u1, u2])
    return out


def refined_rays(se3, poses, cam_idx, dirs_cam):
    """rays_o, rays_d [N, 3]: pose = dataset pose o refinement (R = R_p R_r, t = R_p t_r + t_p); d = R dir, o = t"""
    ref = se3_to_SE3(se3)
    poses = np.asarray(poses, dtype=np.float64)[:, :3, :]
    R = poses[:, :, :3] @ ref[:, :, :3]
    t = np.einsum("cij,cj->ci", poses[:, :, :3], ref[:, :, 3]) + poses[:, :, 3]
    idx = np.asarray(cam_idx)
    rays_d = np.einsum("nij,nj->ni", R[idx], np.asarray(dirs_cam, dtype=np.float64))
    return t[idx], rays_d


def d_se3_numeric(se3, poses, cam_idx, dirs_cam, g_o, g_d, eps=1e-6):
    """central differences of sum(rays_o * g_o) + sum(rays_d * g_d) with respect to se3"""
    se3 = np.asarray(se3, dtype=np.float64)
    grad = np.zeros_like(se3)

    def f(x):
        o, d = refined_rays(x, poses, cam_idx, dirs_cam)
        return float((o * g_o).sum() + (d * g_d).sum())
    for c in range(se3.shape[0]):
        for k in range(6):
            hi, lo = se3.copy(), se3.copy()
            hi[c, k] += eps
            lo[c, k] -= eps
            grad[c, k] = (f(hi) - f(lo)) / (2 * eps)
    return grad
