"""Camera pose refinement and ray generation: the step either side of march_rays_train when
`--pose_opt barf` is on.

Mirrors, with the same names and argument meaning:
  * barf/camera.py:65-153       Lie.se3_to_SE3 / skew_symmetric / taylor_A,B,C, Pose.compose_pair   (torch, differentiable)
  * barf/camera_optimizers.py   CameraOptimizer: se3_refine Embedding [num_cameras, 6] (zeros), forward(poses, indices)
  * nerf/train_utils.py:96-172  get_rays (random pixel sampling, camera-space directions, no normalisation)
and adds the native path: `pose_rays(se3, poses, cam_idx, dirs_cam)` = provide_refined_poses + get_rays' matrix part in ONE
kernel (csrc/pose.cu), differentiable with respect to se3 through a second kernel -- what FusedTrainStep(pose=...) captures.

Precision note: the reference evaluates provide_refined_poses under autocast, so its `@` products (wx @ wx, V @ u, the
compose) are rounded to fp16 before `Pose.__call__` casts back to fp32 (camera.py:33-34); here everything stays fp32.
"""
import torch
from torch import nn
from torch.autograd import Function

from . import _lib


# ---- torch formulation (reference semantics; used by the autograd path and as the check of the kernels) ---------------
def skew_symmetric(w):
    w0, w1, w2 = w.unbind(dim=-1)
    O = torch.zeros_like(w0)
    return torch.stack([torch.stack([O, -w2, w1], dim=-1), torch.stack([w2, O, -w0], dim=-1), torch.stack([-w1, w0, O], dim=-1)], dim=-2)


def _taylor(x, first, nth=10):
    """sum_i (-1)^i x^(2i) / (2i + first)!   (first = 1: sin(x)/x, 2: (1-cos x)/x^2, 3: (x-sin x)/x^3; camera.py:124-153)"""
    ans = torch.zeros_like(x)
    denom = 1.0
    for i in range(nth + 1):
        if first == 1:
            if i > 0:
                denom *= (2 * i) * (2 * i + 1)
        else:
            denom *= (2 * i + first - 1) * (2 * i + first)
        ans = ans + (-1) ** i * x ** (2 * i) / denom
    return ans


def se3_to_SE3(wu):
    """[..., 6] (w, u) -> [..., 3, 4] = [R | V u]   (camera.py:93-105)"""
    w, u = wu.split([3, 3], dim=-1)
    wx = skew_symmetric(w)
    theta = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, device=w.device, dtype=torch.float32)
    A, B, C = _taylor(theta, 1), _taylor(theta, 2), _taylor(theta, 3)
    R = eye + A * wx + B * wx @ wx
    V = eye + B * wx + C * wx @ wx
    return torch.cat([R, V @ u[..., None]], dim=-1)


def compose_pair(pose_a, pose_b):
    """pose_new(x) = pose_b o pose_a (x)   (camera.py:56-63)"""
    R_a, t_a = pose_a[..., :3], pose_a[..., 3:]
    R_b, t_b = pose_b[..., :3], pose_b[..., 3:]
    return torch.cat([R_b @ R_a, R_b @ t_a + t_b], dim=-1)


def pixel_directions(i, j, intrinsics):
    """camera-space directions of pixel centres (i + 0.5, j + 0.5): z flipped, y flipped, not normalised (train_utils.py:150-157)"""
    fx, fy, cx, cy = intrinsics
    return torch.stack(((i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)), dim=-1)


def get_rays(poses, intrinsics, H, W, N=-1, coords=None, ldirs=None, generator=None):
    """poses [N or 1, 4, 4] / [.., 3, 4] cam2world, intrinsics (fx, fy, cx, cy) -> dict(rays_o, rays_d [N, 3], i, j, rays_ldir)
    (train_utils.py:96-172 without the patch sampler)."""
    device = poses.device
    if N > 0:
        if coords is not None:
            inds = coords[:, 0] * W + coords[:, 1]
        else:
            inds = torch.randint(0, H * W, size=[N], device=device, generator=generator)
    else:
        inds = torch.arange(H * W, device=device)
    j = (inds // W).float() + 0.5
    i = (inds % W).float() + 0.5
    directions = pixel_directions(i, j, intrinsics)
    rays_d = (directions.unsqueeze(1) @ poses[:, :3, :3].transpose(-1, -2)).squeeze(1)
    rays_o = poses[:, :3, 3].expand_as(rays_d)
    out = {"rays_o": rays_o, "rays_d": rays_d, "rays_ldir": ldirs.expand_as(rays_d) if ldirs is not None else None}
    if N > 0:
        out["i"], out["j"] = i.long(), j.long()
    return out


# ---- native path ------------------------------------------------------------------------------------------------------
class _pose_rays(Function):
    @staticmethod
    def forward(ctx, se3, poses, cam_idx, dirs_cam):
        se3 = se3.contiguous().float()
        poses = poses.contiguous().float()
        cam_idx = cam_idx.contiguous().int()
        dirs_cam = dirs_cam.contiguous().float()
        N = cam_idx.shape[0]
        stride = poses.shape[-2] * poses.shape[-1]
        rays_o = torch.empty(N, 3, dtype=torch.float32, device=se3.device)
        rays_d = torch.empty(N, 3, dtype=torch.float32, device=se3.device)
        _lib.call("ngp_pose_rays_forward", _lib.ptr(se3), _lib.ptr(poses), stride, _lib.ptr(cam_idx), _lib.ptr(dirs_cam), N,
                  se3.shape[0], _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.stream())
        ctx.save_for_backward(se3, poses, cam_idx, dirs_cam)
        return rays_o, rays_d

    @staticmethod
    def backward(ctx, d_o, d_d):
        se3, poses, cam_idx, dirs_cam = ctx.saved_tensors
        N = cam_idx.shape[0]
        d_o = torch.zeros(N, 3, device=se3.device) if d_o is None else d_o.contiguous().float()
        d_d = torch.zeros(N, 3, device=se3.device) if d_d is None else d_d.contiguous().float()
        d_se3 = torch.zeros_like(se3)
        _lib.call("ngp_pose_rays_backward", _lib.ptr(d_o), _lib.ptr(d_d), _lib.ptr(se3), _lib.ptr(poses),
                  poses.shape[-2] * poses.shape[-1], _lib.ptr(cam_idx), _lib.ptr(dirs_cam), N, se3.shape[0], _lib.ptr(d_se3),
                  _lib.stream())
        return d_se3, None, None, None


def pose_rays(se3, poses, cam_idx, dirs_cam):
    """rays_o, rays_d [N, 3] of pixels `dirs_cam` seen from cameras `cam_idx` whose dataset poses [C, 3|4, 4] are refined by
    se3 [C, 6]; differentiable with respect to se3."""
    return _pose_rays.apply(se3, poses, cam_idx, dirs_cam)


class CameraOptimizer(nn.Module):
    """barf/camera_optimizers.py:14-107: learnable se(3) correction per camera, composed in front of the dataset pose."""

    def __init__(self, num_cameras, device, opt=None):
        super().__init__()
        self.num_cameras, self.device, self.opt = num_cameras, device, opt
        self.annealing = 0.0
        self.pose_noise = None
        noise = float(getattr(opt, "noise", 0.0) or 0.0)
        if noise > 0.0:      # synthetic perturbation of the initial poses (camera_optimizers.py:25-36)
            scale = float(getattr(opt, "scale", 1.0))
            se3_noise_t = torch.randn(num_cameras, 3, device=device) * noise * scale
            se3_noise_r = torch.randn(num_cameras, 3, device=device) * noise
            self.pose_noise = se3_to_SE3(torch.cat([se3_noise_t, se3_noise_r], dim=-1))
        self.se3_refine = nn.Embedding(num_cameras, 6, device=device)
        nn.init.zeros_(self.se3_refine.weight)

    def update_annealing(self, new_value):
        self.annealing = new_value

    def provide_refined_poses(self, poses, indices):
        poses = poses[:, :3, :]
        if self.pose_noise is not None:
            poses = compose_pair(self.pose_noise[indices], poses)
        pose_refine = se3_to_SE3(self.se3_refine.weight[indices])
        return compose_pair(pose_refine, poses)

    def forward(self, poses, indices):
        return self.provide_refined_poses(poses, indices)

    def get_refined_poses(self, poses_gt):
        return self(poses_gt, torch.arange(0, self.num_cameras, device=poses_gt.device).long())

    def get_params(self):
        return list(self.parameters())


def look_at_poses(n, radius=2.0, seed=4):
    """n synthetic camera-to-world poses [n, 4, 4] on the sphere of `radius`, looking at the origin (OpenGL convention: the
    camera looks down -z, y up) -- the synthetic stand-in for a dataset's poses in the benchmarks."""
    g = torch.Generator().manual_seed(seed)
    c = torch.randn(n, 3, generator=g)
    c = c / c.norm(dim=-1, keepdim=True) * radius
    fwd = -c / c.norm(dim=-1, keepdim=True)
    up = torch.tensor([0.0, 0.0, 1.0]).expand(n, 3)
    right = torch.cross(fwd, up, dim=-1)
    right = right / right.norm(dim=-1, keepdim=True).clamp(min=1e-6)
    up2 = torch.cross(right, fwd, dim=-1)
    poses = torch.eye(4).repeat(n, 1, 1)
    poses[:, :3, 0], poses[:, :3, 1], poses[:, :3, 2], poses[:, :3, 3] = right, up2, -fwd, c
    return poses
