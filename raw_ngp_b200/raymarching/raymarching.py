"""raymarching -- occupancy-grid ray marching and compositing operators, B200 backend.

Mirror of the reference operator surface raymarching/raymarching.py:32-476 (same function names, positional
argument order, return values and autograd contract).  All native work goes through libngp_b200.so.
"""
import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib

__all__ = [
    "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "flatten_rays",
    "march_rays_train", "composite_rays_train", "march_rays", "composite_rays", "compact_rays_alive",
]


COOPERATIVE_MARCH = True


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _f32c(t):
    t = _cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32c(t):
    t = _cuda(t)
    if t.dtype != torch.int32:
        t = t.int()
    return t.contiguous()


# ----------------------------------------
# utils
# ----------------------------------------

class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o, rays_d [N, 3]; aabb [6] (xmin, ymin, zmin, xmax, ymax, zmax) -> nears [N], fars [N]."""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        aabb = _f32c(aabb)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        _lib.call("ngp_near_far_from_aabb", _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(aabb), N, float(min_near),
                  _lib.ptr(nears), _lib.ptr(fars), _lib.stream())
        ctx.mark_non_differentiable(nears, fars)
        return nears, fars

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_nears, grad_fars):
        return None, None, None, None


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        """Spherical coordinates (theta, phi in [-1, 1]) where each ray leaves the sphere |x| = radius."""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
        _lib.call("ngp_sph_from_ray", _lib.ptr(rays_o), _lib.ptr(rays_d), float(radius), N, _lib.ptr(coords),
                  _lib.stream())
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """coords [N, 3] int32 in [0, 1024) -> Morton indices [N] int32."""
        coords = _i32c(coords)
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _lib.call("ngp_morton3D", _lib.ptr(coords), N, _lib.ptr(indices), _lib.stream())
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """Morton indices [N] int32 -> coords [N, 3] int32."""
        indices = _i32c(indices)
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _lib.call("ngp_morton3D_invert", _lib.ptr(indices), N, _lib.ptr(coords), _lib.stream())
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        """grid [C, H^3] float, thresh scalar -> bitfield [C * H^3 / 8] uint8 (bit i of byte n = grid[8n+i] > thresh).

        Extension: `thresh` may be a tuple (mean_dev, density_thresh) with mean_dev a 1-element CUDA tensor; the
        threshold is then min(mean_dev, density_thresh) evaluated on the device (no host sync)."""
        grid = _f32c(grid)
        C, H3 = grid.shape[0], grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        thresh_dev = None
        if isinstance(thresh, tuple):
            thresh_dev, thresh = thresh
            thresh_dev = _f32c(thresh_dev)
        _lib.call("ngp_packbits", _lib.ptr(grid), N, float(thresh), _lib.ptr(thresh_dev), _lib.ptr(bitfield),
                  _lib.stream())
        return bitfield


packbits = _packbits.apply


class _flatten_rays(Function):
    @staticmethod
    def forward(ctx, rays, M):
        """rays [N, 2] (offset, count) -> res [M] ray id of every sample."""
        rays = _i32c(rays)
        N = rays.shape[0]
        res = torch.zeros(M, dtype=torch.int, device=rays.device)
        _lib.call("ngp_flatten_rays", _lib.ptr(rays), N, int(M), _lib.ptr(res), _lib.stream())
        return res


flatten_rays = _flatten_rays.apply

# ----------------------------------------
# train functions
# ----------------------------------------


class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, rays_ldir, bound, contract, density_bitfield, C, H, nears, fars, perturb=False,
                dt_gamma=0, max_steps=1024):
        """March rays through the occupancy bitfield and emit the training samples.

        Returns xyzs [M, 3], dirs [M, 3], ts [M, 2] = (t after the step, dt), rays [N, 2] int32 = (offset, count),
        ldirs [M, 3] or None.  Offsets are the exclusive prefix sum of the counts in ray order."""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        rays_ldir = _f32c(rays_ldir).view(-1, 3) if rays_ldir is not None else None
        density_bitfield = _cuda(density_bitfield).contiguous()
        nears = _f32c(nears).view(-1)
        fars = _f32c(fars).view(-1)
        N = rays_o.shape[0]
        dev = rays_o.device

        counter = torch.zeros(2, dtype=torch.int32, device=dev)
        if perturb:
            noises = torch.rand(N, dtype=torch.float32, device=dev)
        else:
            noises = torch.zeros(N, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
        # workspace of the warp-cooperative marcher: t of every kept sample (COOPERATIVE_MARCH = False selects the
        # one-thread-per-ray kernels; both give identical results)
        scratch = torch.empty(N * int(max_steps), dtype=torch.float32, device=dev) if COOPERATIVE_MARCH else None

        st = _lib.stream()
        _lib.call("ngp_march_rays_train_count", _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(density_bitfield),
                  float(bound), int(bool(contract)), float(dt_gamma), int(max_steps), N, int(C), int(H),
                  _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(noises), _lib.ptr(rays), _lib.ptr(counter), _lib.ptr(scratch), st)
        M = int(counter[0].item())  # the one host sync of the op (reference: raymarching.py:303)

        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        ts = torch.empty(M, 2, dtype=torch.float32, device=dev)
        ldirs = torch.empty(M, 3, dtype=torch.float32, device=dev) if rays_ldir is not None else None
        _lib.call("ngp_march_rays_train_write", _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.ptr(rays_ldir),
                  _lib.ptr(density_bitfield), float(bound), int(bool(contract)), float(dt_gamma), int(max_steps), N,
                  int(C), int(H), _lib.ptr(nears), _lib.ptr(fars), _lib.ptr(noises), _lib.ptr(rays), M, None,
                  _lib.ptr(scratch), _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ts), _lib.ptr(ldirs), st)

        ctx.save_for_backward(rays, ts)
        ctx.mark_non_differentiable(rays)
        return xyzs, dirs, ts, rays, ldirs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dxyzs, dL_ddirs, dL_dts, dL_drays, dL_dldirs):
        # dL/do = sum_seg dL/dxyz ; dL/dd = sum_seg (dL/dxyz * t + dL/ddirs)   (reference raymarching.py:319-329)
        rays, ts = ctx.saved_tensors
        N, M = rays.shape[0], ts.shape[0]
        dev = ts.device
        g_xyz = torch.zeros(M, 3, device=dev) if dL_dxyzs is None else dL_dxyzs.float().contiguous()
        g_dir = None if dL_ddirs is None else dL_ddirs.float().contiguous()
        dl_rays_o = torch.empty(N, 3, dtype=torch.float32, device=dev)
        dl_rays_d = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _lib.call("ngp_march_rays_train_backward", _lib.ptr(g_xyz), _lib.ptr(g_dir), _lib.ptr(ts), _lib.ptr(rays), N, M,
                  _lib.ptr(dl_rays_o), _lib.ptr(dl_rays_d), _lib.stream())
        return dl_rays_o, dl_rays_d, None, None, None, None, None, None, None, None, None, None, None


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, ts, rays, T_thresh=1e-4):
        """sigmas [M], rgbs [M, 3], ts [M, 2], rays [N, 2] -> weights [M], weights_sum [N], depth [N], image [N, 3]."""
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        ts = _f32c(ts)
        rays = _i32c(rays)
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights = torch.zeros(M, dtype=torch.float32, device=dev)  # samples after termination keep weight 0
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, 3, dtype=torch.float32, device=dev)
        _lib.call("ngp_composite_rays_train_forward", _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts), _lib.ptr(rays), M,
                  N, float(T_thresh), _lib.ptr(weights), _lib.ptr(weights_sum), _lib.ptr(depth), _lib.ptr(image),
                  _lib.stream())
        ctx.save_for_backward(sigmas, rgbs, ts, rays, weights_sum, depth, image)
        ctx.dims = (M, N, float(T_thresh))
        return weights, weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_weights, grad_weights_sum, grad_depth, grad_image):
        sigmas, rgbs, ts, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_weights = grad_weights.float().contiguous()
        grad_weights_sum = grad_weights_sum.float().contiguous()
        grad_depth = grad_depth.float().contiguous()
        grad_image = grad_image.float().contiguous()
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _lib.call("ngp_composite_rays_train_backward", _lib.ptr(grad_weights), _lib.ptr(grad_weights_sum),
                  _lib.ptr(grad_depth), _lib.ptr(grad_image), _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts),
                  _lib.ptr(rays), _lib.ptr(weights_sum), _lib.ptr(depth), _lib.ptr(image), M, N, T_thresh,
                  _lib.ptr(grad_sigmas), _lib.ptr(grad_rgbs), _lib.stream())
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply

# ----------------------------------------
# infer functions
# ----------------------------------------


class _march_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, contract, density_bitfield, C, H, near,
                far, perturb=False, dt_gamma=0, max_steps=1024):
        """March the first n_alive ids of rays_alive by up to n_step occupied samples each.

        Returns xyzs [n_alive*n_step, 3], dirs [n_alive*n_step, 3], ts [n_alive*n_step, 2]; slots a ray did not
        reach are zero (ts[:, 0] == 0 marks the ray as finished for composite_rays)."""
        rays_o = _f32c(rays_o).view(-1, 3)
        rays_d = _f32c(rays_d).view(-1, 3)
        rays_alive = _i32c(rays_alive)
        rays_t = _f32c(rays_t)
        near = _f32c(near).view(-1)
        far = _f32c(far).view(-1)
        density_bitfield = _cuda(density_bitfield).contiguous()
        dev = rays_o.device
        M = int(n_alive) * int(n_step)
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        ts = torch.empty(M, 2, dtype=torch.float32, device=dev)
        if perturb:
            noises = torch.rand(n_alive, dtype=torch.float32, device=dev)
        else:
            noises = torch.zeros(n_alive, dtype=torch.float32, device=dev)
        _lib.call("ngp_march_rays", int(n_alive), int(n_step), _lib.ptr(rays_alive), _lib.ptr(rays_t), _lib.ptr(rays_o),
                  _lib.ptr(rays_d), float(bound), int(bool(contract)), float(dt_gamma), int(max_steps), int(C), int(H),
                  _lib.ptr(density_bitfield), _lib.ptr(near), _lib.ptr(far), _lib.ptr(xyzs), _lib.ptr(dirs),
                  _lib.ptr(ts), _lib.ptr(noises), _lib.stream())
        return xyzs, dirs, ts


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2):
        """Accumulate n_step samples per alive ray into weights_sum / depth / image IN PLACE; finished rays get
        rays_alive[n] = -1, the others their new rays_t."""
        sigmas = _f32c(sigmas)
        rgbs = _f32c(rgbs)
        for name, t, dt in (("rays_alive", rays_alive, torch.int32), ("rays_t", rays_t, torch.float32),
                            ("weights_sum", weights_sum, torch.float32), ("depth", depth, torch.float32),
                            ("image", image, torch.float32)):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == dt):
                raise RuntimeError(f"composite_rays: in-place argument {name} must be a contiguous CUDA {dt} tensor")
        ts = _f32c(ts)
        _lib.call("ngp_composite_rays", int(n_alive), int(n_step), float(T_thresh), _lib.ptr(rays_alive),
                  _lib.ptr(rays_t), _lib.ptr(sigmas), _lib.ptr(rgbs), _lib.ptr(ts), _lib.ptr(weights_sum),
                  _lib.ptr(depth), _lib.ptr(image), _lib.stream())
        return tuple()


composite_rays = _composite_rays.apply


def compact_rays_alive(rays_alive, n_alive=None):
    """Device-side, order-preserving replacement of `rays_alive[rays_alive >= 0]` (nerf/renderer.py:612).
    Returns (buffer, count_tensor): the first count entries of `buffer` are the surviving ids."""
    rays_alive = _i32c(rays_alive)
    n = rays_alive.shape[0] if n_alive is None else int(n_alive)
    out = torch.empty_like(rays_alive)
    n_out = torch.zeros(1, dtype=torch.int32, device=rays_alive.device)
    ws = torch.empty((n + 4095) // 4096, dtype=torch.int32, device=rays_alive.device) if n > 4096 else None
    _lib.call("ngp_compact_rays_alive", _lib.ptr(rays_alive), n, _lib.ptr(out), _lib.ptr(n_out), _lib.ptr(ws), _lib.stream())
    return out, n_out
