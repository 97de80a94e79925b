"""TEST INFRASTRUCTURE ONLY -- calls the reference's *own* compiled CUDA extensions (oracle/_ref/*.so).

The reference's Python wrappers cannot travel to the GPU box (/root/reference does not exist there), so this
module restates only their tensor plumbing -- allocate outputs, call the pybind function, permute -- citing the
wrapper lines it follows.  All arithmetic happens inside the unmodified reference kernels.
"""
import glob
import importlib.util
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def available():
    return all(glob.glob(os.path.join(_HERE, "_ref", n + ".*.so")) for n in ("_gridencoder", "_raymarching_mob", "_shencoder"))


def module(name):
    if name not in _cache:
        hits = glob.glob(os.path.join(_HERE, "_ref", name + ".*.so"))
        if not hits:
            raise RuntimeError(f"oracle/_ref/{name} not built; run oracle/build_ref.sh where /root/reference exists")
        spec = importlib.util.spec_from_file_location(name, hits[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]


# ---- gridencoder/grid.py:27-69 ---------------------------------------------------------------------------------
def grid_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                 align_corners=False, interpolation=0, max_level=None):
    be = module("_gridencoder")
    inputs = inputs.contiguous()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = np.log2(per_level_scale)
    H = base_resolution
    max_level = L if max_level is None else min(max_level, L)
    outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
    if max_level < L:
        outputs.zero_()
    dy_dx = None
    if calc_grad_inputs:
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
        if max_level < L:
            dy_dx.zero_()
    be.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, max_level, S, H, dy_dx, gridtype,
                           align_corners, interpolation)
    return outputs.permute(1, 0, 2).reshape(B, L * C), dy_dx


# ---- gridencoder/grid.py:74-95 ---------------------------------------------------------------------------------
def grid_backward(grad, inputs, embeddings, offsets, per_level_scale, base_resolution, dy_dx=None, gridtype=0,
                  align_corners=False, interpolation=0, max_level=None):
    be = module("_gridencoder")
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = np.log2(per_level_scale)
    H = base_resolution
    max_level = L if max_level is None else min(max_level, L)
    grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
    grad_embeddings = torch.zeros_like(embeddings)
    grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
    be.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, max_level, S, H, dy_dx,
                            grad_inputs, gridtype, align_corners, interpolation)
    if grad_inputs is not None:
        grad_inputs = grad_inputs.to(inputs.dtype)
    return grad_inputs, grad_embeddings


# ---- gridencoder/grid.py:177-211 -------------------------------------------------------------------------------
def grid_total_variation(inputs, embeddings, grad, offsets, weight, per_level_scale, base_resolution, gridtype=0,
                         align_corners=False):
    be = module("_gridencoder")
    B, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    be.grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, np.log2(per_level_scale),
                            base_resolution, gridtype, align_corners)


def grid_weight_decay(embeddings, grad, offsets, weight):
    be = module("_gridencoder")
    be.grad_weight_decay(embeddings, grad, offsets, weight, embeddings.shape[0], embeddings.shape[1],
                         offsets.shape[0] - 1)


# ---- shencoder/sphere_harmonics.py:17-54 -----------------------------------------------------------------------
def sh_forward(inputs, degree, calc_grad_inputs=False):
    be = module("_shencoder")
    inputs = inputs.contiguous()
    B, D = inputs.shape
    outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
    dy_dx = torch.empty(B, D * degree ** 2, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs else None
    be.sh_encode_forward(inputs, outputs, B, D, degree, dy_dx)
    return outputs, dy_dx


def sh_backward(grad, inputs, degree, dy_dx):
    be = module("_shencoder")
    B, D = inputs.shape
    grad_inputs = torch.zeros_like(inputs)
    be.sh_encode_backward(grad.contiguous(), inputs, B, D, degree, dy_dx, grad_inputs)
    return grad_inputs


# ---- freqencoder/freq.py:17-53 -----------------------------------------------------------------------------------
def freq_forward(inputs, degree):
    be = module("_freqencoder")
    inputs = inputs.contiguous()
    B, D = inputs.shape
    C = D + D * 2 * degree
    outputs = torch.empty(B, C, dtype=inputs.dtype, device=inputs.device)
    be.freq_encode_forward(inputs, B, D, degree, C, outputs)
    return outputs


def freq_backward(grad, outputs, D, degree):
    be = module("_freqencoder")
    B, C = outputs.shape
    grad_inputs = torch.zeros(B, D, dtype=outputs.dtype, device=outputs.device)
    be.freq_encode_backward(grad.contiguous(), outputs, B, D, degree, C, grad_inputs)
    return grad_inputs


# ---- raymarching/raymarching.py ---------------------------------------------------------------------------------
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):  # :35-57
    be = module("_raymarching_mob")
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
    be.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):  # :75-98
    be = module("_raymarching_mob")
    N = rays_o.shape[0]
    coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
    be.sph_from_ray(rays_o, rays_d, radius, N, coords)
    return coords


def morton3D(coords):  # :104-119
    be = module("_raymarching_mob")
    N = coords.shape[0]
    indices = torch.empty(N, dtype=torch.int32, device=coords.device)
    be.morton3D(coords.int(), N, indices)
    return indices


def morton3D_invert(indices):  # :127-142
    be = module("_raymarching_mob")
    N = indices.shape[0]
    coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    be.morton3D_invert(indices.int(), N, coords)
    return coords


def packbits(grid, thresh, bitfield=None):  # :151-175
    be = module("_raymarching_mob")
    grid = grid.contiguous()
    N = grid.shape[0] * grid.shape[1] // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    be.packbits(grid, N, thresh, bitfield)
    return bitfield


def flatten_rays(rays, M):  # :182-200
    be = module("_raymarching_mob")
    res = torch.zeros(M, dtype=torch.int, device=rays.device)
    be.flatten_rays(rays.contiguous(), rays.shape[0], M, res)
    return res


def march_rays_train(rays_o, rays_d, rays_ldir, bound, contract, density_bitfield, C, H, nears, fars, noises,
                     dt_gamma=0, max_steps=1024):  # :254-317 (noises passed in so both sides see the same values)
    be = module("_raymarching_mob")
    N = rays_o.shape[0]
    dev = rays_o.device
    step_counter = torch.zeros(1, dtype=torch.int32, device=dev)
    rays = torch.empty(N, 2, dtype=torch.int32, device=dev)
    be.march_rays_train(rays_o, rays_d, rays_ldir, density_bitfield, bound, contract, dt_gamma, max_steps, N, C, H,
                        nears, fars, None, None, None, None, rays, step_counter, noises)
    M = step_counter.item()
    xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    ts = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
    ldirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev) if rays_ldir is not None else None
    be.march_rays_train(rays_o, rays_d, rays_ldir, density_bitfield, bound, contract, dt_gamma, max_steps, N, C, H,
                        nears, fars, xyzs, dirs, ts, ldirs, rays, step_counter, noises)
    return xyzs, dirs, ts, rays, ldirs


def composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh=1e-4):  # :336-368
    be = module("_raymarching_mob")
    M, N = sigmas.shape[0], rays.shape[0]
    dev = sigmas.device
    weights = torch.zeros(M, dtype=sigmas.dtype, device=dev)
    weights_sum = torch.empty(N, dtype=sigmas.dtype, device=dev)
    depth = torch.empty(N, dtype=sigmas.dtype, device=dev)
    image = torch.empty(N, 3, dtype=sigmas.dtype, device=dev)
    be.composite_rays_train_forward(sigmas, rgbs, ts, rays, M, N, T_thresh, weights, weights_sum, depth, image)
    return weights, weights_sum, depth, image


def composite_rays_train_backward(grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays,
                                  weights_sum, depth, image, T_thresh=1e-4):  # :370-387
    be = module("_raymarching_mob")
    M, N = sigmas.shape[0], rays.shape[0]
    grad_sigmas = torch.zeros_like(sigmas)
    grad_rgbs = torch.zeros_like(rgbs)
    be.composite_rays_train_backward(grad_weights, grad_weights_sum, grad_depth, grad_image, sigmas, rgbs, ts, rays,
                                     weights_sum, depth, image, M, N, T_thresh, grad_sigmas, grad_rgbs)
    return grad_sigmas, grad_rgbs


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, contract, density_bitfield, C, H, near,
               far, noises, dt_gamma=0, max_steps=1024):  # :399-442
    be = module("_raymarching_mob")
    dev = rays_o.device
    M = n_alive * n_step
    xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=dev)
    ts = torch.zeros(M, 2, dtype=rays_o.dtype, device=dev)
    be.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, contract, dt_gamma, max_steps, C, H,
                  density_bitfield, near, far, xyzs, dirs, ts, noises)
    return xyzs, dirs, ts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image,
                   T_thresh=1e-2):  # :450-468
    be = module("_raymarching_mob")
    be.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image)
