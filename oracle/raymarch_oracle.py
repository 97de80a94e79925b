"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libraymarch_oracle.so (C restatement of
raymarching/src/raymarching.cu, see raymarch_oracle.c).  numpy in, numpy out."""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int32, c_uint8, c_uint32

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libraymarch_oracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "raymarch_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "libraymarch_oracle.so"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_march_rays_train.restype = c_uint32
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t=c_float):
    return None if a is None else a.ctypes.data_as(POINTER(t))


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    o, d, bb = _f(rays_o), _f(rays_d), _f(aabb)
    N = o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(o), _p(d), _p(bb), c_uint32(N), c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    o, d = _f(rays_o), _f(rays_d)
    N = o.shape[0]
    out = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(o), _p(d), c_float(radius), c_uint32(N), _p(out))
    return out


def morton3D(coords):
    c = _i(coords)
    out = np.empty(c.shape[0], np.int32)
    lib().orc_morton3D(_p(c, c_int32), c_uint32(c.shape[0]), _p(out, c_int32))
    return out


def morton3D_invert(indices):
    ix = _i(indices)
    out = np.empty((ix.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(ix, c_int32), c_uint32(ix.shape[0]), _p(out, c_int32))
    return out


def packbits(grid, thresh):
    g = _f(grid).reshape(-1)
    N = g.shape[0] // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(g), c_uint32(N), c_float(thresh), _p(out, c_uint8))
    return out


def flatten_rays(rays, M):
    r = _i(rays)
    out = np.zeros(M, np.int32)
    lib().orc_flatten_rays(_p(r, c_int32), c_uint32(r.shape[0]), c_uint32(M), _p(out, c_int32))
    return out


def march_rays_train(rays_o, rays_d, rays_ldir, bound, contract, bitfield, C, H, nears, fars, noises, dt_gamma=0.0,
                     max_steps=1024):
    o, d = _f(rays_o), _f(rays_d)
    l = _f(rays_ldir) if rays_ldir is not None else None
    bf = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nr, fr, nz = _f(nears).reshape(-1), _f(fars).reshape(-1), _f(noises)
    N = o.shape[0]
    rays = np.empty((N, 2), np.int32)
    args = (c_float(bound), c_int(int(contract)), c_float(dt_gamma), c_uint32(max_steps), c_uint32(N), c_uint32(C),
            c_uint32(H), _p(nr), _p(fr), _p(nz), _p(rays, c_int32))
    M = lib().orc_march_rays_train(_p(o), _p(d), _p(l), _p(bf, c_uint8), *args, None, None, None, None)
    xyzs, dirs, ts = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    ldirs = np.zeros((M, 3), np.float32) if l is not None else None
    lib().orc_march_rays_train(_p(o), _p(d), _p(l), _p(bf, c_uint8), *args, _p(xyzs), _p(dirs), _p(ts), _p(ldirs))
    return xyzs, dirs, ts, rays, ldirs


def composite_rays_train_forward(sigmas, rgbs, ts, rays, T_thresh=1e-4):
    s, c, t, r = _f(sigmas), _f(rgbs), _f(ts), _i(rays)
    M, N = s.shape[0], r.shape[0]
    w = np.zeros(M, np.float32)
    ws, dp, im = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().orc_composite_rays_train_forward(_p(s), _p(c), _p(t), _p(r, c_int32), c_uint32(M), c_uint32(N), c_float(T_thresh),
                                           _p(w), _p(ws), _p(dp), _p(im))
    return w, ws, dp, im


def composite_rays_train_backward(gw, gws, gdp, gim, sigmas, rgbs, ts, rays, weights_sum, depth, image, T_thresh=1e-4):
    s, c, t, r = _f(sigmas), _f(rgbs), _f(ts), _i(rays)
    M, N = s.shape[0], r.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().orc_composite_rays_train_backward(_p(_f(gw)), _p(_f(gws)), _p(_f(gdp)), _p(_f(gim)), _p(s), _p(c), _p(t),
                                            _p(r, c_int32), _p(_f(weights_sum)), _p(_f(depth)), _p(_f(image)), c_uint32(M),
                                            c_uint32(N), c_float(T_thresh), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, contract, bitfield, C, H, nears, fars, noises,
               dt_gamma=0.0, max_steps=1024):
    o, d = _f(rays_o), _f(rays_d)
    al, rt = _i(rays_alive), _f(rays_t)
    bf = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    xyzs, dirs, ts = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().orc_march_rays(c_uint32(n_alive), c_uint32(n_step), _p(al, c_int32), _p(rt), _p(o), _p(d), c_float(bound),
                         c_int(int(contract)), c_float(dt_gamma), c_uint32(max_steps), c_uint32(C), c_uint32(H),
                         _p(bf, c_uint8), _p(_f(nears).reshape(-1)), _p(_f(fars).reshape(-1)), _p(xyzs), _p(dirs), _p(ts),
                         _p(_f(noises)))
    return xyzs, dirs, ts


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, ts, weights_sum, depth, image, T_thresh=1e-2):
    """In place on the numpy arrays rays_alive (int32), rays_t, weights_sum, depth, image (float32, contiguous)."""
    for a in (rays_t, weights_sum, depth, image):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    assert rays_alive.dtype == np.int32 and rays_alive.flags.c_contiguous
    lib().orc_composite_rays(c_uint32(n_alive), c_uint32(n_step), c_float(T_thresh), _p(rays_alive, c_int32), _p(rays_t),
                             _p(_f(sigmas)), _p(_f(rgbs)), _p(_f(ts)), _p(weights_sum), _p(depth), _p(image))
