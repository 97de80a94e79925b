"""Fused NeRF field: NeRFNetwork.forward / .density (nerf/network.py:74-156) as one forward and one backward kernel
(csrc/field_ws.cu, csrc/field_bwd_ws.cu; layer widths outside {16, 32, 64}, i.e. the light-stage view_mlp, use the kernel
pairs of csrc/field.cu + csrc/mlp.cu) instead of ~60 PyTorch ops.  Used by raw_ngp_b200.nerf.NeRFNetwork when the
configuration is eligible (fp16 table with F=2, ReLU MLPs, autocast, no gradient w.r.t. positions/directions); otherwise
the network composes the individual operators exactly like the reference does.  (Gradients w.r.t. positions / directions
inside the fused path exist in FusedTrainStep, raw_ngp_b200/trainer.py.)"""
import ctypes

import numpy as np
import torch
from torch.autograd import Function

from . import _lib
from .ffmlp import NGP_ACT_RELU, _pad16, _ptr_array

DENSITY_ACT = {"clamped_exp": 0, "softplus": 1}
COLOR_ACT = {"exp": 1, "sigmoid": 2, "clamped_exp": 3}


def _pad_weight(w, n, k):
    """fp16 copy of w [N, K] zero-padded to [n, k].  The copy is cached ON the weight tensor object and reused until the
    tensor is written again (its version counter changes): inference evaluates the field ~40 times per frame with fixed
    weights.  The native optimizers write parameters through raw pointers, also from replayed CUDA graphs, where torch's
    version counter does not move: _lib.weights_epoch stands in for it."""
    ver = (w._version, _lib.weights_epoch, w.data_ptr(), tuple(w.shape), n, k, w.device)
    hit = getattr(w, "_ngp_padded", None)
    if hit is not None and hit[0] == ver:
        return hit[1]
    wp = torch.zeros(n, k, dtype=torch.float16, device=w.device)
    wp[:w.shape[0], :w.shape[1]] = w.detach()
    try:
        w._ngp_padded = (ver, wp)
    except (AttributeError, RuntimeError):
        pass
    return wp


def _grid_scalars(enc):
    return (float(np.log2(enc.per_level_scale)), int(enc.base_resolution), int(enc.num_levels), int(enc.gridtype_id),
            int(bool(enc.align_corners)), int(enc.interp_id))


def density_only(enc, grid_weights, xyzs, bound, density_act, beta, feat_weights=None):
    """sigma [M] fp32 for NeRFNetwork.density (no gradient)."""
    xyzs = xyzs.contiguous().float()
    M = xyzs.shape[0]
    dims = [w.shape[1] for w in grid_weights[:1]] + [w.shape[0] for w in grid_weights]
    pdims = [_pad16(d) for d in dims]
    w16 = [_pad_weight(w, pdims[i + 1], pdims[i]) for i, w in enumerate(grid_weights)]
    sigma = torch.empty(M, dtype=torch.float32, device=xyzs.device)
    S, H, L, gt, ac, ip = _grid_scalars(enc)
    cd = (ctypes.c_uint32 * len(pdims))(*pdims)
    if all(d in (16, 32, 64) for d in pdims[:3]) and L % 8 == 0:      # warp-specialised kernel, grid_mlp only
        _lib.call("ngp_field_forward_full", _lib.ptr(xyzs), None, None, _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets),
                  _lib.ptr(feat_weights), float(bound), S, H, L, gt, ac, ip, _ptr_array(w16), cd, None, None, M, None,
                  int(density_act), float(beta), 1, None, None, None, None, _lib.ptr(sigma), None, None, _lib.stream())
    else:
        _lib.call("ngp_field_forward_density", _lib.ptr(xyzs), None, None, _lib.ptr(enc.embeddings), _lib.ptr(enc.offsets),
                  _lib.ptr(feat_weights), float(bound), S, H, L, gt, ac, ip, _ptr_array(w16), cd, len(w16), M, None,
                  int(density_act), float(beta), None, None, _lib.ptr(sigma), None, 0, _lib.stream())
    return sigma


def _ws_ok(p1, p2):
    """The warp-specialised kernels keep every layer input as a swizzled tile of width 16 / 32 / 64 (csrc/tile_sw.cuh); wider
    layers (rfield: 48-wide input, 80-wide hidden) take the two-kernel path."""
    return all(d in (16, 32, 64) for d in list(p1[:3]) + list(p2[:3]))


def _ws_fwd_ok(p1, p2, L):
    """The warp-specialised FORWARD takes any layer width that is a multiple of 16 up to 128 (16-column panels, csrc/tile_sw.cuh):
    the light-stage view_mlp (48 -> 80 -> 80) runs in it, with the saved tensors written as plain rows for the kernel-pair
    backward.  (The warp-specialised backward keeps all six weight-gradient accumulators in TMEM and both chains' tiles in
    shared memory; at 80-wide layers neither fits -- 544 of 512 columns, 245 of 227 KB.)"""
    return L % 8 == 0 and all(d % 16 == 0 and 16 <= d <= 128 for d in list(p1) + list(p2))


class _fused_field(Function):
    @staticmethod
    def forward(ctx, xyzs, dirs, ldirs, table, feat_weights, cfg, *weights):
        enc, bound, density_act, beta, color_act = cfg
        n1 = len(weights) // 2
        gw, vw = weights[:n1], weights[n1:]
        xyzs = xyzs.contiguous().float()
        dirs = dirs.contiguous().float()
        ldirs = ldirs.contiguous().float() if ldirs is not None else None
        M, dev = xyzs.shape[0], xyzs.device
        d1 = [gw[0].shape[1]] + [w.shape[0] for w in gw]
        d2 = [vw[0].shape[1]] + [w.shape[0] for w in vw]
        p1, p2 = [_pad16(d) for d in d1], [_pad16(d) for d in d2]
        w1 = [_pad_weight(w, p1[i + 1], p1[i]) for i, w in enumerate(gw)]
        w2 = [_pad_weight(w, p2[i + 1], p2[i]) for i, w in enumerate(vw)]
        keep = any(ctx.needs_input_grad)   # (grad mode is off inside Function.forward; this is the real signal)
        ws = _ws_ok(p1, p2)
        ws_fwd = ws or _ws_fwd_ok(p1, p2, enc.num_levels)
        f16 = dict(dtype=torch.float16, device=dev)
        Mt = (M + 127) // 128 * 128 if ws else M      # the WS kernels save whole 128-row tiles (tile-panel layout)
        enc_buf = torch.empty(Mt, p1[0], **f16) if keep else None
        acts1 = [torch.empty(Mt, p1[l + 1], **f16) if keep else None for l in range(len(w1) - 1)]
        acts2 = [torch.empty(Mt, p2[l + 1], **f16) if keep else None for l in range(len(w2) - 1)]
        in2 = torch.empty(Mt, p2[0], **f16) if (keep or not ws_fwd) else None
        sigma = torch.empty(M, dtype=torch.float32, device=dev)
        rgb = torch.empty(M, 3, dtype=torch.float32, device=dev)
        S, H, L, gt, ac, ip = _grid_scalars(enc)
        st = _lib.stream()
        c1 = (ctypes.c_uint32 * len(p1))(*p1)
        c2 = (ctypes.c_uint32 * len(p2))(*p2)
        if ws_fwd:
            _lib.call("ngp_field_forward_full", _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ldirs), _lib.ptr(table), _lib.ptr(enc.offsets),
                      _lib.ptr(feat_weights), float(bound), S, H, L, gt, ac, ip, _ptr_array(w1), c1, _ptr_array(w2), c2, M, None,
                      int(density_act), float(beta), int(color_act), _lib.ptr(enc_buf), _ptr_array(acts1) if keep else None,
                      _lib.ptr(in2), _ptr_array(acts2) if keep else None, _lib.ptr(sigma), _lib.ptr(rgb), None, st)
        else:
            _lib.call("ngp_field_forward_density", _lib.ptr(xyzs), _lib.ptr(dirs), _lib.ptr(ldirs), _lib.ptr(table),
                      _lib.ptr(enc.offsets), _lib.ptr(feat_weights), float(bound), S, H, L, gt, ac, ip, _ptr_array(w1), c1, len(w1), M,
                      None, int(density_act), float(beta), _lib.ptr(enc_buf), _ptr_array(acts1) if keep else None, _lib.ptr(sigma),
                      _lib.ptr(in2), p2[0], st)
            _lib.call("ngp_mlp_forward_rgb", _lib.ptr(in2), p2[0], _ptr_array(w2), c2, len(w2), M, None, NGP_ACT_RELU, int(color_act),
                      _lib.ptr(rgb), _ptr_array(acts2) if keep else None, st)
        if keep:
            ctx.save_for_backward(xyzs, enc_buf, in2, sigma, rgb, table, feat_weights if feat_weights is not None else xyzs.new_empty(0),
                                  *acts1, *acts2, *w1, *w2)
            ctx.meta = (enc, bound, density_act, beta, color_act, d1, d2, p1, p2, [w.dtype for w in weights],
                        feat_weights is not None)
        ctx.mark_non_differentiable()
        return sigma, rgb

    @staticmethod
    def backward(ctx, d_sigma, d_rgb):
        enc, bound, density_act, beta, color_act, d1, d2, p1, p2, wdt, has_fw = ctx.meta
        sv = ctx.saved_tensors
        xyzs, enc_buf, in2, sigma, rgb, table, fw = sv[:7]
        n1, n2 = len(p1) - 1, len(p2) - 1
        o = 7
        acts1 = list(sv[o:o + n1 - 1]); o += n1 - 1
        acts2 = list(sv[o:o + n2 - 1]); o += n2 - 1
        w1 = list(sv[o:o + n1]); o += n1
        w2 = list(sv[o:o + n2])
        M, dev = xyzs.shape[0], xyzs.device
        d_sigma = d_sigma.contiguous().float()
        d_rgb = d_rgb.contiguous().float()
        dw1 = [torch.zeros(p1[l + 1], p1[l], dtype=torch.float32, device=dev) for l in range(n1)]
        dw2 = [torch.zeros(p2[l + 1], p2[l], dtype=torch.float32, device=dev) for l in range(n2)]
        sink = enc.grad_sink
        if sink is not None:
            if sink.shape != table.shape or sink.dtype != table.dtype:
                raise RuntimeError("grad_sink must match the table")
            gtable = sink
        else:
            gtable = torch.zeros_like(table)
        S, H, L, gt, ac, ip = _grid_scalars(enc)
        st = _lib.stream()
        c1 = (ctypes.c_uint32 * len(p1))(*p1)
        c2 = (ctypes.c_uint32 * len(p2))(*p2)
        fwp = _lib.ptr(fw) if has_fw else None
        if _ws_ok(p1, p2):
            _lib.call("ngp_field_backward_full", _lib.ptr(xyzs), _lib.ptr(d_sigma), _lib.ptr(sigma), _lib.ptr(d_rgb), _lib.ptr(rgb),
                      _lib.ptr(enc_buf), _ptr_array(acts1), _lib.ptr(in2), _ptr_array(acts2), _lib.ptr(enc.offsets), fwp, float(bound),
                      S, H, L, gt, ac, ip, _ptr_array(w1), c1, _ptr_array(w2), c2, M, None, int(density_act), float(beta),
                      int(color_act), _lib.ptr(gtable), _ptr_array(dw1), _ptr_array(dw2), None, None, None, None, st)
        else:
            d_in2 = torch.empty(M, p2[0], dtype=torch.float16, device=dev)
            _lib.call("ngp_mlp_backward_rgb", _lib.ptr(d_rgb), _lib.ptr(rgb), int(color_act), _lib.ptr(in2), p2[0], _ptr_array(w2),
                      _ptr_array(acts2), c2, n2, M, None, NGP_ACT_RELU, _lib.ptr(d_in2), p2[0], _ptr_array(dw2), st)
            _lib.call("ngp_field_backward_density", _lib.ptr(xyzs), _lib.ptr(d_sigma), _lib.ptr(sigma), _lib.ptr(d_in2), p2[0],
                      _lib.ptr(enc_buf), None, _lib.ptr(enc.offsets), fwp, float(bound), S, H, L, gt, ac, ip, _ptr_array(w1),
                      _ptr_array(acts1), c1, n1, M, None, int(density_act), float(beta), _lib.ptr(gtable), _ptr_array(dw1), None, st)
        gw = [dw1[l][:d1[l + 1], :d1[l]].to(wdt[l]) for l in range(n1)]
        vw = [dw2[l][:d2[l + 1], :d2[l]].to(wdt[n1 + l]) for l in range(n2)]
        return (None, None, None, None if sink is not None else gtable, None, None, *gw, *vw)


def fused_field(xyzs, dirs, ldirs, enc, grid_weights, view_weights, bound, density_act, beta, color_act, feat_weights=None):
    """Returns sigma [M] fp32, rgb [M,3] fp32."""
    cfg = (enc, bound, density_act, beta, color_act)
    return _fused_field.apply(xyzs, dirs, ldirs, enc.embeddings, feat_weights, cfg, *grid_weights, *view_weights)
