"""Fused bias-free ReLU MLP (the `MLP` of nerf/network.py:12-35) on the tcgen05 tensor cores.

`fused_mlp(x, w0, w1, ...)` is numerically the reference's  relu(...relu(x @ w0.T) @ w1.T...) @ wn.T  evaluated
with fp16 operands and fp32 accumulation (what nn.Linear does under torch.cuda.amp.autocast, renderer.py:546), in ONE
kernel per direction (csrc/mlp.cu).  Feature dimensions are zero-padded to multiples of 16 (31 -> 32, 3 -> 16).
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib

NGP_ACT_RELU = 1


def _pad16(n):
    return (n + 15) // 16 * 16


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _pad_cols(x, cols):
    """[M, c] -> contiguous fp16 [M, cols] (zero padded); no copy when already in that form."""
    if x.dtype == torch.float16 and x.shape[1] == cols and x.is_contiguous():
        return x
    out = torch.zeros(x.shape[0], cols, dtype=torch.float16, device=x.device) if x.shape[1] != cols else \
        torch.empty(x.shape[0], cols, dtype=torch.float16, device=x.device)
    out[:, :x.shape[1]] = x
    return out


class _fused_mlp(Function):
    @staticmethod
    def forward(ctx, x, *weights):
        _lib.require_cuda(x, *weights)
        M, K0 = x.shape
        dims = [K0] + [w.shape[0] for w in weights]
        pdims = [_pad16(d) for d in dims]
        L = len(weights)
        dev = x.device
        x16 = _pad_cols(x, pdims[0])
        w16 = []
        for l, w in enumerate(weights):
            wp = torch.zeros(pdims[l + 1], pdims[l], dtype=torch.float16, device=dev)
            wp[:dims[l + 1], :dims[l]] = w
            w16.append(wp)
        acts = [torch.empty(M, pdims[l + 1], dtype=torch.float16, device=dev) for l in range(L - 1)]
        y = torch.empty(M, pdims[L], dtype=torch.float16, device=dev)
        cd = (ctypes.c_uint32 * (L + 1))(*pdims)
        _lib.call("ngp_mlp_forward", _lib.ptr(x16), pdims[0], _ptr_array(w16), cd, L, M, NGP_ACT_RELU, _lib.ptr(y), pdims[L],
                  _ptr_array(acts), _lib.stream())
        ctx.save_for_backward(x16, *acts, *w16)
        ctx.meta = (dims, pdims, L, M, [w.dtype for w in weights], x.dtype)
        ctx.need_dx = x.requires_grad
        return y[:, :dims[L]]

    @staticmethod
    def backward(ctx, dy):
        dims, pdims, L, M, wdtypes, xdtype = ctx.meta
        saved = ctx.saved_tensors
        x16, acts, w16 = saved[0], list(saved[1:L]), list(saved[L:])
        dev = x16.device
        dy16 = _pad_cols(dy, pdims[L])
        dws = [torch.zeros(pdims[l + 1], pdims[l], dtype=torch.float32, device=dev) for l in range(L)]
        dx = torch.empty(M, pdims[0], dtype=torch.float16, device=dev) if ctx.need_dx else None
        cd = (ctypes.c_uint32 * (L + 1))(*pdims)
        _lib.call("ngp_mlp_backward", _lib.ptr(dy16), pdims[L], _lib.ptr(x16), pdims[0], _ptr_array(w16), _ptr_array(acts),
                  cd, L, M, NGP_ACT_RELU, _lib.ptr(dx), pdims[0], _ptr_array(dws), _lib.stream())
        gx = dx[:, :dims[0]].to(xdtype) if dx is not None else None
        gws = [dws[l][:dims[l + 1], :dims[l]].to(wdtypes[l]) for l in range(L)]
        return (gx, *gws)


def fused_mlp(x, *weights):
    """x [M, K0]; weights[l] [dims[l+1], dims[l]] (nn.Linear layout).  Returns [M, dims[-1]] fp16."""
    return _fused_mlp.apply(x, *weights)
