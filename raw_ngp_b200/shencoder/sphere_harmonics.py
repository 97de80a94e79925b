"""SHEncoder -- real spherical harmonics of view / light directions, B200 backend.

Mirror of shencoder/sphere_harmonics.py:14-89 of the reference (same names and arguments).  Backward
recomputes the Jacobian from the saved 12-byte inputs instead of storing dy_dx [B, 3*degree^2].
"""
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib


class _sh_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        # inputs [B, 3] float in [-1, 1]  ->  [B, degree^2]
        _lib.require_cuda(inputs)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, input_dim = inputs.shape
        if input_dim != 3:
            raise RuntimeError("SH encoder only supports input dim == 3")
        outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
        _lib.call("ngp_sh_encode_forward", _lib.ptr(inputs), _lib.ptr(outputs), B, int(degree), None, _lib.NGP_F32,
                  _lib.stream())
        ctx.save_for_backward(inputs)
        ctx.dims = (B, input_dim, int(degree))
        ctx.calc_grad_inputs = bool(calc_grad_inputs)
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        (inputs,) = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad = grad.contiguous()
        if grad.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            grad = grad.float()
        grad_inputs = torch.zeros_like(inputs)
        _lib.call("ngp_sh_encode_backward", _lib.ptr(grad), _lib.ptr(inputs), B, degree, _lib.ptr(grad_inputs),
                  _lib.dtype_id(grad.dtype), _lib.stream())
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


def sh_encode_with_jacobian(inputs, degree):
    """Forward that also returns dy_dx [B, 3*degree^2] in the reference's layout (shencoder.cu:124-127)."""
    inputs = inputs.contiguous().float()
    B = inputs.shape[0]
    outputs = torch.empty(B, degree ** 2, dtype=torch.float32, device=inputs.device)
    dy_dx = torch.empty(B, 3 * degree ** 2, dtype=torch.float32, device=inputs.device)
    _lib.call("ngp_sh_encode_forward", _lib.ptr(inputs), _lib.ptr(outputs), B, int(degree), _lib.ptr(dy_dx),
              _lib.NGP_F32, _lib.stream())
    return outputs, dy_dx


class SHEncoder(nn.Module):
    """Real spherical harmonics of a direction, degree 1..8 -> degree^2 coefficients (SHEncoder of the reference)."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        if input_dim != 3:
            raise AssertionError("SH encoder only support input dim == 3")
        if not 0 < degree <= 8:
            raise AssertionError("SH encoder only supports degree in [1, 8]")
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        # [..., 3] in [-size, size]: scaled to the unit cube, renormalised to the unit sphere (sphere_harmonics.py:76-81),
        # encoded; gradients with respect to the directions only when they are asked for
        lead = inputs.shape[:-1]
        unit = inputs / size
        unit = unit / torch.norm(unit, dim=-1, keepdim=True)
        flat = unit.reshape(-1, self.input_dim)
        return sh_encode(flat, self.degree, flat.requires_grad).reshape(*lead, self.output_dim)
