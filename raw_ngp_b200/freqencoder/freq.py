"""FreqEncoder -- frequency (positional) encoding, B200 backend.

Mirror of freqencoder/freq.py:15-77 of the reference (same names and arguments): [.., D] -> [.., D + 2 D degree].  The
backward uses the saved outputs as the derivatives, like the reference (freqencoder.cu:60-94)."""
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib


class _freq_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        # inputs [B, input_dim] float -> [B, output_dim]
        if not inputs.is_cuda:
            inputs = inputs.cuda()
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=inputs.dtype, device=inputs.device)
        _lib.call("ngp_freq_encode_forward", _lib.ptr(inputs), B, input_dim, int(degree), int(output_dim), _lib.ptr(outputs), _lib.stream())
        ctx.save_for_backward(outputs)
        ctx.dims = (B, input_dim, int(degree), int(output_dim))
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        (outputs,) = ctx.saved_tensors
        B, input_dim, degree, output_dim = ctx.dims
        grad = grad.contiguous().float()
        grad_inputs = torch.empty(B, input_dim, dtype=torch.float32, device=grad.device)
        _lib.call("ngp_freq_encode_backward", _lib.ptr(grad), _lib.ptr(outputs), B, input_dim, degree, output_dim, _lib.ptr(grad_inputs),
                  _lib.stream())
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        # inputs [..., input_dim] -> [..., output_dim]
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        outputs = freq_encode(inputs, self.degree, self.output_dim)
        return outputs.reshape(prefix_shape + [self.output_dim])
