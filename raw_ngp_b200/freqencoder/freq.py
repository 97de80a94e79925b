"""Frequency (positional) encoding on the B200 backend: [..., D] -> [..., D + 2 * D * degree].

Public surface as in the reference (freqencoder/freq.py:15-77): `freq_encode(inputs, degree, output_dim)` and the module
`FreqEncoder(input_dim=3, degree=4)` with `.output_dim`.  The kernels are `ngp_freq_encode_forward / _backward`
(csrc/freq_encode.cu); like the reference's, the backward derives d sin / d x from the stored outputs (the cosine column is
the derivative of the sine column and vice versa), so the forward keeps its result for autograd."""
import torch
from torch import nn

from .. import _lib


def _launch_forward(points, degree, width):
    count, dim = points.shape
    encoded = points.new_empty(count, width)
    _lib.call("ngp_freq_encode_forward", _lib.ptr(points), count, dim, degree, width, _lib.ptr(encoded), _lib.stream())
    return encoded


def _launch_backward(upstream, encoded, dim, degree):
    count, width = encoded.shape
    grad_points = encoded.new_empty(count, dim)
    _lib.call("ngp_freq_encode_backward", _lib.ptr(upstream), _lib.ptr(encoded), count, dim, degree, width, _lib.ptr(grad_points),
              _lib.stream())
    return grad_points


class _FreqEncode(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        points = (inputs if inputs.is_cuda else inputs.cuda()).float().contiguous()
        encoded = _launch_forward(points, int(degree), int(output_dim))
        ctx.save_for_backward(encoded)
        ctx.shape_info = (points.shape[1], int(degree))
        return encoded

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        (encoded,) = ctx.saved_tensors
        dim, degree = ctx.shape_info
        return _launch_backward(grad.float().contiguous(), encoded, dim, degree), None, None


def freq_encode(inputs, degree, output_dim):
    return _FreqEncode.apply(inputs, degree, output_dim)


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim * (1 + 2 * degree)

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        lead = inputs.shape[:-1]
        flat = freq_encode(inputs.reshape(-1, self.input_dim), self.degree, self.output_dim)
        return flat.reshape(*lead, self.output_dim)
