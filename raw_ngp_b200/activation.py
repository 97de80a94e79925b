"""trunc_exp: the density activation of the reference (activation.py:9-21).

Forward y = exp(x) evaluated in fp32 regardless of autocast; the backward multiplies the incoming gradient by exp of the input
clamped to [-80, 80], so an overflowing activation cannot turn a zero upstream gradient into inf * 0 = nan."""
import torch


class _TruncExp(torch.autograd.Function):
    LIMIT = 80.0

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, logits):
        ctx.save_for_backward(logits)
        return logits.exp()

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        (logits,) = ctx.saved_tensors
        return grad_out * logits.clamp(min=-_TruncExp.LIMIT, max=_TruncExp.LIMIT).exp()


def trunc_exp(x):
    return _TruncExp.apply(x)
