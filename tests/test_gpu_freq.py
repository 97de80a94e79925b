"""GPU tests: frequency encoder (csrc/freq_encode.cu) against the reference kernel (oracle/_ref/_freqencoder, unmodified
source) -- bit-exact, both sides evaluate the same __sinf -- and against the CPU oracle."""
import numpy as np
import pytest
import torch

from raw_ngp_b200.encoding import get_encoder
from raw_ngp_b200.freqencoder import FreqEncoder, freq_encode

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_cuda
    try:
        ref_cuda.module("_freqencoder")
    except RuntimeError as e:
        pytest.skip(str(e))
    return ref_cuda


@pytest.mark.parametrize("B,D,deg", [(0, 3, 4), (1, 3, 6), (4099, 3, 10), (1000, 2, 1), (777, 5, 3), (262144, 3, 6)])
def test_freq_forward_backward_match_reference(B, D, deg):
    ref_cuda = _ref()
    g = torch.Generator().manual_seed(B + deg)
    x = (torch.rand(B, D, generator=g) * 4 - 2).cuda()
    C = D + 2 * D * deg
    out = freq_encode(x, deg, C)
    assert out.shape == (B, C)
    if B == 0:
        return
    ref = ref_cuda.freq_forward(x, deg)
    assert torch.equal(out, ref)
    grad = torch.randn(B, C, generator=g).cuda()
    xg = x.clone().requires_grad_(True)
    freq_encode(xg, deg, C).backward(grad)
    assert torch.equal(xg.grad, ref_cuda.freq_backward(grad, ref, D, deg))


def test_freq_matches_oracle_and_autograd():
    from oracle import freq_oracle
    torch.manual_seed(0)
    x = (torch.rand(2048, 3) * 2 - 1).cuda()
    enc, dim = get_encoder("frequency", input_dim=3, multires=6)
    assert isinstance(enc, FreqEncoder) and dim == 39 and enc.output_dim == 39
    out = enc(x.reshape(32, 64, 3))
    assert out.shape == (32, 64, 39)
    o = freq_oracle.forward(x.cpu().numpy(), 6)
    # __sinf vs the correctly rounded sine: absolute error grows with the argument (|x 2^5| <= 32 here)
    np.testing.assert_allclose(out.reshape(-1, 39).cpu().numpy(), o, rtol=0, atol=2e-5)
    grad = torch.randn(2048, 39).cuda()
    xg = x.clone().requires_grad_(True)
    enc(xg).backward(grad)
    np.testing.assert_allclose(xg.grad.cpu().numpy(), freq_oracle.backward(grad.cpu().numpy(), enc(x).cpu().numpy(), 3, 6), rtol=1e-5, atol=1e-5)
    # derivative check against torch autograd of the same formula in fp64
    xd = x.double().requires_grad_(True)
    cols = [xd]
    for f in range(6):
        cols += [torch.sin(xd * 2 ** f), torch.cos(xd * 2 ** f)]
    torch.cat(cols, dim=-1).backward(grad.double())
    torch.testing.assert_close(xg.grad.double(), xd.grad, rtol=1e-3, atol=2e-3)
