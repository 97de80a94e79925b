"""GPU parity: SH encoder vs the reference's compiled extension (oracle/_ref/_shencoder)."""
import pytest
import torch

from raw_ngp_b200 import synthetic
from raw_ngp_b200.shencoder import SHEncoder, sh_encode
from raw_ngp_b200.shencoder import sphere_harmonics as sh_mod

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("degree", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_forward_and_jacobian(degree, ref_sh):
    from oracle import ref_cuda
    d = synthetic.unit_vectors(10007, seed=3).cuda()
    d[:3] = torch.eye(3).cuda()          # axis-aligned directions
    d[3:6] = -torch.eye(3).cuda()
    ours, ours_j = sh_mod.sh_encode_with_jacobian(d, degree)
    ref, ref_j = ref_cuda.sh_forward(d, degree, True)
    # same polynomials, different (factored) evaluation order: fp32 cancellation noise grows with the degree
    torch.testing.assert_close(ours, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ours_j, ref_j, rtol=1e-5, atol=1e-4)
    plain = sh_encode(d, degree, False)
    assert torch.equal(plain, ours)


@pytest.mark.parametrize("degree", [1, 4, 8])
def test_sh_backward(degree, ref_sh):
    from oracle import ref_cuda
    d = (synthetic.unit_vectors(4099, seed=4) * 0.9).cuda()   # not exactly unit: polynomials, not just the sphere
    grad = torch.randn(4099, degree ** 2, generator=torch.Generator().manual_seed(7)).cuda()
    _, ref_j = ref_cuda.sh_forward(d, degree, True)
    ref_g = ref_cuda.sh_backward(grad, d, degree, ref_j)
    x = d.clone().requires_grad_(True)
    sh_encode(x, degree, True).backward(grad)
    scale = ref_g.abs().max().clamp(min=1.0)
    torch.testing.assert_close(x.grad / scale, ref_g / scale, rtol=0, atol=1e-5)


def test_sh_module_matches_reference_pipeline(ref_sh):
    from oracle import ref_cuda
    enc = SHEncoder(degree=4)
    v = torch.randn(2 ** 16, 3, generator=torch.Generator().manual_seed(9)).cuda() * 3
    out = enc(v)
    n = v / torch.norm(v, dim=-1, keepdim=True)
    ref, _ = ref_cuda.sh_forward(n.contiguous(), 4, False)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=2e-6)
    # orthonormality on the sphere (size-independent property): mean of Y_i*Y_j*4pi ~ delta_ij
    gram = (out.double().T @ out.double()) * (4 * 3.141592653589793 / out.shape[0])
    assert (gram - torch.eye(16, dtype=torch.float64, device=gram.device)).abs().max().item() < 0.05
