"""GPU tests: ngp_fused_adam / ngp_check_finite (csrc/optim.cu) against torch.optim.Adam under GradScaler semantics
(main.py:245, nerf/train_utils.py:897-904)."""
import pytest
import torch

from raw_ngp_b200 import _lib
from raw_ngp_b200.trainer import FusedAdam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [4096, 100003])
@pytest.mark.parametrize("gdtype", [torch.float16, torch.float32])
def test_fused_adam_matches_torch(n, gdtype):
    torch.manual_seed(0)
    dev = "cuda"
    master = torch.randn(n, device=dev)
    ref = torch.nn.Parameter(master.clone())
    opt_ref = torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    grad = torch.zeros(n, device=dev, dtype=gdtype)
    lp = master.half()
    opt = FusedAdam(lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    opt.add_group(master, grad, lp)
    scale = 128.0
    inv_scale = torch.full((1,), 1.0 / scale, device=dev)
    found_inf = torch.zeros(1, device=dev)
    for it in range(5):
        g = torch.randn(n, device=dev) * 0.1
        grad.copy_((g * scale).to(gdtype))
        ref.grad = grad.float() / scale
        found_inf.zero_()
        _lib.call("ngp_check_finite", _lib.ptr(grad), _lib.dtype_id(gdtype), n, _lib.ptr(found_inf), _lib.stream())
        assert found_inf.item() == 0.0
        opt.step(inv_scale, found_inf, zero_grad=True)
        opt_ref.step()
        assert grad.abs().max().item() == 0.0                       # gradient cleared in the same pass
        torch.testing.assert_close(master, ref.data, rtol=1e-5, atol=1e-6)
        assert torch.equal(lp, master.half())                        # low-precision copy written in the same pass


@pytest.mark.parametrize("pos", [0, 77777, 100002])
def test_inf_skips_the_step(pos):
    n, dev = 100003, "cuda"
    master = torch.randn(n, device=dev)
    before = master.clone()
    grad = (torch.randn(n, device=dev) * 0.1).half()
    grad[pos] = float("inf")
    opt = FusedAdam()
    opt.add_group(master, grad, None)
    found_inf = torch.zeros(1, device=dev)
    _lib.call("ngp_check_finite", _lib.ptr(grad), _lib.NGP_F16, n, _lib.ptr(found_inf), _lib.stream())
    assert found_inf.item() == 1.0
    opt.step(torch.ones(1, device=dev), found_inf, zero_grad=True)
    assert torch.equal(master, before)                                # parameters, m, v untouched
    assert opt.groups[0]["m"].abs().max().item() == 0.0
    assert grad.float().abs().max().item() == 0.0                     # but the gradient buffer is cleared for the next step
