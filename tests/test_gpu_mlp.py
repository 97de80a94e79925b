"""GPU numerics: fused tcgen05 MLP vs a plain PyTorch fp32 evaluation of the same layers on the same fp16-rounded
inputs and weights (tolerance: fp16 activations between layers, rel 1e-2 of the tensor scale)."""
import pytest
import torch

from raw_ngp_b200.ffmlp import fused_mlp

pytestmark = pytest.mark.gpu


def _ref(x, ws):
    h = x.float()
    for i, w in enumerate(ws):
        h = h @ w.half().float().T
        if i + 1 < len(ws):
            h = torch.relu(h).half().float()
    return h


@pytest.mark.parametrize("dims", [(32, 64, 64, 16), (31, 64, 64, 3), (47, 80, 80, 3), (32, 64, 16), (32, 48, 48, 48, 16)],
                         ids=lambda d: "x".join(map(str, d)))
@pytest.mark.parametrize("M", [128, 1000, 70001])
def test_fused_mlp_forward_backward(dims, M):
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(M, dims[0], generator=g) * 0.5).half().cuda().requires_grad_(True)
    ws = [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).cuda().requires_grad_(True) for i in range(len(dims) - 1)]
    y = fused_mlp(x, *ws)
    assert y.shape == (M, dims[-1]) and y.dtype == torch.float16
    xr = x.detach().clone().requires_grad_(True)
    wr = [w.detach().clone().requires_grad_(True) for w in ws]
    yr = _ref(xr, wr)
    scale = yr.abs().max()
    assert ((y.float() - yr).abs().max() / scale).item() < 1e-2
    dy = (torch.randn(M, dims[-1], generator=g) * 0.1).half().cuda()
    y.backward(dy)
    yr.backward(dy.float())
    for a, b in zip(ws, wr):
        s = b.grad.abs().max()
        assert ((a.grad - b.grad).abs().max() / s).item() < 2e-2, "weight gradient"
    # a hidden pre-activation within fp16 rounding of zero can flip its ReLU mask between the two evaluations, which
    # changes that sample's input gradient discretely: bound the bulk tightly and the rare outliers loosely
    err = ((x.grad.float() - xr.grad.float()).abs() / xr.grad.float().abs().max()).flatten()
    assert err.mean().item() < 2e-3, "input gradient (mean)"
    k = max(1, int(err.numel() * 1e-4))
    assert torch.topk(err, k).values[-1].item() < 2e-2 and err.max().item() < 0.3, "input gradient (tail)"


def test_fused_mlp_matches_network_mlp_under_autocast():
    from raw_ngp_b200.nerf import MLP, default_opt
    torch.manual_seed(0)
    mlp = MLP(32, 16, 64, 3, default_opt(), bias=False).cuda()
    x = torch.randn(5000, 32, device="cuda").half()
    with torch.amp.autocast("cuda"):
        ref = mlp(x)
    out = fused_mlp(x, *[l.weight for l in mlp.net])
    assert ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item() < 1e-2
