"""GPU tests: the fused field (csrc/field.cu) against the op-by-op network path (GridEncoder -> MLP -> trunc_exp -> SHEncoder
-> MLP -> activation), whose operators are individually verified against the reference kernels."""
import pytest
import torch

from raw_ngp_b200 import synthetic
from raw_ngp_b200.nerf import NeRFNetwork, default_opt

pytestmark = pytest.mark.gpu


def _net(**kw):
    torch.manual_seed(0)
    m = NeRFNetwork(default_opt(**kw)).cuda()
    m.grid_encoder.embeddings.data = (torch.rand_like(m.grid_encoder.embeddings.data) * 2 - 1).half()
    for mlp in (m.grid_mlp, m.view_mlp):
        for l in mlp.net:
            l.weight.data.mul_(1.5)
    return m


CASES = [
    dict(bound=1),
    dict(bound=2, rfield=True),
    dict(bound=1, density_activation="softplus", color_activation="sigmoid"),
    dict(bound=1, color_activation="exp", pose_opt="barf", start_annealing=0.0, end_annealing=0.5),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
@pytest.mark.parametrize("M", [1000, 50003])
def test_fused_field_matches_unfused(case, M):
    model = _net(hashmap_size=16, hashgrid_resolution=512, **case)
    model.annealing = 0.2
    b = model.bound
    x = synthetic.uniform_points(M, seed=1, lo=-b * 1.02, hi=b * 1.02).cuda()      # a few points outside the box
    d = (synthetic.unit_vectors(M, seed=2) * 1.7).cuda()
    ld = synthetic.unit_vectors(M, seed=3).cuda() if case.get("rfield") else None
    gs = (torch.randn(M, generator=torch.Generator().manual_seed(4)) * 0.1).cuda()
    gc = (torch.randn(M, 3, generator=torch.Generator().manual_seed(5)) * 0.1).cuda()
    params = [model.grid_encoder.embeddings] + [l.weight for l in model.grid_mlp.net] + [l.weight for l in model.view_mlp.net]

    def run(fused):
        model.FUSED = fused
        for p in params:
            p.grad = None
        with torch.amp.autocast("cuda"):
            dn = d / torch.norm(d, dim=-1, keepdim=True)               # renderer.py:544
            out = model(x, dn, ld)
        sigma, color = out["sigma"].float(), out["color"].float()
        (sigma * gs).sum().backward(retain_graph=True)
        (color * gc).sum().backward()
        return sigma.detach(), color.detach(), [p.grad.float().clone() for p in params]

    s0, c0, g0 = run(False)
    s1, c1, g1 = run(True)
    assert torch.isfinite(s1).all() and torch.isfinite(c1).all()
    # forward: same kernels' arithmetic, same rounding points -> tight
    torch.testing.assert_close(s1, s0, rtol=2e-3, atol=1e-4)
    torch.testing.assert_close(c1, c0, rtol=2e-3, atol=1e-4)
    for name, a, bb in zip(["table"] + ["w"] * 6, g1, g0):
        scale = bb.abs().max().clamp(min=1e-8)
        err = (a - bb).abs() / scale
        assert err.max().item() < 3e-2 and err.mean().item() < 1e-3, (name, err.max().item(), err.mean().item())


def test_density_only_matches_forward():
    model = _net(bound=1, hashmap_size=16, hashgrid_resolution=512)
    x = synthetic.uniform_points(20000, seed=7).cuda()
    with torch.no_grad(), torch.amp.autocast("cuda"):
        model.FUSED = True
        a = model.density(x)["sigma"]
        model.FUSED = False
        b = model.density(x)["sigma"]
    torch.testing.assert_close(a.float(), b.float(), rtol=2e-3, atol=1e-5)


def test_fused_field_with_grad_sink_and_inference():
    model = _net(bound=1, hashmap_size=16, hashgrid_resolution=512)
    enc = model.grid_encoder
    enc.grad_sink = torch.zeros_like(enc.embeddings.data)
    x = synthetic.uniform_points(4096, seed=1).cuda()
    d = synthetic.unit_vectors(4096, seed=2).cuda()
    with torch.amp.autocast("cuda"):
        out = model(x, d)
    (out["sigma"].sum() + out["color"].sum()).backward()
    assert enc.embeddings.grad is None and enc.grad_sink.float().abs().sum().item() > 0
    with torch.no_grad(), torch.amp.autocast("cuda"):
        out2 = model(x, d)
    torch.testing.assert_close(out2["color"], out["color"].detach())
