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


def test_ws_forward_with_light_stage_widths_matches_kernel_pairs():
    """The warp-specialised forward on 16-column panels (view_mlp 47 -> 80 -> 80 -> 3, 48-wide input with SH of the light
    direction) against the density-field + view-MLP kernel pair it replaces: sigma / rgb and every saved tensor (plain rows
    for the pair backward), ragged M, and through the autograd Function (forward in the new kernel, backward in the pairs)
    against the op-by-op network."""
    import ctypes
    from raw_ngp_b200 import _lib, field as F_
    from raw_ngp_b200.ffmlp import _pad16, _ptr_array
    torch.manual_seed(0)
    model = NeRFNetwork(default_opt(bound=2, contract=True, rfield=True, hashmap_size=15, hashgrid_resolution=256, grid_size=32)).cuda()
    model.grid_encoder.embeddings.data = model.grid_encoder.embeddings.data.uniform_(-0.5, 0.5).half()
    enc = model.grid_encoder
    M = 128 * 37 + 45
    g = torch.Generator().manual_seed(4)
    xyzs = ((torch.rand(M, 3, generator=g) * 2 - 1) * 2).cuda()
    dirs = (torch.randn(M, 3, generator=g) * 1.7).cuda()
    ldirs = torch.randn(M, 3, generator=g).cuda()
    gw, vw = [l.weight for l in model.grid_mlp.net], [l.weight for l in model.view_mlp.net]
    p1 = [_pad16(d) for d in [gw[0].shape[1]] + [w.shape[0] for w in gw]]
    p2 = [_pad16(d) for d in [vw[0].shape[1]] + [w.shape[0] for w in vw]]
    assert p2 == [48, 80, 80, 16] and not F_._ws_ok(p1, p2) and F_._ws_fwd_ok(p1, p2, enc.num_levels)
    w1 = [F_._pad_weight(w, p1[i + 1], p1[i]) for i, w in enumerate(gw)]
    w2 = [F_._pad_weight(w, p2[i + 1], p2[i]) for i, w in enumerate(vw)]
    S, H, L, gt, ac, ip = F_._grid_scalars(enc)
    c1, c2 = (ctypes.c_uint32 * 4)(*p1), (ctypes.c_uint32 * 4)(*p2)
    P, st = _lib.ptr, _lib.stream()
    f16 = dict(dtype=torch.float16, device="cuda")

    def bufs():
        return dict(enc=torch.zeros(M, p1[0], **f16), a1=[torch.zeros(M, p1[l + 1], **f16) for l in range(2)], in2=torch.zeros(M, p2[0], **f16),
                    a2=[torch.zeros(M, p2[l + 1], **f16) for l in range(2)], sigma=torch.zeros(M, device="cuda"), rgb=torch.zeros(M, 3, device="cuda"))
    a, b = bufs(), bufs()
    _lib.call("ngp_field_forward_full", P(xyzs), P(dirs), P(ldirs), P(enc.embeddings), P(enc.offsets), None, 2.0, S, H, L, gt, ac, ip,
              _ptr_array(w1), c1, _ptr_array(w2), c2, M, None, 0, 1.0, 3, P(a["enc"]), _ptr_array(a["a1"]), P(a["in2"]), _ptr_array(a["a2"]),
              P(a["sigma"]), P(a["rgb"]), None, st)
    _lib.call("ngp_field_forward_density", P(xyzs), P(dirs), P(ldirs), P(enc.embeddings), P(enc.offsets), None, 2.0, S, H, L, gt, ac, ip,
              _ptr_array(w1), c1, 3, M, None, 0, 1.0, P(b["enc"]), _ptr_array(b["a1"]), P(b["sigma"]), P(b["in2"]), p2[0], st)
    _lib.call("ngp_mlp_forward_rgb", P(b["in2"]), p2[0], _ptr_array(w2), c2, 3, M, None, 1, 3, P(b["rgb"]), _ptr_array(b["a2"]), st)
    torch.cuda.synchronize()
    assert torch.equal(a["enc"], b["enc"])
    for x, y in zip(a["a1"] + [a["in2"]] + a["a2"], b["a1"] + [b["in2"]] + b["a2"]):
        torch.testing.assert_close(x.float(), y.float(), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(a["sigma"], b["sigma"], rtol=2e-3, atol=1e-5)
    torch.testing.assert_close(a["rgb"], b["rgb"], rtol=3e-3, atol=1e-5)
    # through autograd: forward in the warp-specialised kernel, backward in the kernel pairs, vs the op-by-op network
    model.train()
    with torch.autocast("cuda", dtype=torch.float16):
        out = model(xyzs, dirs, ldirs)
        loss = (out["sigma"].float().clamp(max=50).mean() + out["color"].float().mean())
    loss.backward()
    g_fused = [p.grad.clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    model.FUSED = False
    with torch.autocast("cuda", dtype=torch.float16):
        out2 = model(xyzs, dirs, ldirs)
        loss2 = (out2["sigma"].float().clamp(max=50).mean() + out2["color"].float().mean())
    loss2.backward()
    torch.testing.assert_close(loss, loss2, rtol=2e-3, atol=1e-5)
    for ga, gb in zip(g_fused, [p.grad for p in model.parameters()]):
        scale = gb.float().abs().max().clamp(min=1e-8)
        err = (ga.float() - gb.float()).abs() / scale
        assert err.max().item() < 3e-2 and err.mean().item() < 2e-3
