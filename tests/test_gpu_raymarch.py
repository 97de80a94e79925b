"""GPU parity: ray marching / compositing operators vs the reference's compiled extension
(oracle/_ref/_raymarching_mob).  Integer outputs (morton3D, packbits, per-ray sample counts, alive ids) must be
bit-exact; sample buffers are compared per ray segment (the reference's offsets depend on atomic arrival order,
ours are the ray-ordered prefix sum); composited floats within rel 1e-5 (+ abs floor)."""
import pytest
import torch

from raw_ngp_b200 import raymarching, synthetic

pytestmark = pytest.mark.gpu


def _scene(H=128, cascade=1, bound=1.0, radius=0.5):
    grid = synthetic.ball_density_grid(H=H, cascade=cascade, bound=bound, radius=radius).cuda()
    thresh = min(grid.clamp(min=0).mean().item(), 10.0)
    return grid, thresh


def _rays(N, seed=2, bound=1.0, min_near=0.05):
    o, d = synthetic.sphere_rays(N, seed=seed)
    o, d = o.cuda(), d.cuda()
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device="cuda")
    nears, fars = synthetic.near_far_torch(o, d, aabb, min_near)
    return o, d, aabb, nears, fars


def test_morton_packbits_bit_exact(ref_march):
    from oracle import ref_cuda
    g = torch.Generator().manual_seed(11)
    coords = torch.randint(0, 1024, (100003, 3), generator=g, dtype=torch.int32).cuda()
    ours = raymarching.morton3D(coords)
    assert torch.equal(ours, ref_cuda.morton3D(coords))
    assert torch.equal(raymarching.morton3D_invert(ours), ref_cuda.morton3D_invert(ours))
    assert torch.equal(raymarching.morton3D_invert(ours), coords)           # round trip
    idx = torch.randint(0, 2 ** 30, (50000,), generator=g, dtype=torch.int32).cuda()
    assert torch.equal(raymarching.morton3D_invert(idx), ref_cuda.morton3D_invert(idx))
    assert raymarching.morton3D(torch.empty(0, 3, dtype=torch.int32, device="cuda")).numel() == 0

    grid = torch.randn(2, 128 ** 3, generator=g).cuda()
    grid[0, :64] = -1.0
    for th in (0.0, 0.37, -2.0, 10.0):
        assert torch.equal(raymarching.packbits(grid, th), ref_cuda.packbits(grid, th))
    # device-side threshold = min(mean, density_thresh)
    mean = grid.clamp(min=0).mean().reshape(1)
    assert torch.equal(raymarching.packbits(grid, (mean, 10.0)), ref_cuda.packbits(grid, min(mean.item(), 10.0)))
    # preallocated bitfield is written in place and returned
    bf = torch.zeros(2 * 128 ** 3 // 8, dtype=torch.uint8, device="cuda")
    assert raymarching.packbits(grid, 0.1, bf).data_ptr() == bf.data_ptr()


def test_near_far_sph_flatten(ref_march):
    from oracle import ref_cuda
    o, d, aabb, _, _ = _rays(50000)
    d[:100] *= 3.7                       # un-normalised directions
    d[100:110, 0] = 0.0                  # axis-parallel rays (1/0 = inf path)
    o[110:120] = o[110:120] * 0.1        # origins inside the box
    aabb2 = torch.tensor([-0.7, -0.5, -1.0, 0.6, 0.9, 0.3], device="cuda")
    for bb in (aabb, aabb2):
        n1, f1 = raymarching.near_far_from_aabb(o, d, bb, 0.05)
        n2, f2 = ref_cuda.near_far_from_aabb(o, d, bb, 0.05)
        assert torch.equal(n1, n2) and torch.equal(f1, f2)
    inside = o * 0.2
    c1 = raymarching.sph_from_ray(inside, d, 1.5)
    c2 = ref_cuda.sph_from_ray(inside, d, 1.5)
    torch.testing.assert_close(c1, c2, rtol=1e-5, atol=1e-6)


MARCH_CASES = [
    dict(name="cfg2", N=4096, H=128, cascade=1, bound=1.0, contract=False, dt_gamma=0.0, perturb=True, ldir=False),
    dict(name="cfg2-noperturb", N=4096, H=128, cascade=1, bound=1.0, contract=False, dt_gamma=0.0, perturb=False, ldir=False),
    dict(name="cone", N=4096, H=128, cascade=1, bound=1.0, contract=False, dt_gamma=1 / 128, perturb=True, ldir=False),
    dict(name="cascade3", N=8192, H=128, cascade=3, bound=4.0, contract=False, dt_gamma=0.0, perturb=True, ldir=True),
    dict(name="cfg3-contract-C4", N=8192, H=128, cascade=4, bound=8.0, contract=True, dt_gamma=0.0, perturb=True, ldir=True),
    dict(name="cfg3-contract-C2", N=8192, H=128, cascade=2, bound=2.0, contract=True, dt_gamma=0.0, perturb=True, ldir=True),
    dict(name="H64-odd-bound", N=3001, H=64, cascade=2, bound=1.5, contract=False, dt_gamma=0.0, perturb=True, ldir=False),
    dict(name="maxsteps64", N=2048, H=128, cascade=1, bound=1.0, contract=False, dt_gamma=0.0, perturb=True, ldir=False, max_steps=64),
    # lattice far longer than one 1024-point window of the cooperative marcher (8x the cube diagonal at dt_min)
    dict(name="long-lattice-bound8", N=4096, H=128, cascade=4, bound=8.0, contract=False, dt_gamma=0.0, perturb=True, ldir=False),
    dict(name="long-lattice-sparse", N=4096, H=128, cascade=4, bound=8.0, contract=False, dt_gamma=0.0, perturb=True, ldir=False, radius=0.05),
]


def _march_both(case):
    from oracle import ref_cuda
    grid, thresh = _scene(H=case["H"], cascade=case["cascade"], bound=case["bound"], radius=case.get("radius", 0.5))
    bitfield = raymarching.packbits(grid, thresh)
    o, d, aabb, nears, fars = _rays(case["N"], bound=case["bound"])
    ldir = synthetic.unit_vectors(case["N"], seed=3).cuda() if case["ldir"] else None
    max_steps = case.get("max_steps", 1024)
    torch.manual_seed(1234)
    ours = raymarching.march_rays_train(o, d, ldir, case["bound"], case["contract"], bitfield, case["cascade"], case["H"],
                                        nears, fars, case["perturb"], case["dt_gamma"], max_steps)
    torch.manual_seed(1234)
    N = o.shape[0]
    noises = torch.rand(N, device="cuda") if case["perturb"] else torch.zeros(N, device="cuda")
    ref = ref_cuda.march_rays_train(o, d, ldir, case["bound"], case["contract"], bitfield, case["cascade"], case["H"],
                                    nears.view(-1).contiguous(), fars.view(-1).contiguous(), noises, case["dt_gamma"], max_steps)
    return ours, ref


@pytest.mark.parametrize("coop", [True, False], ids=["warp-cooperative", "thread-per-ray"])
@pytest.mark.parametrize("case", MARCH_CASES, ids=lambda c: c["name"])
def test_march_rays_train_bit_exact(case, coop, ref_march, monkeypatch):
    from raw_ngp_b200.raymarching import raymarching as rm_mod
    monkeypatch.setattr(rm_mod, "COOPERATIVE_MARCH", coop)
    (xyzs, dirs, ts, rays, ldirs), (rx, rd, rt, rrays, rl) = _march_both(case)
    N = rays.shape[0]
    # counts and total bit-exact
    assert torch.equal(rays[:, 1], rrays[:, 1]), f"{(rays[:, 1] != rrays[:, 1]).sum().item()} rays differ in sample count"
    M = int(rays[:, 1].sum().item())
    assert xyzs.shape[0] == M == rx.shape[0] and M > 0
    # our offsets are the ray-ordered exclusive scan
    excl = torch.cumsum(rays[:, 1], 0) - rays[:, 1]
    assert torch.equal(rays[:, 0].long(), excl.long())
    # the reference's offsets are a permutation of segments: gather its samples into ray order and compare bitwise
    counts = rays[:, 1].long()
    ray_of_sample = torch.repeat_interleave(torch.arange(N, device="cuda"), counts)
    within = torch.arange(M, device="cuda") - excl.long()[ray_of_sample]
    ref_pos = rrays[:, 0].long()[ray_of_sample] + within
    assert torch.equal(xyzs, rx[ref_pos])
    assert torch.equal(dirs, rd[ref_pos])
    assert torch.equal(ts, rt[ref_pos])
    if case["ldir"]:
        assert torch.equal(ldirs, rl[ref_pos])
    else:
        assert ldirs is None and rl is None
    assert torch.equal(raymarching.flatten_rays(rays, M), ray_of_sample.int())


def test_march_rays_train_backward_segment_sums(ref_march):
    case = MARCH_CASES[0]
    grid, thresh = _scene()
    bitfield = raymarching.packbits(grid, thresh)
    o, d, aabb, nears, fars = _rays(2048)
    o.requires_grad_(True)
    d.requires_grad_(True)
    xyzs, dirs, ts, rays, _ = raymarching.march_rays_train(o, d, None, 1.0, False, bitfield, 1, 128, nears.detach(), fars.detach(), False, 0.0, 1024)
    g = torch.Generator().manual_seed(5)
    gx = torch.randn(xyzs.shape, generator=g).cuda()
    gd = torch.randn(dirs.shape, generator=g).cuda()
    (xyzs * gx).sum().backward(retain_graph=True)
    go1 = o.grad.clone(); gd1 = d.grad.clone()
    o.grad = None; d.grad = None
    ((xyzs * gx).sum() + (dirs * gd).sum()).backward()
    # reference semantics (raymarching.py:319-329): segment sums of dL/dxyz and dL/dxyz * t (+ dL/ddirs)
    seg = torch.repeat_interleave(torch.arange(2048, device="cuda"), rays[:, 1].long())
    exp_o = torch.zeros(2048, 3, device="cuda", dtype=torch.float64).index_add_(0, seg, gx.double())
    exp_d = torch.zeros(2048, 3, device="cuda", dtype=torch.float64).index_add_(0, seg, (gx * ts[:, :1]).double())
    exp_d2 = exp_d.clone().index_add_(0, seg, gd.double())
    torch.testing.assert_close(go1.double(), exp_o, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(gd1.double(), exp_d, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(o.grad.double(), exp_o, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(d.grad.double(), exp_d2, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("T_thresh", [1e-8, 1e-4, 0.3])
def test_composite_rays_train_forward_backward(T_thresh, ref_march):
    from oracle import ref_cuda
    (xyzs, dirs, ts, rays, _), _ = _march_both(MARCH_CASES[0])
    M, N = xyzs.shape[0], rays.shape[0]
    g = torch.Generator().manual_seed(21)
    sigmas = (torch.rand(M, generator=g) * 30).cuda()
    sigmas[::7] = 0.0
    rgbs = torch.rand(M, 3, generator=g).cuda()
    rays2 = rays.clone()
    rays2[5, 1] = 0                       # empty ray
    rays2[9, 0] = M - 3; rays2[9, 1] = 50  # offset + count > M  -> zero output (raymarching.cu:540-547)

    s1 = sigmas.clone().requires_grad_(True); c1 = rgbs.clone().requires_grad_(True)
    w, ws, dp, im = raymarching.composite_rays_train(s1, c1, ts, rays2, T_thresh)
    rw, rws, rdp, rim = ref_cuda.composite_rays_train_forward(sigmas, rgbs, ts, rays2, T_thresh)
    torch.testing.assert_close(w, rw, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ws, rws, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dp, rdp, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(im, rim, rtol=1e-5, atol=1e-6)
    assert (w == 0).sum().item() >= (rw == 0).sum().item() - 2 and abs((w == 0).sum().item() - (rw == 0).sum().item()) <= 2

    gw = torch.randn(M, generator=g).cuda() * 0.1
    gws = torch.randn(N, generator=g).cuda()
    gdp = torch.randn(N, generator=g).cuda()
    gim = torch.randn(N, 3, generator=g).cuda()
    torch.autograd.backward([w, ws, dp, im], [gw, gws, gdp, gim])
    rgs, rgc = ref_cuda.composite_rays_train_backward(gw, gws, gdp, gim, sigmas, rgbs, ts, rays2, rws, rdp, rim, T_thresh)
    torch.testing.assert_close(c1.grad, rgc, rtol=1e-5, atol=1e-7)
    scale = rgs.abs().max()
    torch.testing.assert_close(s1.grad / scale, rgs / scale, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("contract,cascade,bound", [(False, 1, 1.0), (True, 2, 2.0), (False, 3, 4.0)])
def test_inference_loop_matches_reference(contract, cascade, bound, ref_march):
    """Runs the renderer's march_rays / composite_rays loop (nerf/renderer.py:588-616) with both backends in lock-step
    on an analytic density/colour field."""
    from oracle import ref_cuda
    grid, thresh = _scene(cascade=cascade, bound=bound)
    bitfield = raymarching.packbits(grid, thresh)
    N = 20000
    o, d, aabb, nears, fars = _rays(N, bound=bound)
    nears = nears.view(-1).contiguous(); fars = fars.view(-1).contiguous()

    def field(x):
        sig = 40.0 * torch.exp(-4 * (x ** 2).sum(-1))
        col = torch.sigmoid(3 * x)
        return sig.contiguous(), col.contiguous()

    state = []
    for _ in range(2):
        state.append(dict(ws=torch.zeros(N, device="cuda"), dp=torch.zeros(N, device="cuda"), im=torch.zeros(N, 3, device="cuda"),
                          alive=torch.arange(N, dtype=torch.int32, device="cuda"), t=nears.clone()))
    a, b = state
    step, T_thresh = 0, 1e-4
    while step < 1024:
        n_alive = a["alive"].shape[0]
        assert n_alive == b["alive"].shape[0]
        if n_alive == 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        noises = torch.zeros(n_alive, device="cuda")
        x1, d1, t1 = raymarching.march_rays(n_alive, n_step, a["alive"], a["t"], o, d, bound, contract, bitfield, cascade, 128, nears, fars, False, 0.0, 1024)
        x2, d2, t2 = ref_cuda.march_rays(n_alive, n_step, b["alive"], b["t"], o, d, bound, contract, bitfield, cascade, 128, nears, fars, noises, 0.0, 1024)
        assert torch.equal(x1, x2) and torch.equal(d1, d2) and torch.equal(t1, t2)
        s, c = field(x1)
        raymarching.composite_rays(n_alive, n_step, a["alive"], a["t"], s, c, t1, a["ws"], a["dp"], a["im"], T_thresh)
        ref_cuda.composite_rays(n_alive, n_step, b["alive"], b["t"], s, c, t2, b["ws"], b["dp"], b["im"], T_thresh)
        assert torch.equal(a["alive"], b["alive"])
        comp, cnt = raymarching.compact_rays_alive(a["alive"])
        expect = a["alive"][a["alive"] >= 0]
        assert cnt.item() == expect.shape[0] and torch.equal(comp[:cnt.item()], expect)
        a["alive"] = expect
        b["alive"] = b["alive"][b["alive"] >= 0]
        step += n_step
    for k in ("ws", "dp", "im", "t"):
        assert torch.equal(a[k], b[k]), k
    assert a["ws"].max().item() > 0.5


def test_march_count_at_scale_and_properties(ref_march):
    """>= 10^5 rays (full-size property test): counts bit-exact vs the reference, every sample inside the box, t
    strictly increasing within a ray, sample count <= max_steps."""
    case = dict(name="big", N=200000, H=128, cascade=1, bound=1.0, contract=False, dt_gamma=0.0, perturb=True, ldir=False)
    (xyzs, dirs, ts, rays, _), (rx, rd, rt, rrays, _) = _march_both(case)
    assert torch.equal(rays[:, 1], rrays[:, 1])
    assert rays[:, 1].max().item() <= 1024
    assert xyzs.abs().max().item() <= 1.0
    seg = torch.repeat_interleave(torch.arange(rays.shape[0], device="cuda"), rays[:, 1].long())
    same = seg[1:] == seg[:-1]
    assert (ts[1:, 0][same] > ts[:-1, 0][same]).all()


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 70001, 2073600])
def test_compact_rays_alive_sizes(n):
    """single-block and two-pass multi-block compaction against the boolean mask of renderer.py:612, ragged sizes"""
    g = torch.Generator().manual_seed(n)
    alive = torch.arange(n, dtype=torch.int32)
    dead = torch.rand(n, generator=g) < (0.85 if n > 4096 else 0.4)
    alive[dead] = -1
    alive = alive.cuda()
    out, cnt = raymarching.compact_rays_alive(alive)
    expect = alive[alive >= 0]
    assert cnt.item() == expect.shape[0]
    assert torch.equal(out[:cnt.item()], expect)


@pytest.mark.parametrize("n_step", [1, 3, 8, 19, 32])
def test_march_rays_step_counts_match_reference(n_step, ref_march):
    """one inference march call: staged (shared memory, coalesced stores; n_step <= 19) and direct output paths, ragged
    last warp, against the reference kernel"""
    from oracle import ref_cuda
    grid, thresh = _scene(cascade=2, bound=2.0)
    bitfield = raymarching.packbits(grid, thresh)
    N = 5003
    o, d, aabb, nears, fars = _rays(N, bound=2.0)
    nears = nears.view(-1).contiguous(); fars = fars.view(-1).contiguous()
    alive = torch.arange(N, dtype=torch.int32, device="cuda")[torch.randperm(N, device="cuda")][:4001].contiguous()
    n_alive = alive.shape[0]
    t = nears.clone()
    noises = torch.zeros(n_alive, device="cuda")
    x1, d1, t1 = raymarching.march_rays(n_alive, n_step, alive, t, o, d, 2.0, False, bitfield, 2, 128, nears, fars, False, 0.0, 1024)
    x2, d2, t2 = ref_cuda.march_rays(n_alive, n_step, alive, t, o, d, 2.0, False, bitfield, 2, 128, nears, fars, noises, 0.0, 1024)
    assert torch.equal(x1, x2) and torch.equal(d1, d2) and torch.equal(t1, t2)
    assert (t1[:, 0] > 0).float().mean().item() > 0.2
