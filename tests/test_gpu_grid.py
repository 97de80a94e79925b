"""GPU parity: raw_ngp_b200 grid encoder vs the reference's own compiled extension (oracle/_ref/_gridencoder),
through the C ABI, on identical seeded inputs.  Tolerances (BASELINE.json north_star): fp32 rel 1e-5, fp16 rel 1e-3;
table gradients are atomically accumulated on both sides, so only summation order differs."""
import numpy as np
import pytest
import torch

from raw_ngp_b200 import synthetic
from raw_ngp_b200.gridencoder import GridEncoder, grid_encode
from raw_ngp_b200.gridencoder import grid as grid_mod

pytestmark = pytest.mark.gpu


def _setup(D=3, C=2, L=16, log2T=19, base=16, desired=2048, dtype=torch.float32, B=20000, seed=0, gridtype="hash",
           align_corners=False, interpolation="linear", oob=True):
    enc = GridEncoder(input_dim=D, num_levels=L, level_dim=C, base_resolution=base, log2_hashmap_size=log2T,
                      desired_resolution=desired, gridtype=gridtype, align_corners=align_corners,
                      interpolation=interpolation)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, D, generator=g)
    if oob:  # a few points outside [0,1] and exactly on the faces
        x[:16] = x[:16] * 1.5 - 0.25
        x[16:24] = 0.0
        x[24:32] = 1.0
    emb = synthetic.table_values(enc.embeddings.shape[0], C, seed=seed, scale=1.0, dtype=dtype)
    return enc, x.cuda(), emb.cuda(), enc.offsets.cuda()


CASES = [
    dict(),                                              # model config: D3 C2 L16 T19 16->2048 hash linear
    dict(desired=4096),                                  # hash grid of the contracted scene (network.py:48)
    dict(C=1, L=8, log2T=14), dict(C=4, L=8, log2T=15), dict(C=8, L=4, log2T=15),
    dict(D=2, L=8, log2T=14, desired=512),
    dict(gridtype="tiled", log2T=15),
    dict(align_corners=True), dict(interpolation="smoothstep"),
    dict(interpolation="smoothstep", align_corners=True, gridtype="tiled", log2T=16),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_forward_matches_reference(case, dtype, ref_grid):
    from oracle import ref_cuda
    if dtype == torch.float16 and case.get("C") == 1:
        pytest.skip("reference drops fp16 C=1 gradients and its fp16 C=1 path is unused (gridencoder.cu:22-26)")
    enc, x, emb, off = _setup(dtype=dtype, **case)
    ours = grid_encode(x, emb, off, enc.per_level_scale, enc.base_resolution, False, enc.gridtype_id, enc.align_corners,
                       enc.interp_id, None)
    ref, _ = ref_cuda.grid_forward(x, emb, off, enc.per_level_scale, enc.base_resolution, False, enc.gridtype_id,
                                   enc.align_corners, enc.interp_id, None)
    assert ours.shape == ref.shape and ours.dtype == ref.dtype
    if dtype == torch.float32:
        torch.testing.assert_close(ours, ref, rtol=1e-5, atol=1e-6)
    else:
        torch.testing.assert_close(ours.float(), ref.float(), rtol=1e-3, atol=1e-3)
    # the arithmetic is meant to be the reference's, operation for operation
    exact = (ours == ref).float().mean().item()
    assert exact > 0.999, f"only {exact:.5f} of outputs bit-identical"


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_forward_max_level_and_jacobian(dtype, ref_grid):
    from oracle import ref_cuda
    enc, x, emb, off = _setup(dtype=dtype, B=5000)
    for max_level in (None, 7, 1):   # (the reference itself cannot launch max_level=0: grid.y == 0)
        ours, ours_j = grid_mod.grid_encode_with_jacobian(x, emb, off, enc.per_level_scale, enc.base_resolution,
                                                          enc.gridtype_id, enc.align_corners, enc.interp_id, max_level)
        ref, ref_j = ref_cuda.grid_forward(x, emb, off, enc.per_level_scale, enc.base_resolution, True, enc.gridtype_id,
                                           enc.align_corners, enc.interp_id, max_level)
        tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=2e-3, atol=2e-2)
        torch.testing.assert_close(ours.float(), ref.float(), **tol)
        # dy_dx is O(resolution * |table|): compare relative to its scale
        scale = ref_j.float().abs().max().clamp(min=1.0)
        torch.testing.assert_close(ours_j.float() / scale, ref_j.float() / scale, rtol=0,
                                   atol=1e-6 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize("case", CASES[:5] + CASES[6:9], ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_backward_matches_reference(case, dtype, ref_grid):
    from oracle import ref_cuda
    if dtype == torch.float16 and case.get("C") == 1:
        pytest.skip("reference silently drops fp16 C=1 table gradients (gridencoder.cu:22-26,341-346)")
    enc, x, emb, off = _setup(dtype=dtype, B=30000, **case)
    B, F = x.shape[0], enc.output_dim
    grad = (torch.randn(B, F, generator=torch.Generator().manual_seed(1)) * 1e-2).to(dtype).cuda()

    xr = x.clone().requires_grad_(True)
    er = emb.clone().requires_grad_(True)
    out = grid_encode(xr, er, off, enc.per_level_scale, enc.base_resolution, True, enc.gridtype_id, enc.align_corners,
                      enc.interp_id, None)
    out.backward(grad)

    _, dy_dx = ref_cuda.grid_forward(x, emb, off, enc.per_level_scale, enc.base_resolution, True, enc.gridtype_id,
                                     enc.align_corners, enc.interp_id, None)
    ref_gx, ref_ge = ref_cuda.grid_backward(grad, x, emb, off, enc.per_level_scale, enc.base_resolution, dy_dx,
                                            enc.gridtype_id, enc.align_corners, enc.interp_id, None)
    assert er.grad.dtype == ref_ge.dtype and er.grad.shape == ref_ge.shape
    if dtype == torch.float32:
        torch.testing.assert_close(er.grad, ref_ge, rtol=1e-4, atol=1e-7)
        scale = ref_gx.abs().max().clamp(min=1e-6)
        torch.testing.assert_close(xr.grad / scale, ref_gx / scale, rtol=0, atol=2e-5)
    else:
        # both sides accumulate in fp16 with atomics in arbitrary order: the reference does not reproduce itself run to run
        # (max ~ one fp16 ulp of the largest sums on a handful of the 12 M entries, tests/test_gpu_oracle_step.py measures up to
        # 2.7e-2 at configs[1] size), so the bound is on the mean, with a loose cap on the single worst entry
        scale = ref_ge.float().abs().max().clamp(min=1e-6)
        err = (er.grad.float() - ref_ge.float()).abs() / scale
        assert err.max().item() < 1e-2 and err.mean().item() < 1e-4, (err.max().item(), err.mean().item())
        # the reference accumulates dy_dx and the input gradient in half (gridencoder.cu:367-377); ours is fp32
        scale = ref_gx.abs().max().clamp(min=1e-6)
        err = ((xr.grad - ref_gx).abs() / scale)
        assert err.mean().item() < 5e-3 and err.max().item() < 0.1


def test_max_level_zero_writes_zeros():
    enc, x, emb, off = _setup(B=1000)
    out = grid_encode(x, emb, off, enc.per_level_scale, enc.base_resolution, False, 0, False, 0, 0)
    assert out.shape == (1000, 32) and (out == 0).all()


def test_input_gradient_matches_finite_difference_free_formula(ref_grid):
    """fp32: ours (recomputed in backward) == reference dy_dx contracted with grad, both exact formulas."""
    from oracle import ref_cuda
    enc, x, emb, off = _setup(B=4096, oob=False)
    grad = torch.randn(x.shape[0], enc.output_dim, generator=torch.Generator().manual_seed(3)).cuda()
    _, dy_dx = ref_cuda.grid_forward(x, emb, off, enc.per_level_scale, enc.base_resolution, True, 0, False, 0, None)
    L, C, D = enc.num_levels, enc.level_dim, 3
    expect = torch.einsum("blc,bldc->bd", grad.view(-1, L, C).double(), dy_dx.view(-1, L, D, C).double()).float()
    xr = x.clone().requires_grad_(True)
    out = grid_encode(xr, emb, off, enc.per_level_scale, enc.base_resolution, True, 0, False, 0, None)
    out.backward(grad)
    scale = expect.abs().max()
    torch.testing.assert_close(xr.grad / scale, expect / scale, rtol=0, atol=1e-5)


def test_bf16_against_fp32(ref_grid):
    """bf16 has no reference implementation (not in AT_DISPATCH_FLOATING_TYPES_AND_HALF): compare with the fp32
    reference evaluated on the bf16-rounded table."""
    from oracle import ref_cuda
    enc, x, emb, off = _setup(dtype=torch.bfloat16, B=8192)
    ours = grid_encode(x, emb, off, enc.per_level_scale, enc.base_resolution, False, 0, False, 0, None)
    ref, _ = ref_cuda.grid_forward(x, emb.float(), off, enc.per_level_scale, enc.base_resolution, False, 0, False, 0, None)
    torch.testing.assert_close(ours.float(), ref, rtol=8e-3, atol=8e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16], ids=["fp32", "fp16"])
def test_tv_and_weight_decay(dtype, ref_grid):
    """fp32: against the reference kernels.  fp16: the reference's TV kernel adds through the do-nothing
    atomicAdd(at::Half*) stub (gridencoder.cu:22-26,628), i.e. it is a silent no-op; ours applies the update, so
    the fp16 result is checked against the reference run in fp32 on the same (fp16-rounded) table."""
    from oracle import ref_cuda
    from raw_ngp_b200 import _lib
    enc, x, emb, off = _setup(dtype=dtype, B=50000, L=8, log2T=15, desired=256, oob=False)
    g0 = (torch.randn(emb.shape, generator=torch.Generator().manual_seed(5)) * 1e-3).to(dtype).cuda()
    L, C = enc.num_levels, enc.level_dim
    S = float(np.log2(enc.per_level_scale))

    ours = g0.clone()
    xin = x.to(dtype).contiguous()
    _lib.call("ngp_grid_grad_total_variation", xin.data_ptr(), emb.data_ptr(), ours.data_ptr(), off.data_ptr(), 1e-3,
              xin.shape[0], 3, C, L, S, enc.base_resolution, 0, 0, _lib.dtype_id(dtype), _lib.stream())
    ref = g0.float().clone()
    ref_cuda.grid_total_variation(xin.float().contiguous(), emb.float().contiguous(), ref, off, 1e-3, enc.per_level_scale,
                                  enc.base_resolution, 0, False)
    scale = (ref - g0.float()).abs().max()
    assert scale.item() > 1e-6
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert ((ours.float() - ref).abs() / scale).max().item() < tol
    if dtype == torch.float16:   # document the reference's no-op
        noop = g0.clone()
        ref_cuda.grid_total_variation(xin, emb, noop, off, 1e-3, enc.per_level_scale, enc.base_resolution, 0, False)
        assert torch.equal(noop, g0)

    ours, ref = g0.clone(), g0.clone()
    _lib.call("ngp_grid_grad_weight_decay", emb.data_ptr(), ours.data_ptr(), off.data_ptr(), 0.1, emb.shape[0], C, L,
              _lib.dtype_id(dtype), _lib.stream())
    ref_cuda.grid_weight_decay(emb, ref, off, 0.1)
    torch.testing.assert_close(ours.float(), ref.float(), rtol=1e-5 if dtype == torch.float32 else 2e-3, atol=1e-7)


def test_module_forward_backward_full_size(ref_grid):
    """BASELINE config 1 at full size (2^18 points): module API, fp16 table, against the reference kernels."""
    from oracle import ref_cuda
    torch.manual_seed(0)
    enc = GridEncoder(desired_resolution=2048).cuda()
    enc.embeddings.data = synthetic.table_values(enc.embeddings.shape[0], 2, seed=0, scale=1.0).cuda().half()
    x = synthetic.uniform_points(2 ** 18, seed=0).cuda()
    out = enc(x, bound=1)
    assert out.shape == (2 ** 18, 32) and out.dtype == torch.float16
    ref, _ = ref_cuda.grid_forward((x + 1) / 2, enc.embeddings.data, enc.offsets, enc.per_level_scale, 16)
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-3, atol=1e-3)
    grad = torch.randn(2 ** 18, 32, generator=torch.Generator().manual_seed(1)).half().cuda() * 1e-3
    out.backward(grad)
    _, ref_ge = ref_cuda.grid_backward(grad, (x + 1) / 2, enc.embeddings.data, enc.offsets, enc.per_level_scale, 16)
    scale = ref_ge.float().abs().max()
    assert ((enc.embeddings.grad.float() - ref_ge.float()).abs() / scale).max().item() < 1e-2
    # linearity property (size independent): encode(a*T) == a*encode(T) up to rounding
    enc2 = GridEncoder(desired_resolution=2048).cuda()
    enc2.embeddings.data = (enc.embeddings.data.float() * 0.5).half()
    torch.testing.assert_close(enc2(x).float(), out.detach().float() * 0.5, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("case", [dict(), dict(L=8, log2T=14), dict(gridtype="tiled", log2T=15), dict(L=32, log2T=12, desired=512),
                                  dict(interpolation="smoothstep", align_corners=True)],
                         ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_forward_tile_kernel_equals_point_level_kernels(case, dtype):
    """The tile kernel (128 points per CTA, thread = (point, level group), rows assembled in shared memory) against the
    one-thread-per-(point, level) kernel that is itself pinned to the reference above: ragged B, max_level < L (zero fill),
    out-of-range points, all three table dtypes."""
    import numpy as np
    from raw_ngp_b200 import _lib
    enc, x, emb, off = _setup(dtype=dtype, B=20001, **case)
    L = enc.num_levels
    S = float(np.log2(enc.per_level_scale))
    base_flags = _lib.NGP_GRID_REF_ROUNDING if dtype == torch.float16 else 0
    for max_level in (L, 7, 1):
        outs = []
        for flags in (base_flags, base_flags | _lib.NGP_GRID_POINT_LEVEL_KERNELS):
            out = torch.full((x.shape[0], L * 2), 7.0, device="cuda", dtype=dtype)
            _lib.call("ngp_grid_encode_forward", x.data_ptr(), emb.data_ptr(), off.data_ptr(), out.data_ptr(), x.shape[0], 3, 2, L, max_level,
                      S, enc.base_resolution, None, enc.gridtype_id, int(enc.align_corners), enc.interp_id, _lib.dtype_id(dtype), flags,
                      _lib.stream())
            outs.append(out)
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), (case, dtype, max_level, (outs[0] != outs[1]).float().mean().item())
        assert (outs[0][:, 2 * max_level:] == 0).all() and outs[0][64:, :2].abs().max() > 0
