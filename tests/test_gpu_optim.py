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


def _ptrs(tensors):
    import ctypes
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dp_fused_adam_matches_allreduce_then_adam(world):
    """csrc/optim.cu dp_fused_adam_kernel (reduce-scatter + Adam + all-gather over peer pointers), with the `world` ranks
    emulated by `world` sets of buffers on one GPU: every rank's launch updates its shard from ALL gradient buffers and
    stores into ALL parameter copies.  Expected: all copies bit-identical and equal to fused_adam on the fp32 sum."""
    torch.manual_seed(world)
    dev, n = "cuda", 8 * 12345
    per = (n + 8 * world - 1) // (8 * world) * 8
    master0 = torch.randn(n, device=dev)
    grads = [(torch.randn(n, device=dev) * 0.1 * 128).half() for _ in range(world)]
    tables = [master0.half().clone() for _ in range(world)]
    step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    inv_scale = torch.full((1,), 1.0 / (128.0 * world), device=dev)
    found_inf = torch.zeros(1, device=dev)
    shards = []
    for r in range(world):
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        shards.append(dict(lo=lo, hi=hi, master=master0[lo:hi].clone(), m=torch.zeros(hi - lo, device=dev), v=torch.zeros(hi - lo, device=dev)))
    # reference: sum in fp32 in rank order, then the single-GPU fused Adam
    ref_master = master0.clone()
    ref_lp = master0.half()
    ref_opt = FusedAdam(lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    ref_grad = torch.zeros(n, device=dev)
    ref_opt.add_group(ref_master, ref_grad, ref_lp)
    st = _lib.stream()
    for it in range(3):
        for g in grads:
            g.copy_((torch.randn(n, device=dev) * 0.1 * 128).half())
        ref_grad.zero_()
        for g in grads:
            ref_grad += g.float()
        ref_opt.step(inv_scale, found_inf, zero_grad=False)
        _lib.call("ngp_adam_step_counter", _lib.ptr(step_dev), _lib.ptr(found_inf), st)
        for r, sh in enumerate(shards):
            _lib.call("ngp_dp_fused_adam", _ptrs(grads), _lib.NGP_F16, _ptrs(tables), _lib.NGP_F16, world, world, _lib.ptr(sh["master"]),
                      _lib.ptr(sh["m"]), _lib.ptr(sh["v"]), sh["lo"], sh["hi"], 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(step_dev), None,
                      _lib.ptr(inv_scale), _lib.ptr(found_inf), None, 0, st)
        torch.cuda.synchronize()
        for t in tables[1:]:
            assert torch.equal(t, tables[0])
        full = torch.cat([sh["master"] for sh in shards])
        torch.testing.assert_close(full, ref_master, rtol=1e-6, atol=1e-7)
        assert (tables[0].float() - ref_lp.float()).abs().max().item() <= 2e-3      # same values up to one fp16 rounding
    # inf on any rank: nothing moves
    found_inf.fill_(1.0)
    before = [t.clone() for t in tables]
    sh = shards[0]
    _lib.call("ngp_dp_fused_adam", _ptrs(grads), _lib.NGP_F16, _ptrs(tables), _lib.NGP_F16, world, world, _lib.ptr(sh["master"]),
              _lib.ptr(sh["m"]), _lib.ptr(sh["v"]), sh["lo"], sh["hi"], 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(step_dev), None,
              _lib.ptr(inv_scale), _lib.ptr(found_inf), None, 0, st)
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(before, tables))


@pytest.mark.parametrize("bad_rank", [None, 1])
def test_dp_slim_chain_check_publish_adam_finish(bad_rank):
    """The data-parallel update chain as FusedTrainStep._peer_update launches it (check + flag publish | Adam with the flags
    merged in the kernel and step = counter + 1 | finish: merged flag, counter, gradient clear), `world` ranks emulated on one GPU
    in the order the barriers impose: every rank's check first, then every rank's Adam, then every rank's finish."""
    import ctypes
    torch.manual_seed(5)
    dev, world, n, nw = "cuda", 3, 8 * 4000, 256
    per = (n + 8 * world - 1) // (8 * world) * 8
    master0 = torch.randn(n, device=dev)
    grads = [(torch.randn(n, device=dev) * 0.1 * 128).half() for _ in range(world)]
    wgrads = [torch.randn(nw, device=dev) for _ in range(world)]
    if bad_rank is not None:
        grads[bad_rank][1234] = float("inf")
    tables = [master0.half().clone() for _ in range(world)]
    flags = [torch.zeros(8, device=dev) for _ in range(world)]
    steps = [torch.full((1,), 4, dtype=torch.int32, device=dev) for _ in range(world)]
    founds = [torch.full((1,), 9.0, device=dev) for _ in range(world)]
    scratch = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(world)]
    inv_scale = torch.full((1,), 1.0 / (128.0 * world), device=dev)
    st = _lib.stream()
    shards = []
    for r in range(world):
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        shards.append(dict(lo=lo, hi=hi, master=master0[lo:hi].clone(), m=torch.zeros(hi - lo, device=dev), v=torch.zeros(hi - lo, device=dev)))
    for r in range(world):
        args = ((ctypes.c_void_p * 2)(grads[r].data_ptr(), wgrads[r].data_ptr()), (ctypes.c_int * 2)(_lib.NGP_F16, _lib.NGP_F32),
                (ctypes.c_uint64 * 2)(n, nw))
        _lib.call("ngp_dp_check_publish", *args, 2, _lib.ptr(founds[r]), _lib.ptr(scratch[r]), _ptrs(flags), world, r, st)
    for r, sh in enumerate(shards):
        _lib.call("ngp_dp_fused_adam", _ptrs(grads), _lib.NGP_F16, _ptrs(tables), _lib.NGP_F16, world, world, _lib.ptr(sh["master"]),
                  _lib.ptr(sh["m"]), _lib.ptr(sh["v"]), sh["lo"], sh["hi"], 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(steps[r]), None,
                  _lib.ptr(inv_scale), None, _lib.ptr(flags[r]), world, st)
    for r in range(world):
        _lib.call("ngp_dp_finish", _lib.ptr(flags[r]), world, _lib.ptr(founds[r]), _lib.ptr(steps[r]), _lib.ptr(grads[r]), n * 2,
                  _lib.ptr(wgrads[r]), nw * 4, st)
    torch.cuda.synchronize()
    for r in range(world):
        assert flags[r][:world].tolist() == [1.0 if r2 == bad_rank else 0.0 for r2 in range(world)]
        assert founds[r].item() == (1.0 if bad_rank is not None else 0.0)
        assert steps[r].item() == (4 if bad_rank is not None else 5)
        assert grads[r].abs().max().item() == 0 and wgrads[r].abs().max().item() == 0
    if bad_rank is not None:
        assert all(torch.equal(t, master0.half()) for t in tables)          # skipped on every rank
    else:
        assert all(torch.equal(t, tables[0]) for t in tables[1:]) and not torch.equal(tables[0], master0.half())
        # the update used step count 5 (= counter + 1): compare with the plain kernel at step 5
        ref_tables = [master0.half().clone()]
        ref = dict(master=master0.clone(), m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev))
        # (gradients were cleared by finish: regenerate the same ones)
        torch.manual_seed(5)
        _ = torch.randn(n, device=dev)
        g2 = [(torch.randn(n, device=dev) * 0.1 * 128).half() for _ in range(world)]
        step5 = torch.full((1,), 5, dtype=torch.int32, device=dev)
        _lib.call("ngp_dp_fused_adam", _ptrs(g2), _lib.NGP_F16, _ptrs(ref_tables), _lib.NGP_F16, world, 1, _lib.ptr(ref["master"]),
                  _lib.ptr(ref["m"]), _lib.ptr(ref["v"]), 0, n, 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(step5), None, _lib.ptr(inv_scale),
                  None, None, 0, st)
        torch.cuda.synchronize()
        assert torch.equal(ref_tables[0], tables[0])


def test_dp_flags_and_argument_checks():
    dev, world = "cuda", 4
    flags = [torch.zeros(8, device=dev) for _ in range(world)]
    st = _lib.stream()
    for r in range(world):
        local = torch.full((1,), 1.0 if r == 2 else 0.0, device=dev)
        _lib.call("ngp_dp_publish_flag", _lib.ptr(local), _ptrs(flags), world, r, st)
    merged = torch.zeros(1, device=dev)
    for r in range(world):
        assert flags[r][:world].tolist() == [0.0, 0.0, 1.0, 0.0]
        _lib.call("ngp_dp_merge_flags", _lib.ptr(flags[r]), world, _lib.ptr(merged), st)
        assert merged.item() == 1.0
    # shard bounds must be multiples of the 16-byte vector, at most 8 peers
    g, t = [torch.zeros(64, device=dev).half()], [torch.zeros(64, device=dev).half()]
    m = torch.zeros(64, device=dev)
    step_dev = torch.ones(1, dtype=torch.int32, device=dev)
    with pytest.raises(RuntimeError):
        _lib.call("ngp_dp_fused_adam", _ptrs(g), _lib.NGP_F16, _ptrs(t), _lib.NGP_F16, 1, 1, _lib.ptr(m), _lib.ptr(m.clone()), _lib.ptr(m.clone()),
                  3, 35, 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(step_dev), None, None, None, None, 0, st)
    with pytest.raises(RuntimeError):
        _lib.call("ngp_dp_fused_adam", _ptrs(g * 9), _lib.NGP_F16, _ptrs(t * 9), _lib.NGP_F16, 9, 9, _lib.ptr(m), _lib.ptr(m.clone()),
                  _lib.ptr(m.clone()), 0, 64, 1e-2, 0.9, 0.99, 1e-15, 0.0, _lib.ptr(step_dev), None, None, None, None, 0, st)


@pytest.mark.parametrize("bad", [None, ("table", 0), ("table", 12196239), ("mlp", 13000), ("mlp", 3)])
def test_check_finite_multi(bad):
    """one launch: inf/nan over two buffers of different dtype -> found_inf overwritten (no zero-fill), step counted unless
    skipped, scratch self-resetting (called repeatedly)."""
    import ctypes
    dev = "cuda"
    table = (torch.randn(12196240, device=dev) * 0.1).half()
    mlp = torch.randn(13440, device=dev)
    found = torch.full((1,), 7.0, device=dev)              # garbage: must be overwritten
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    args = ((ctypes.c_void_p * 2)(table.data_ptr(), mlp.data_ptr()), (ctypes.c_int * 2)(_lib.NGP_F16, _lib.NGP_F32),
            (ctypes.c_uint64 * 2)(table.numel(), mlp.numel()))
    for rep in range(3):
        if bad is not None and rep == 1:
            (table if bad[0] == "table" else mlp)[bad[1]] = float("nan") if rep % 2 else float("inf")
        if rep == 2 and bad is not None:
            (table if bad[0] == "table" else mlp)[bad[1]] = 0.0
        _lib.call("ngp_check_finite_multi", *args, 2, _lib.ptr(found), _lib.ptr(step), _lib.ptr(scratch), _lib.stream())
        expect_inf = bad is not None and rep == 1
        assert found.item() == (1.0 if expect_inf else 0.0)
        assert scratch.tolist() == [0, 0]
    assert step.item() == (2 if bad is not None else 3)


def test_small_adam_matches_torch_and_skips_on_inf():
    """ngp_small_adam (one single-block launch: inf check + step count + unscale + Adam + clear) against torch.optim.Adam"""
    torch.manual_seed(3)
    dev, n = "cuda", 600
    master = torch.randn(n, device=dev) * 0.01
    ref = torch.nn.Parameter(master.clone())
    opt_ref = torch.optim.Adam([ref], lr=1e-3)                      # torch defaults, like barf/camera_optimizers.py:40
    grad, m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    lr_dev = torch.full((1,), 1e-3, device=dev)
    inv_scale = torch.full((1,), 1 / 128.0, device=dev)
    found = torch.full((1,), 5.0, device=dev)

    def launch():
        _lib.call("ngp_small_adam", _lib.ptr(master), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), n, 123.0, 0.9, 0.999, 1e-8, 0.0,
                  _lib.ptr(step), _lib.ptr(lr_dev), _lib.ptr(inv_scale), _lib.ptr(found), _lib.stream())

    for it in range(4):
        g = torch.randn(n, device=dev)
        grad.copy_(g * 128.0)
        ref.grad = g.clone()
        launch()
        opt_ref.step()
        assert found.item() == 0.0 and step.item() == it + 1 and grad.abs().max().item() == 0.0
        torch.testing.assert_close(master, ref.data, rtol=1e-5, atol=1e-7)
    before = master.clone()
    grad.fill_(1.0)
    grad[577] = float("nan")
    launch()
    assert found.item() == 1.0 and step.item() == 4 and torch.equal(master, before) and grad.abs().max().item() == 0.0


def test_dp_small_adam_sums_the_peer_buffers():
    """ngp_dp_small_adam (the pose optimizer's update in peer-memory data parallel: gradient = sum of every rank's buffer, read
    through the peer mappings) against torch.optim.Adam on the summed gradient; three local buffers stand in for three ranks.
    The buffers are left alone (ngp_dp_finish clears them behind the closing barrier); an inf in ANY rank's buffer skips the step."""
    import ctypes
    torch.manual_seed(4)
    dev, n, world = "cuda", 600, 3
    master = torch.randn(n, device=dev) * 0.01
    ref = torch.nn.Parameter(master.clone())
    opt_ref = torch.optim.Adam([ref], lr=1e-3)
    grads = [torch.zeros(n, device=dev) for _ in range(world)]
    peers = (ctypes.c_void_p * world)(*[g.data_ptr() for g in grads])
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    lr_dev = torch.full((1,), 1e-3, device=dev)
    inv_scale = torch.full((1,), 1 / (128.0 * world), device=dev)      # GradScaler scale x world: the mean over ranks
    found = torch.full((1,), 5.0, device=dev)

    def launch():
        _lib.call("ngp_dp_small_adam", peers, world, _lib.ptr(master), _lib.ptr(m), _lib.ptr(v), n, 123.0, 0.9, 0.999, 1e-8, 0.0,
                  _lib.ptr(step), _lib.ptr(lr_dev), _lib.ptr(inv_scale), _lib.ptr(found), _lib.stream())

    for it in range(4):
        gs = [torch.randn(n, device=dev) for _ in range(world)]
        for buf, g in zip(grads, gs):
            buf.copy_(g * 128.0)
        ref.grad = sum(gs) / world
        launch()
        opt_ref.step()
        assert found.item() == 0.0 and step.item() == it + 1
        assert all(torch.equal(buf, g * 128.0) for buf, g in zip(grads, gs))
        torch.testing.assert_close(master, ref.data, rtol=1e-5, atol=1e-7)
    before = master.clone()
    grads[2][13] = float("inf")
    launch()
    assert found.item() == 1.0 and step.item() == 4 and torch.equal(master, before)
