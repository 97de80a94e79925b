"""CPU tests of the N>1 host logic with gloo, world_size 2: the data-parallel gradient exchange of
raw_ngp_b200.trainer (SUM all-reduce of the flat gradient buffers, MAX of the inf flag, division by
loss_scale * world inside the optimizer) and the ray sharding used by bench.py / inference tiles."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from raw_ngp_b200 import parallel
        g = torch.Generator().manual_seed(100 + rank)
        table_grad = torch.randn(1000, 2, generator=g)
        mlp_grad = torch.randn(333, generator=g)
        found_inf = torch.tensor([1.0 if rank == 1 else 0.0])
        loss_scale = 128.0
        local = (table_grad.clone(), mlp_grad.clone())
        parallel.all_reduce_gradients([table_grad, mlp_grad], found_inf)
        inv = parallel.unscale_factor(loss_scale, world)
        # every rank must hold the same averaged, unscaled gradient
        gathered = [torch.zeros_like(table_grad) for _ in range(world)]
        dist.all_gather(gathered, table_grad)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        assert found_inf.item() == 1.0                      # MAX over ranks
        # ray sharding: disjoint, covering, contiguous tiles
        lo, hi = parallel.shard_range(1920 * 1080, rank, world)
        out[rank] = dict(sum=table_grad.clone(), local=local[0], inv=inv, lo=lo, hi=hi, mlp=mlp_grad.clone(), mlp_local=local[1])
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    torch.testing.assert_close(r0["sum"], r0["local"] + r1["local"])
    torch.testing.assert_close(r0["mlp"], r0["mlp_local"] + r1["mlp_local"])
    assert r0["inv"] == pytest.approx(1.0 / (128.0 * 2))
    assert (r0["lo"], r0["hi"]) == (0, 1036800) and (r1["lo"], r1["hi"]) == (1036800, 2073600)


def test_shard_range_properties():
    from raw_ngp_b200 import parallel
    for n in (0, 1, 7, 4096, 2073600):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
