"""Probe: which way of getting peer-accessible buffers across ranks works on this box (symmetric memory, CUDA IPC)."""
import os, traceback
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = {}
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float16, device=dev)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float16)
    v = float(peer[:4].float().sum().item())
    ok["symm_mem"] = (v, [hex(p) for p in hdl.buffer_ptrs][:world], type(hdl).__name__)
    hdl.barrier()
except Exception as e:
    ok["symm_mem"] = "FAILED: " + repr(e)[:300]
try:
    x = torch.full((1 << 20,), float(rank + 1), dtype=torch.float16, device=dev)
    h = x.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    pr = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*hs[pr])
    y = torch.empty(0, dtype=torch.float16, device=torch.device("cuda", hs[pr][0])).set_(st, 0, (1 << 20,))
    torch.cuda.synchronize(); dist.barrier()
    ok["ipc"] = (float(y[:4].float().sum().item()), str(y.device), hex(y.data_ptr()), torch.cuda.can_device_access_peer(local, pr))
    # read the peer buffer from a kernel running on THIS device
    z = torch.empty(4, dtype=torch.float16, device=dev)
    try:
        z.copy_(y[:4]); ok["ipc_copy"] = z.float().tolist()
    except Exception as e:
        ok["ipc_copy"] = repr(e)[:200]
    dist.barrier()
except Exception as e:
    ok["ipc"] = "FAILED: " + repr(e)[:300] + traceback.format_exc()[-400:]
print(rank, ok, flush=True)
dist.destroy_process_group()
