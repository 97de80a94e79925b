"""Developer tool: the gather / reduction ceilings of tools/l2_ceiling.py as a function of how many SMs (512-thread blocks) take
part -- shows which side bounds them (per-SM L1 sector rate for gathers, the chip-wide L2 rate for reductions)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from raw_ngp_b200 import _lib
dev = torch.device("cuda:0")
n_rows = 1 << 23
table = torch.zeros(n_rows, dtype=torch.int32, device=dev); sink = torch.zeros(1, dtype=torch.int32, device=dev)
def rate(mode, blocks, rounds=64):
    ops = blocks * 512 * rounds * (8 if mode < 2 else 4)
    for _ in range(2): _lib.call("ngp_diag_l2_rate", _lib.ptr(table), n_rows, blocks, rounds, mode, _lib.ptr(sink), _lib.stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): _lib.call("ngp_diag_l2_rate", _lib.ptr(table), n_rows, blocks, rounds, mode, _lib.ptr(sink), _lib.stream())
    e1.record(); torch.cuda.synchronize()
    return ops * 5 / (e0.elapsed_time(e1) * 1e-3) / 1e9
for blocks in (37, 74, 148, 296, 592):
    print(f"blocks {blocks:4d} ({blocks*512/148:6.0f} threads/SM avg): gather {rate(0, blocks, 256):6.1f}  red {rate(1, blocks, 256):6.1f}  red.v2 {rate(2, blocks, 256):6.1f}  G ops/s")
