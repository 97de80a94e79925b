#!/usr/bin/env python3
"""Developer tool: render the configs[3] frame a few times (for an ncu launch list: the last frame's launches are the tail)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
t = bench.Timer(1, dev)
for i in range(n):
    out = bench.config3_block(dev, 0, 1, t, frames=1)
print(out["ms_per_frame"])
