#!/usr/bin/env python3
"""Developer tool: a few steps of the configs[2] light-stage training step (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda:0")
t = bench.Timer(1, dev)
r = bench.config2_block(dev, t, int(sys.argv[1]) if len(sys.argv) > 1 else 3, 1)
print(r["ms_per_step"], r["step_kernels_ms"])
