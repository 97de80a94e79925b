#!/usr/bin/env python3
"""Developer tool: BASELINE configs[0] micro-benchmark only (GridEncoder fwd / bwd on 2^18 points)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
print(json.dumps(bench.encoder_micro(torch.device("cuda:0"), bench._peaks()[0]), indent=1))
