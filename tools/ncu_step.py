#!/usr/bin/env python3
"""Developer tool: a few eager (non-graph) FusedTrainStep steps on the bench scene, for `ncu` captures.
Usage: python tools/ncu_step.py [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from raw_ngp_b200.trainer import FusedTrainStep  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda:0")
    model, o, d, tgt = bench.build_scene(dev, 0)
    step = FusedTrainStep(model, bench.RAYS_PER_GPU, use_graph=False)
    o, d, tgt = o.to(dev), d.to(dev), tgt.to(dev)
    for _ in range(steps):
        loss = step.step(o, d, tgt, update_grid=False)
    torch.cuda.synchronize()
    print("loss", float(loss.item()), "samples", step.last_num_points)


if __name__ == "__main__":
    main()
