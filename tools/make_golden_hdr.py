#!/usr/bin/env python3
"""Generates tests/golden/hdr_loss.npz: the clipped, tone-curve weighted MSE of the raw / HDR training path, by EXECUTING the
reference's own lines (nerf/train_utils.py:512-541, the `image_mode == 'HDR'` branch of Trainer.train_step) on seeded inputs,
with the gradient of the loss with respect to the rendered colours from autograd.

    python tools/make_golden_hdr.py tests/golden

The lines are taken from the file where it lies under /root/reference and executed in a namespace that provides the few names
they use (`self.opt`, `self.device`, `data`, `gt_rgb`, `pred_rgb`, `raw_utils` unused with loss_weight='none').
tests/test_oracle_golden.py pins the torch restatement used by the GPU tests to it; tests/test_gpu_trainstep.py pins the fused
composite + loss kernel."""
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference/nerf/train_utils.py"


def main(out_dir):
    lines = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "if(self.opt.image_mode == 'HDR'):" in l)
    end = next(i for i in range(start, len(lines)) if "loss = (data_loss * lossmult_tensor * loss_weight).sum() / lossmult_tensor.sum()" in lines[i])
    body = textwrap.dedent("\n".join(lines[start + 1:end + 1]))
    g = torch.Generator().manual_seed(17)
    N = 193
    pred = (torch.rand(N, 3, generator=g) * 2.5).requires_grad_(True)          # some products exceed 1: the clip is active
    gt = torch.rand(N, 3, generator=g)
    exposure = np.array([1.0, 0.25, 1.0 / 16], dtype=np.float32)[np.arange(N) % 3]
    ns = dict(torch=torch, np=np, raw_utils=None, gt_rgb=gt, pred_rgb=pred, data={"exposure": exposure},
              self=types.SimpleNamespace(device="cpu", opt=types.SimpleNamespace(loss_weight="none")))
    exec(compile(body, "train_utils.py:hdr-branch", "exec"), ns)
    loss = ns["loss"]
    loss.backward()
    np.savez_compressed(os.path.join(out_dir, "hdr_loss.npz"), pred_rgb=pred.detach().numpy(), gt_rgb=gt.numpy(), exposure=exposure,
                        loss=np.float32(loss.item()), d_pred=pred.grad.numpy())
    print("loss", loss.item(), "clipped fraction", float((pred.detach() * torch.from_numpy(exposure)[:, None] > 1).float().mean()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden")
