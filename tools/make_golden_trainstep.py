#!/usr/bin/env python3
"""Generates tests/golden/train_step_loss.npz by EXECUTING the reference's own lines of Trainer.train_step
(nerf/train_utils.py:494-557: background selection incl. per-ray random, RGBA ground truth, the render call, the HDR loss with
lossmult / loss_weight or the MSE, the entropy regulariser) on seeded inputs, gradients from autograd:

    python tools/make_golden_trainstep.py tests/golden

The lines run where they lie (/root/reference), in a namespace that provides the names they use: `self.opt`, `self.device`,
`self.criterion`, `data`, `images`, `N`, `C`, `rays_*` and a `self.model.render` stub that turns given composited colours /
opacities into `outputs` the way run_cuda's last lines do (renderer.py:672: image = image + (1 - weights_sum)[..., None] *
bg_color).  raw_utils is the reference's own raw/raw_utils.py (gaussian / planck / hanning weighting).
tests/test_gpu_trainstep.py feeds the same numbers to ngp_composite_train_loss as rays with one sample each."""
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"


def _raw_utils():
    spec = importlib.util.spec_from_file_location("ref_raw_utils", os.path.join(REF, "raw", "raw_utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def run_case(body, raw_utils, seed, N, image_mode, background, rgba, lossmult, loss_weight, lambda_entropy):
    g = torch.Generator().manual_seed(seed)
    colour = torch.rand(N, 3, generator=g) * (2.5 if image_mode == "HDR" else 1.0)       # colour of the ray's single sample
    ws = torch.rand(N, generator=g)
    ws[:4] = torch.tensor([0.0, 1.0, 5e-6, 1 - 5e-6])        # the clamp of the entropy term is active on these
    comp = (ws[:, None] * colour).requires_grad_(True)         # composited colour of a one-sample ray: w * c with w = alpha = ws
    ws = ws.requires_grad_(True)
    images = torch.rand(N, 4 if rgba else 3, generator=g)
    exposure = np.array([1.0, 0.25, 1.0 / 16], dtype=np.float32)[np.arange(N) % 3]
    data = {"exposure": exposure}
    if lossmult:
        bayer = np.zeros((N, 3), dtype=np.float32)
        bayer[np.arange(N), np.arange(N) % 3] = 1.0            # one colour per pixel, like a mosaiced sensor
        data["lossmult"] = bayer
    captured = {}

    def render(rays_o, rays_d, **kw):
        bg = kw["bg_color"]
        captured["bg"] = bg
        image = comp + (1 - ws).unsqueeze(-1) * bg           # renderer.py:672
        return {"image": image, "weights_sum": ws, "depth": torch.zeros(N), "num_points": N}
    opt = types.SimpleNamespace(background=background, image_mode=image_mode, loss_weight=loss_weight, lambda_entropy=lambda_entropy,
                                lambda_proposal=0, lambda_orientation=0, lambda_distort=0, diffuse_step=0, rfield=False)
    selfobj = types.SimpleNamespace(device="cpu", opt=opt, global_step=10, criterion=torch.nn.MSELoss(reduction="none"),
                                    model=types.SimpleNamespace(render=render))
    ns = dict(torch=torch, np=np, raw_utils=raw_utils, self=selfobj, data=data, images=images, N=N, C=images.shape[1], rays_o=None,
              rays_d=None, rays_ldir=None, cam_near_far=None)
    torch.manual_seed(seed + 1)                                # the per-ray random background is torch.rand(N, 3)
    exec(compile(body, "train_utils.py:494-557", "exec"), ns)
    loss = ns["loss"]
    loss.backward()
    bg = captured["bg"]
    bg = bg.numpy() if torch.is_tensor(bg) else np.full((N, 3), float(bg), dtype=np.float32)
    lw = ns.get("loss_weight", 1.0) if image_mode == "HDR" else 1.0
    lw = torch.broadcast_to(torch.as_tensor(lw, dtype=torch.float32), (N, 3)).numpy().copy()
    return dict(colour=colour.numpy(), comp=comp.detach().numpy(), ws=ws.detach().numpy(), images=images.numpy(), exposure=exposure, bg=bg.astype(np.float32),
                lossmult=data.get("lossmult", np.ones((N, 3), np.float32)), loss_weight=lw, gt=ns["gt_rgb"].detach().numpy(),
                loss=np.float32(loss.item()), d_comp=comp.grad.numpy(), d_ws=ws.grad.numpy(), lambda_entropy=np.float32(lambda_entropy),
                hdr=np.int32(image_mode == "HDR"))


def main(out_dir):
    lines = open(os.path.join(REF, "nerf", "train_utils.py")).read().splitlines()
    start = next(i for i, l in enumerate(lines) if "if self.opt.background == 'random':" in l)
    end = next(i for i, l in enumerate(lines) if "loss = loss + self.opt.lambda_entropy * (entropy.mean())" in l)
    body = "\n".join(lines[start:end + 1])
    # the tensorboard histogram line sits inside the slice only after `end`; the slice is executed verbatim
    body = textwrap.dedent(body)
    ru = _raw_utils()
    cases = {
        "hdr_random_rgba_bayer_gaussian_entropy": dict(seed=1, N=257, image_mode="HDR", background="random", rgba=True, lossmult=True,
                                                       loss_weight="gaussian", lambda_entropy=1e-2),
        "hdr_white_planck": dict(seed=2, N=130, image_mode="HDR", background="white", rgba=False, lossmult=False, loss_weight="planck",
                                 lambda_entropy=0.0),
        "hdr_black_hanning_bayer": dict(seed=3, N=96, image_mode="HDR", background="black", rgba=True, lossmult=True, loss_weight="hanning",
                                        lambda_entropy=0.0),
        "mse_random_rgba_entropy": dict(seed=4, N=200, image_mode="LDR", background="random", rgba=True, lossmult=False, loss_weight="none",
                                        lambda_entropy=5e-3),
    }
    out = {}
    for name, kw in cases.items():
        r = run_case(body, ru, **kw)
        print(name, "loss", float(r["loss"]))
        for k, v in r.items():
            out[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(out_dir, "train_step_loss.npz"), **out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden")
