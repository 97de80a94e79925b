#!/usr/bin/env python3
"""Evaluates the reference's hard-coded SH expressions *verbatim from its source text*
(/root/reference/shencoder/src/shencoder.cu:43-121 and :130-350) at seeded directions and stores the result as
tests/golden/sh_ref_source.npz.  Run in the build container (the reference tree is not on the GPU box):

    python tools/eval_reference_sh_source.py

No reference source is copied: the file is parsed at run time, each `outputs[i] = expr;` / `d{x,y,z}[i] = expr;`
line is turned into a Python expression (strip the f suffix, pow -> **) and evaluated in float64.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raw_ngp_b200 import synthetic  # noqa: E402

SRC = "/root/reference/shencoder/src/shencoder.cu"


def main():
    text = open(SRC).read()
    d = synthetic.unit_vectors(64, seed=3).numpy().astype(np.float64)
    d[32:] *= 0.8
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    env = dict(x=x, y=y, z=z, xy=x * y, xz=x * z, yz=y * z, x2=x * x, y2=y * y, z2=z * z, xyz=x * y * z)
    env.update(x4=env["x2"] ** 2, y4=env["y2"] ** 2, z4=env["z2"] ** 2)
    env.update(x6=env["x4"] * env["x2"], y6=env["y4"] * env["y2"], z6=env["z4"] * env["z2"])
    env["pow"] = np.power
    out = {k: np.zeros((64, 64)) for k in ("outputs", "dx", "dy", "dz")}
    pat = re.compile(r"^\s*(outputs|dx|dy|dz)\[(\d+)\]\s*=\s*(.*?);", re.M)
    n = 0
    for name, idx, expr in pat.findall(text):
        expr = re.sub(r"(\d+\.?\d*(?:[eE][-+]?\d+)?)f\b", r"\1", expr)
        out[name][:, int(idx)] = eval(expr, {"__builtins__": {}}, env) * np.ones(64)
        n += 1
    assert n == 4 * 64, n
    dst = os.path.join(ROOT, "tests", "golden", "sh_ref_source.npz")
    np.savez_compressed(dst, inputs=d, outputs=out["outputs"], dx=out["dx"], dy=out["dy"], dz=out["dz"])
    print("wrote", dst)


if __name__ == "__main__":
    main()
