"""CPU tests: the C-ABI library builds, loads and exports every symbol include/ngp_b200.h declares; host-side logic
(table layout, wrappers refusing CPU tensors, no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import grid_oracle
from raw_ngp_b200 import _lib
from raw_ngp_b200.gridencoder import GridEncoder
from raw_ngp_b200.gridencoder.grid import level_table_offsets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ngp_b200.h")).read()
    declared = set(re.findall(r"\b(ngp_[a-zA-Z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ngp_b200.h but not exported"
    assert declared == set(_lib.exported_symbols()), "ctypes signature table out of sync with the header"
    loaded = _lib.load()
    assert loaded.ngp_abi_version() == 3
    assert b"aligned" in loaded.ngp_status_string(-4)


def test_no_cpu_fallback():
    enc = GridEncoder(num_levels=4, log2_hashmap_size=10, desired_resolution=64)
    with pytest.raises(RuntimeError, match="CUDA-only|CPU tensor"):
        enc(torch.rand(8, 3))
    from raw_ngp_b200.shencoder import sh_encode
    with pytest.raises(RuntimeError):
        sh_encode(torch.rand(4, 3), 4, False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "raw_ngp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports oracle/"
                assert "/root/reference" not in src


@pytest.mark.parametrize("desired,bound", [(2048, 1), (2048, 2), (4096, 1)])
def test_table_layout_matches_reference_formula(desired, bound):
    enc = GridEncoder(desired_resolution=desired * bound)
    offs = enc.offsets.numpy()
    assert np.array_equal(offs, grid_oracle.table_offsets(3, 16, enc.per_level_scale, 16, 19))
    assert offs[0] == 0 and np.all(np.diff(offs) % 8 == 0) and np.all(np.diff(offs) <= 2 ** 19)
    if desired * bound == 2048:
        assert offs[-1] == 6098120 and enc.embeddings.shape == (6098120, 2)
        assert list(np.diff(offs)[:5]) == [4096, 12168, 29792, 79512, 205384]
    assert enc.output_dim == 32 and int(enc.n_params) == offs[-1] * 2
    assert level_table_offsets(3, 16, enc.per_level_scale, 16, 19) == list(offs)


def test_state_dict_keys_match_reference_checkpoint_layout():
    from raw_ngp_b200.nerf import NeRFNetwork, default_opt
    m = NeRFNetwork(default_opt(bound=2, rfield=True))
    keys = set(m.state_dict().keys())
    assert {"grid_encoder.embeddings", "grid_encoder.offsets", "grid_mlp.net.0.weight", "grid_mlp.net.2.weight",
            "view_mlp.net.0.weight", "view_mlp.net.2.weight", "density_grid", "density_bitfield", "aabb_train",
            "aabb_infer"} <= keys
    assert m.view_mlp.net[0].weight.shape == (80, 47) and m.grid_mlp.net[0].weight.shape == (64, 32)
    assert m.cascade == 2 and m.density_grid.shape == (2, 128 ** 3)
    m2 = NeRFNetwork(default_opt(bound=8, contract=True))
    assert m2.bound == 2 and m2.real_bound == 8 and m2.cascade == 2
