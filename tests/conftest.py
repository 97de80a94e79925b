import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _load_ref_module(name):
    """Imports one of the reference's own CUDA extensions built by oracle/build_ref.sh into oracle/_ref/."""
    import glob
    import torch  # noqa: F401  (the extension links against libtorch)
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + ".*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ref_grid():
    m = _load_ref_module("_gridencoder")
    if m is None:
        pytest.skip("oracle/_ref/_gridencoder not built (run oracle/build_ref.sh where /root/reference exists)")
    return m


@pytest.fixture(scope="session")
def ref_march():
    m = _load_ref_module("_raymarching_mob")
    if m is None:
        pytest.skip("oracle/_ref/_raymarching_mob not built")
    return m


@pytest.fixture(scope="session")
def ref_sh():
    m = _load_ref_module("_shencoder")
    if m is None:
        pytest.skip("oracle/_ref/_shencoder not built")
    return m
