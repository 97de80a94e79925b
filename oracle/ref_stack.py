"""TEST INFRASTRUCTURE ONLY -- the reference's own, UNMODIFIED Python modules of the hot path (nerf/network.py, nerf/renderer.py,
encoding.py, activation.py, the gridencoder / raymarching / shencoder / freqencoder wrappers, barf/camera*.py), executed

  * ops="ref":     over the reference's own CUDA extensions (oracle/_ref/*.so)             -> the oracle
  * ops="dropin":  over raw_ngp_b200/dropin (this repository's operators, through the C ABI) -> the drop-in proof

The sources are byte copies staged by oracle/build_ref.sh into the git-ignored oracle/_ref/py/ (sha256 manifest beside
them), because /root/reference does not exist on the GPU box.  Third-party imports of those files that are absent from this
image and never touched by the hot path (mcubes, trimesh, tensorboardX, torch_efficient_distloss, pymeshlab via meshutils,
easydict, the trainer module nerf/train_utils.py with its dozen logging / metric imports) are replaced by empty stub
modules, exactly as SURVEY.md section 8(c) prescribes; `torch_scatter.segment_csr` (only used by
_march_rays_train.backward, raymarching.py:327-328) is shimmed with torch.segment_reduce.

Both variants can live in one process: each RefStack imports its own private copies of the modules (sys.modules is restored
afterwards); `with stack.active():` re-installs them for code with call-time imports (encoding.get_encoder).
"""
import contextlib
import importlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
PY_DIR = os.path.join(_HERE, "_ref", "py")
DROPIN_DIR = os.path.join(os.path.dirname(_HERE), "raw_ngp_b200", "dropin")

_OWNED = ("gridencoder", "raymarching", "shencoder", "freqencoder", "encoding", "activation", "nerf", "barf", "meshutils",
          "_gridencoder", "_raymarching_mob", "_shencoder", "_freqencoder", "torch_scatter", "mcubes", "trimesh",
          "tensorboardX", "torch_efficient_distloss", "easydict", "camera")


def available():
    return os.path.exists(os.path.join(PY_DIR, "nerf", "renderer.py"))


def _owned(name):
    return any(name == p or name.startswith(p + ".") for p in _OWNED)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__dict__.setdefault("__all__", [])
    return m


def _segment_csr(src, indptr, out=None, reduce="sum"):      # torch_scatter.segment_csr(src, indptr): CSR segmented sum
    assert reduce == "sum" and out is None
    return torch.segment_reduce(src, "sum", offsets=indptr, axis=0)


class RefStack:
    def __init__(self, ops):
        assert ops in ("ref", "dropin")
        if not available():
            raise RuntimeError("oracle/_ref/py not staged; run oracle/build_ref.sh where /root/reference exists")
        self.ops = ops
        saved = {k: v for k, v in sys.modules.items() if _owned(k)}
        saved_path = list(sys.path)
        for k in saved:
            del sys.modules[k]
        try:
            stubs = {
                "mcubes": _stub("mcubes"), "trimesh": _stub("trimesh"), "tensorboardX": _stub("tensorboardX"),
                "torch_efficient_distloss": _stub("torch_efficient_distloss", eff_distloss=None),
                "meshutils": _stub("meshutils"), "easydict": _stub("easydict", EasyDict=dict),
                "torch_scatter": _stub("torch_scatter", segment_csr=_segment_csr),
                "barf.pose_analysis": _stub("barf.pose_analysis"),
                # renderer.py:17 needs only custom_meshgrid from the trainer module (train_utils.py:41-46)
                "nerf.train_utils": _stub("nerf.train_utils", custom_meshgrid=lambda *a: torch.meshgrid(*a, indexing="ij")),
            }
            nerf_pkg = _stub("nerf")
            nerf_pkg.__path__ = [os.path.join(PY_DIR, "nerf")]
            barf_pkg = _stub("barf")
            barf_pkg.__path__ = [os.path.join(PY_DIR, "barf")]
            barf_pkg.pose_analysis = stubs["barf.pose_analysis"]
            sys.modules.update(stubs)
            sys.modules["nerf"], sys.modules["barf"] = nerf_pkg, barf_pkg
            if ops == "ref":
                from . import ref_cuda
                for n in ("_gridencoder", "_raymarching_mob", "_shencoder", "_freqencoder"):
                    sys.modules[n] = ref_cuda.module(n)         # the names the wrappers try first (grid.py:9-12)
                sys.path.insert(0, PY_DIR)
            else:
                sys.path.insert(0, PY_DIR)
                sys.path.insert(0, DROPIN_DIR)                   # dropin/ first on sys.path
            self.network = importlib.import_module("nerf.network")
            self.renderer = importlib.import_module("nerf.renderer")
            self.raymarching = importlib.import_module("raymarching")
            self.gridencoder = importlib.import_module("gridencoder")
            self.shencoder = importlib.import_module("shencoder")
            self.encoding = importlib.import_module("encoding")
            self.camera_optimizers = importlib.import_module("barf.camera_optimizers")
            if ops == "ref":        # raymarching.py:14-25 imports its extension lazily, on the first call: resolve it now
                importlib.import_module("raymarching.raymarching").get_backend()
            self.modules = {k: v for k, v in sys.modules.items() if _owned(k)}
        finally:
            for k in [k for k in sys.modules if _owned(k)]:
                del sys.modules[k]
            sys.modules.update(saved)
            sys.path[:] = saved_path
        want = DROPIN_DIR if ops == "dropin" else PY_DIR
        for pkg in (self.gridencoder, self.raymarching, self.shencoder):
            assert os.path.dirname(os.path.dirname(os.path.abspath(pkg.__file__))) == want, (pkg.__file__, want)
        assert os.path.abspath(self.network.__file__).startswith(PY_DIR) and os.path.abspath(self.renderer.__file__).startswith(PY_DIR)

    @contextlib.contextmanager
    def active(self):
        saved = {k: v for k, v in sys.modules.items() if _owned(k)}
        for k in saved:
            del sys.modules[k]
        sys.modules.update(self.modules)
        try:
            yield self
        finally:
            for k in [k for k in sys.modules if _owned(k)]:
                del sys.modules[k]
            sys.modules.update(saved)

    def make_opt(self, **overrides):
        """The fields of the reference CLI namespace (main.py:31-127) that network.py / renderer.py read."""
        opt = types.SimpleNamespace(
            bound=2, contract=False, grid_size=128, min_near=0.05, density_thresh=10, cuda_ray=True, dt_gamma=0,
            max_steps=1024, T_thresh=1e-8, fp16=True, hashmap_size=19, hashgrid_resolution=2048, rfield=False,
            pose_opt="none", internal_activation="relu", beta=1.0, density_activation="clamped_exp",
            color_activation="clamped_exp", start_annealing=0.0, end_annealing=0.5, lambda_orientation=0,
            compute_normals=False, device="cuda", num_cameras=0, noise=0.0, scale=1.0, c_lr=1e-3, iters=1000,
            lambda_proposal=0, lambda_distort=0)
        for k, v in overrides.items():
            setattr(opt, k, v)
        return opt

    def build_network(self, opt):
        with self.active():
            return self.network.NeRFNetwork(opt)


_cache = {}


def get(ops):
    if ops not in _cache:
        _cache[ops] = RefStack(ops)
    return _cache[ops]
