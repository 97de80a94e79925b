#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference CUDA extensions
# (sources stay where they lie under /root/reference) for sm_100a into oracle/_ref/.
# The resulting pybind modules (_gridencoder, _raymarching_mob, _shencoder, _freqencoder) are the
# GPU-side parity oracle used by tests/ -m gpu and by bench.py's reference_cuda leg.
# Nothing under raw_ngp_b200/ may import them.
#
# Why not the reference's own build glue: */backend.py and */setup.py pass -std=c++14,
# which torch 2.11 headers reject (c10/util/C++17.h); we pass -std=c++17, nothing else
# differs from the reference flags (gridencoder/backend.py:6-9).
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT/obj"
if [ ! -d "$REF" ]; then echo "[build_ref] $REF absent -- using prebuilt $OUT"; exit 0; fi

PY=${PY:-python}
read -r TORCH_INC TORCH_LIB PY_INC EXT_SUFFIX < <($PY - <<'PYEOF'
import sysconfig, torch, os
from torch.utils import cpp_extension as c
inc = " ".join("-I"+p for p in c.include_paths(device_type="cuda") if os.path.isdir(p))
print(inc.replace(" ", ","), os.path.join(os.path.dirname(torch.__file__), "lib"),
      sysconfig.get_paths()["include"], sysconfig.get_config_var("EXT_SUFFIX"))
PYEOF
)
TORCH_INC=${TORCH_INC//,/ }
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
COMMON="-O3 -std=c++17 -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=1 $TORCH_INC -I$PY_INC"
NVFLAGS="-gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC \
 -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__"

build_one() { # dir  cu-basename  module-name
  local d=$1 cu=$2 mod=$3
  local so="$OUT/${mod}${EXT_SUFFIX}"
  if [ -f "$so" ] && [ "$so" -nt "$REF/$d/src/$cu.cu" ]; then echo "[build_ref] $mod up to date"; return; fi
  echo "[build_ref] compiling $d/src/$cu.cu -> $mod"
  $NVCC $COMMON $NVFLAGS -DTORCH_EXTENSION_NAME=$mod -c "$REF/$d/src/$cu.cu" -o "$OUT/obj/$mod.cu.o"
  g++ $COMMON -fPIC -DTORCH_EXTENSION_NAME=$mod -c "$REF/$d/src/bindings.cpp" -o "$OUT/obj/$mod.bind.o"
  g++ -shared "$OUT/obj/$mod.cu.o" "$OUT/obj/$mod.bind.o" -o "$so" \
      -L"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_cuda -ltorch_python -lc10 -lc10_cuda \
      -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TORCH_LIB"
  echo "[build_ref] built $so"
}

# The reference's own Python modules of the hot path, staged UNMODIFIED (byte copies + sha256 manifest) next to the compiled
# extensions so that tests on the GPU box (where /root/reference does not exist) can run the reference's nerf/network.py and
# nerf/renderer.py::run_cuda (a) over the reference kernels and (b) over raw_ngp_b200/dropin -- oracle/ref_stack.py.
# oracle/_ref/ is git-ignored: no reference source enters the history.
stage_py() {
  local PYOUT="$OUT/py"
  rm -rf "$PYOUT"; mkdir -p "$PYOUT"
  : > "$OUT/py_manifest.sha256"
  for f in nerf/network.py nerf/renderer.py encoding.py activation.py \
           gridencoder/__init__.py gridencoder/grid.py raymarching/__init__.py raymarching/raymarching.py \
           shencoder/__init__.py shencoder/sphere_harmonics.py freqencoder/__init__.py freqencoder/freq.py \
           barf/camera.py barf/camera_optimizers.py; do
    mkdir -p "$PYOUT/$(dirname "$f")"
    cp "$REF/$f" "$PYOUT/$f"
    (cd "$REF" && sha256sum "$f") >> "$OUT/py_manifest.sha256"
  done
  echo "[build_ref] staged $(wc -l < "$OUT/py_manifest.sha256") reference python modules under $PYOUT"
}
stage_py

build_one raymarching raymarching _raymarching_mob &
build_one shencoder   shencoder   _shencoder &
build_one gridencoder gridencoder _gridencoder &
build_one freqencoder freqencoder _freqencoder &
wait
ls -la "$OUT"
