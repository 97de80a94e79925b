"""Builds raw_ngp_b200/lib/libngp_b200.so (C ABI, include/ngp_b200.h) from csrc/*.cu with nvcc for sm_100a.

In-tree build so that the .so travels with the repository snapshot; nothing is JIT-compiled at run time.
Usage:  python -m raw_ngp_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "_build")
LIB = os.path.join(LIBDIR, "libngp_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# No --use_fast_math: the marcher's sample counts are bit-compared with the reference.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-diag-suppress", "177",
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".inc", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(OBJDIR, "digest.txt")
    digest = _deps_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print(f"[raw_ngp_b200.build] {LIB} up to date")
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            if verbose:
                print(f"[raw_ngp_b200.build] nvcc not found, using prebuilt {LIB}")
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libngp_b200.so")

    def compile_one(src):
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        cmd = [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print(f"[raw_ngp_b200.build] built {LIB} from {len(objs)} translation units")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
