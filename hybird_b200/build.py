"""Build liblbgpu.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("LBGPU_LIB") or os.path.join(HERE, "liblbgpu.so")  # LBGPU_LIB: A/B experiments with alternative builds
SOURCES = [os.path.join(HERE, "csrc", "lbgpu.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", f) for f in ("lb_kernels.cuh", "lb_d3q19.cuh", "lb_dem.cuh", "lbgpu_comm.h")] + \
    [os.path.join(ROOT, "include", "lbgpu.h")]

# -fmad=false: the reference is built for x86-64 without FMA contraction; fusing a*b+c on the
# device would change results in the last bit and, through the free-surface thresholds
# (mass > n, mass < 0), the cell-type map.  fp64 division and sqrt are IEEE-rounded by default.
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("LBGPU_EXTRA_FLAGS", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + SOURCES + ["-o", LIB]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
