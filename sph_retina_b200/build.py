"""Builds the C-ABI CUDA library ``sph_retina_b200/_lib/libsphk.so`` for sm_100a with nvcc.

In-tree on purpose: the built ``.so`` travels to the GPU box with the repo snapshot and is the only
compute path of the package (there is no CPU or PyTorch fallback)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsphk.so")
SOURCES = ["sphk_kernels.cu"]
HEADERS = ["sphk_math.cuh", "sphk_fast.cuh", "sphk_grad.cuh", "sphk_coder.cuh", "sphk_obbloss.cuh", os.path.join("..", "..", "include", "sphk.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put it on PATH)")


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # build next to the target and rename atomically: several ranks may find the library stale at the same time
    tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    cmd += ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed building libsphk.so")
    os.replace(tmp, LIB_PATH)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
