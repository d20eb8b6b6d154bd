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
TUNING_LIB_PATH = os.path.join(LIB_DIR, "libsphk_tuning.so")   # -DSPHK_TUNING: the A/B hooks of tools/ (never loaded by default)
CHECKED_LIB_PATH = os.path.join(LIB_DIR, "libsphk_checked.so") # -DSPHK_CHECKED: ring invariants / index ranges become traps (tools/checked_run.sh)
TIMELINE_LIB_PATH = os.path.join(LIB_DIR, "libsphk_tl.so")     # -DSPHK_TIMELINE: per-CTA / per-warp timestamps (tools/timeline_*.py)
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


def source_hash(tuning: bool = False) -> str:
    """sha256 over the sources, headers, compiler flags and this file: what the built library is a function of."""
    import hashlib
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]:
        if os.path.exists(f):
            h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS + (["-DSPHK_TUNING"] if tuning else [])).encode())
    return h.hexdigest()


def is_stale(target: str = LIB_PATH) -> bool:
    """True when the library is missing or was built from other sources.  Decided by content (a hash stamped next to
    the library at build time), not by modification times: a snapshot copied to a GPU box keeps no useful mtimes."""
    stamp = target + ".srchash"
    if not os.path.isfile(target) or not os.path.isfile(stamp):
        return True
    return open(stamp).read().strip() != source_hash(tuning=(target == TUNING_LIB_PATH))


def build(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """tuning=True builds the instrumented twin (environment-variable A/B hooks compiled in) next to the product
    library; tools/ scripts load it through SPHK_PROBE_LIB.  The product library reads no environment variable."""
    target = TUNING_LIB_PATH if tuning else LIB_PATH
    if not force and not is_stale(target):
        return target
    os.makedirs(LIB_DIR, exist_ok=True)
    # build next to the target and rename atomically: several ranks may find the library stale at the same time
    tmp = "%s.tmp.%d" % (target, os.getpid())
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DSPHK_TUNING"] if tuning else [])
    cmd += ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed building libsphk.so")
    os.replace(tmp, target)
    with open(target + ".srchash", "w") as f:
        f.write(source_hash(tuning))
    if verbose:
        sys.stderr.write(res.stderr)
    return target


def build_checked() -> str:
    """The twin whose shared-memory ring invariants and compacted-store index ranges trap (SPHK_CHECK in csrc/): loaded by
    tools/checked_run.sh through SPHK_PROBE_LIB -- the stand-in for compute-sanitizer where that tool is not available."""
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-DSPHK_CHECKED", "-o", CHECKED_LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return CHECKED_LIB_PATH


def build_timeline() -> str:
    """The timestamp-instrumented twin (tools/timeline_probe.py, tools/timeline_aligned.py)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-DSPHK_TIMELINE", "-DSPHK_TUNING", "-o", TIMELINE_LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True)
    return TIMELINE_LIB_PATH


if __name__ == "__main__":
    if "--checked" in sys.argv:
        print(build_checked())
        sys.exit(0)
    if "--timeline" in sys.argv:
        print(build_timeline())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))
