"""
In-tree build of ``libmdcraft_b200.so`` (hand-written sm_100a CUDA + the C ABI
of ``include/mdcraft_b200.h``).  ``nvcc`` cross-compiles without a GPU, so this
runs in the build container; the resulting ``mdcraft_b200/_lib/*.so`` travels to
the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libmdcraft_b200.so")
SOURCES = ("context.cu", "sq.cu", "sq_nufft.cu", "rdf.cu", "transforms.cu", "dump.cu", "profile.cu", "ncdf.cu")
HEADERS = ("common.cuh", "sq_math.cuh", "sq_nufft.cuh", os.path.join("..", "..", "include", "mdcraft_b200.h"))

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmdcraft_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles the library if it is missing or older than its sources; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("MDC_NVCC_EXTRA", "").split()   # kernel experiments only (e.g. -DMDC_NU_FFT_BLOCKS=8)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB_PATH, *SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
