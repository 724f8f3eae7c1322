"""Builds libbsgpu.so (the CUDA kernels + C ABI) in-tree for sm_100a.

    python -m biostrings_b200.build

nvcc cross-compiles without a GPU; the resulting biostrings_b200/lib/libbsgpu.so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libbsgpu.so")
SOURCES = [os.path.join(CSRC, "bsgpu.cu")]
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
        if f.endswith((".cu", ".cuh", ".inc", ".h"))] + \
    [os.path.join(os.path.dirname(HERE), "include", "bsgpu.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--diag-suppress", "177"]


def nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(SO):
        return False
    t = os.path.getmtime(SO)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("BSGPU_NVCC_EXTRA", "").split()      # tuning sweeps: -DQS2_GROUP=8 ...
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + SOURCES
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
