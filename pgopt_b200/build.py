"""Builds libpgopt_b200.so (hand-written sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

The library is a plain C-ABI shared object: no torch, no Python in it.  nvcc cross-compiles without
a GPU, so this runs in the CPU-only build container; the built .so travels to the GPU box.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpgopt_b200.so")
SOURCES = ["pgb_capi.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + \
    [os.path.join("..", "..", "include", "pgopt_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # products/sums never fuse unless written as __fma_rn (bit parity with the oracle)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, f) for f in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
