"""Builds libpgopt_b200.so (hand-written sm_100a CUDA kernels + the C ABI) in-tree with nvcc.

The library is a plain C-ABI shared object: no torch, no Python in it.  nvcc cross-compiles without
a GPU, so this runs in the CPU-only build container; the built .so travels to the GPU box.  The
translation units are compiled in parallel (one nvcc per .cu).
"""
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libpgopt_b200.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + \
    [os.path.join("..", "..", "include", "pgopt_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # products/sums never fuse unless written as __fma_rn (bit parity with the oracle)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def _deps_mtime():
    deps = [os.path.join(CSRC, f) for f in HEADERS] + [os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps)


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return _deps_mtime() > t or any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES)


def build(force=False, verbose=False, extra_flags=(), out=None):
    out = out or LIB
    if not force and out == LIB and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = nvcc_path()
    flags = list(NVCC_FLAGS) + list(extra_flags)
    hdr_t = _deps_mtime()
    tag = "" if out == LIB else "_" + os.path.basename(out).replace(".", "_")

    def compile_one(src):
        obj = os.path.join(OBJDIR, src[:-3] + tag + ".o")
        spath = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(spath)):
            return obj, ""
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", spath, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        return obj, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    # NCCL (pgb_multi_*) is bound at run time with dlopen("libnccl.so.2"): the copy torch already loaded when there is
    # one, else the system library — the .so itself has no link-time NCCL dependency
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", out] + [o for o, _ in results] + ["-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
