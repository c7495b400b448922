"""Builds qodeapplications_b200/libxr_b200.so (hand-written CUDA for sm_100a only) in-tree with nvcc.

    python -m qodeapplications_b200.build [--force] [--verbose]

The library has no Python/torch dependency: it is the C ABI of include/xr_b200.h.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libxr_b200.so")
SOURCES = ["xr_api.cu", "xr_gemm.cu", "xr_gemm_tma.cu", "xr_trimer.cu", "xr_embed.cu", "xr_linalg.cu", "xr_density.cu", "xr_scalar.cu"]
HEADERS = [os.path.join(CSRC, "xr_common.cuh"), os.path.join(HERE, "..", "include", "xr_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if verbose:
        flags += ["-Xptxas", "-v"]
    cmd = [_nvcc()] + flags + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
