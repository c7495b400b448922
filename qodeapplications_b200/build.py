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
SOURCES = ["xr_api.cu", "xr_gemm.cu", "xr_gemm_tma.cu", "xr_trimer.cu", "xr_embed.cu", "xr_linalg.cu", "xr_density.cu", "xr_scalar.cu", "xr_probe.cu"]
HEADERS = [os.path.join(CSRC, "xr_common.cuh"), os.path.join(HERE, "..", "include", "xr_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


OBJ_DIR = os.path.join(CSRC, "_obj")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def needs_build():
    return _stale(LIB, [os.path.join(CSRC, s) for s in SOURCES] + HEADERS)


def build(force=False, verbose=False):
    """One object per source (compiled side by side, rebuilt only when the source or a header is newer), then one link."""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared" and not f.startswith("--use_fast_math")]
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs, objects = [], []
    for src in SOURCES:
        path, obj = os.path.join(CSRC, src), os.path.join(OBJ_DIR, src[:-3] + ".o")
        objects.append(obj)
        if force or verbose or _stale(obj, [path] + HEADERS):
            jobs.append((src, subprocess.Popen([_nvcc()] + flags + ["-c", "-o", obj, path], stdout=subprocess.PIPE,
                                               stderr=subprocess.STDOUT, text=True)))
    for src, proc in jobs:
        out = proc.communicate()[0]
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            print(out)
    proc = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objects,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
