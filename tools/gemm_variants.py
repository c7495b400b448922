"""Builds libxr_b200.so once per scatter-GEMM scheduling variant (csrc/xr_gemm_tma.cu, XR_GEMM_VARIANT) so that one GPU call
can check and time them all on the real dimer-class shapes:

    python tools/gemm_variants.py build            # here (nvcc cross-compiles); objects of the other sources are reused
    python tools/gemm_variants.py run              # on the GPU box: parity tests + class timings of every variant, JSON lines

Timing = the five charge-transfer classes of one cfg4 dimer (40 000^2 H2, offset-table epilogue into the final layout) through
build_matrix_elements.H2_device, plus two plain products (contiguous C) that separate the write pattern from the schedule.
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qodeapplications_b200 import build as xr_build

OUT = os.path.join(ROOT, "tools", "variants")
VARIANTS = {
    "wholek": ["-DXR_GEMM_WHOLEK=1"],
    "per_ktile_ring": ["-DXR_GEMM_WHOLEK=0"],
}


def build():
    os.makedirs(OUT, exist_ok=True)
    xr_build.build()                                   # fresh objects of every source in csrc/_obj
    flags = [f for f in xr_build.NVCC_FLAGS if not f.startswith("--use_fast_math") and f != "-shared"]
    nvcc = xr_build._nvcc()
    others = [os.path.join(xr_build.OBJ_DIR, s[:-3] + ".o") for s in xr_build.SOURCES if s != "xr_gemm_tma.cu"]
    procs = []
    for name, defs in VARIANTS.items():
        o = os.path.join(OUT, "xr_gemm_tma_%s.o" % name)
        procs.append((name, o, subprocess.Popen([nvcc] + flags + defs + ["-c", os.path.join(xr_build.CSRC, "xr_gemm_tma.cu"), "-o", o])))
    for name, o, proc in procs:
        assert proc.wait() == 0, name
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(OUT, "libxr_%s.so" % name), o] + others)


TIMING = r'''
import json, sys, itertools
import numpy, torch
sys.path.insert(0, %r)
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_H import build_matrix_elements
dev = Device(0)
system = synth.make_system(n_frag=2, n_orb=18, n_states={0: 96, +1: 34, -1: 70}, seed=4, ops=synth.OPS_GENERAL, general_ccaa="random")
eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
out = dev.empty((40000, 40000))
eng.H2_device(0, 1, out=out)
torch.cuda.synchronize()
eng.profile = []
for _ in range(3):
    eng.H2_device(0, 1, out=out)
torch.cuda.synchronize()
best = {}
for label, flops, nbytes, e0, e1 in eng.profile:
    ms = e0.elapsed_time(e1)
    if label not in best or ms < best[label][0]:
        best[label] = (ms, flops, nbytes)
rec = {label: {"ms": ms, "tflops": f / ms / 1e9, "gbs": b / ms / 1e6} for label, (ms, f, b) in sorted(best.items())}
del out
rng = numpy.random.default_rng(0)
for (M, N, K) in [(9984, 9984, 36), (15272, 15272, 326)]:
    ld = K + (K & 1)
    A = torch.randn((M, ld), dtype=torch.float64, device=dev.torch_device)
    B = torch.randn((N, ld), dtype=torch.float64, device=dev.torch_device)
    C = dev.empty((M, N))
    run = lambda: dev.ctx.gemm_scatter(M, N, K, 1.0, A, ld, B, ld, C, None, N, None, False)
    run(); torch.cuda.synchronize()
    t = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        t = min(t, e0.elapsed_time(e1))
    rec["plain_%%dx%%dx%%d" %% (M, N, K)] = {"ms": t, "tflops": 2.0 * M * N * K / t / 1e9, "gbs": 8.0 * M * N / t / 1e6}
    del A, B, C
print(json.dumps(rec))
''' % ROOT


def run():
    for name in VARIANTS:
        env = dict(os.environ, XR_B200_LIB=os.path.join(OUT, "libxr_%s.so" % name))
        t = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
                            os.path.join(ROOT, "tests", "test_general_gpu.py"), "-k", "gemm_scatter or blocks_match or cfg1_dimer"],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        rec = {"variant": name, "parity": "ok" if t.returncode == 0 else t.stdout[-600:]}
        r = subprocess.run([sys.executable, "-c", TIMING], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        try:
            rec["classes"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            rec["classes"] = r.stdout[-1500:]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    build() if sys.argv[1] == "build" else run()
