"""Builds libxr_b200.so variants side by side (tools/variants/libxr_<name>.so) so that one GPU call can time them all:

    python tools/trimer_variants.py build            # here (nvcc cross-compiles)
    python tools/trimer_variants.py run [Pa]         # on the GPU box: parity test + timing of every variant

Variants are the compile-time switches of csrc/xr_trimer.cu (see its header).  The switches that lost on B200
(profiles/r01q_trimer_variants.jsonl) were removed from the kernel again.
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qodeapplications_b200 import build as xr_build

OUT = os.path.join(ROOT, "tools", "variants")
VARIANTS = {
    "sep": [],                                  # shipped: DFMA k-tail in its own basic block
    "nosep": ["-DXR_TRIMER_SEP_TAIL=0"],        # ptxas free to interleave the tail with the DMMAs
    "stream_sum": ["-DXR_TRIMER_STREAM_SUM=1"],  # first moment accumulated element by element as well
}
SHAPES = {}
ONLY_TRIMER_CU = True     # the variants differ in xr_trimer.cu only: compile that file per variant, link the rest once


def build():
    os.makedirs(OUT, exist_ok=True)
    flags = [f for f in xr_build.NVCC_FLAGS if not f.startswith("--use_fast_math") and f != "-shared"]
    nvcc = xr_build._nvcc()
    objs = []
    for s in xr_build.SOURCES:
        if s == "xr_trimer.cu":
            continue
        o = os.path.join(OUT, s.replace(".cu", ".o"))
        subprocess.check_call([nvcc] + flags + ["-c", os.path.join(xr_build.CSRC, s), "-o", o])
        objs.append(o)
    for name, defs in VARIANTS.items():
        o = os.path.join(OUT, "xr_trimer_%s.o" % name)
        subprocess.check_call([nvcc] + flags + defs + ["-c", os.path.join(xr_build.CSRC, "xr_trimer.cu"), "-o", o])
        subprocess.check_call([nvcc, "-shared", "-o", os.path.join(OUT, "libxr_%s.so" % name), o] + objs)


def run(args):
    out = []
    for name in list(VARIANTS):
        env = dict(os.environ, XR_B200_LIB=os.path.join(OUT, "libxr_%s.so" % name))
        t = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", os.path.join(ROOT, "tests", "test_general_gpu.py"),
                            "-k", "test_trimer_stream"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        rec = {"variant": name, "parity": "ok" if t.returncode == 0 else t.stdout[-300:]}
        for shape in (SHAPES.get(name) or ["18"] + (args or ["1184", "9984", "9984"]),):     # 148 x 624 work items: 624 per persistent CTA
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "trimer_sweep.py")] + shape, env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True)
            try:
                rec.setdefault("timing", []).append(json.loads(r.stdout.strip().splitlines()[-1]))
            except Exception:
                rec.setdefault("timing", []).append(r.stdout[-1500:])
        print(json.dumps(rec), flush=True)
        out.append(rec)
    return out


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2:])
