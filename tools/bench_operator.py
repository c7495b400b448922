"""Matrix-free H.v (general/operator.py) on three fragments of cfg4 (200 states each: the product basis has 8e6 states, the
dense H3 would have 6.4e13 elements).  Prints one JSON line: preparation and apply times, and the cross-check
sum(H3 . 1) == streamed sum of all H3 elements (xr_trimer_stream moments) -- two independent GPU paths.
    python tools/bench_operator.py [config] [n_vectors]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_H import build_matrix_elements
from qodeapplications_b200.general.operator import xr_operator

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
nvec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
system = synth.make_system(name, general_ccaa="random")
dev = Device(0)
eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
frags = [0, 1, 2]
t0 = time.perf_counter()
op = xr_operator(eng, fragments=frags)
torch.cuda.synchronize()
prepare = time.perf_counter() - t0
dims = op.dims
v = torch.randn(tuple(dims) + ((nvec,) if nvec > 1 else ()), dtype=torch.float64, device=dev.torch_device)
op.apply(v); torch.cuda.synchronize()
times = []
for _ in range(3):
    n0 = dev.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0 = op.contractor.flops
    e0.record(); y = op.apply(v); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / 1e3)
    launches, flops = dev.ctx.launch_count() - n0, op.contractor.flops - f0
tri = xr_operator(eng, fragments=frags, monomers=False, dimers=False)
ones = torch.ones(tuple(dims), dtype=torch.float64, device=dev.torch_device)
total = float(tri.apply(ones).sum())
t0 = time.perf_counter()
streamed, sumsq = eng.H3_moments(*frags)
stream_s = time.perf_counter() - t0
print(json.dumps({"what": "matrix-free H.v from class factors", "config": name, "fragments": frags, "dims": dims, "n_vectors": nvec,
                  "terms": len(op.terms), "prepare_seconds": prepare, "apply_seconds_best": min(times), "apply_seconds_all": times,
                  "apply_gemm_tflop": flops / 1e12, "apply_tflops": flops / min(times) / 1e12, "xr_kernel_launches": launches,
                  "dense_H3_elements": float(numpy.prod(dims)) ** 2,
                  "check_sum_H3_operator": total, "check_sum_H3_streamed": streamed,
                  "check_abs_diff_over_norm": abs(total - streamed) / (sumsq * float(numpy.prod(dims)) ** 2) ** 0.5,
                  "streamed_build_seconds_for_comparison": stream_s}))
