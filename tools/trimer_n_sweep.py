"""xr_trimer_stream (reduce mode) over the orbital counts n the dispatcher covers: algorithmic TFLOP/s = 2*n*Pa*Pb*Pc / time
(padding of k to the instantiated cover is NOT counted as work).   python tools/trimer_n_sweep.py [Pa]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200.device import Device
from qodeapplications_b200 import lib as xr
Pa = int(sys.argv[1]) if len(sys.argv) > 1 else 592
Pb = Pc = 9984
dev = Device(0)
rng = numpy.random.default_rng(0)
for n in (int(x) for x in os.environ.get("XR_SWEEP_N", "4,6,8,9,10,12,13,14,16,17,18,19,20,24,28,32,36,40,44,48").split(",")):
    W, B, G = (dev.upload(rng.standard_normal(s)) for s in ((Pa, n * n + (n * n) % 2), (Pb, n), (Pc, n)))
    mom = dev.zeros((2,))
    run = lambda: dev.ctx.trimer_stream(n, Pa, Pb, Pc, 1.0, W, n * n + (n * n) % 2, B, n, G, n, 0, Pa, xr.TRIMER_REDUCE, mom)
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flops = 2.0 * Pa * Pb * Pc * n
    print(json.dumps({"n": n, "Pa": Pa, "Pb": Pb, "Pc": Pc, "ms": round(best, 3), "alg_tflops": round(flops / best / 1e9, 2),
                      "frac_of_37.19": round(flops / best / 1e9 / 37.19, 4)}), flush=True)
