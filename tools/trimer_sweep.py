"""Times xr_trimer_stream alone at cfg4-like class sizes (development tool).
    python tools/trimer_sweep.py [n] [Pa] [Pb] [Pc]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200.device import Device
from qodeapplications_b200 import lib as xr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
Pa = int(sys.argv[2]) if len(sys.argv) > 2 else 512
Pb = int(sys.argv[3]) if len(sys.argv) > 3 else 9984
Pc = int(sys.argv[4]) if len(sys.argv) > 4 else 9984
dev = Device(0)
rng = numpy.random.default_rng(0)
W, B, G = (dev.upload(rng.standard_normal(s)) for s in ((Pa, n * n), (Pb, n), (Pc, n)))
mom = dev.zeros((2,))
def run():
    dev.ctx.trimer_stream(n, Pa, Pb, Pc, 1.0, W, n * n, B, n, G, n, 0, Pa, xr.TRIMER_REDUCE, mom)
run(); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
flops = 2.0 * Pa * Pb * Pc * n
print(json.dumps({"epi": os.environ.get("XR_TRIMER_EPI", "0"), "n": n, "Pa": Pa, "Pb": Pb, "Pc": Pc, "ms": best,
                  "alg_tflops": flops / best / 1e9, "frac_of_37.19": flops / best / 1e9 / 37.19}))
