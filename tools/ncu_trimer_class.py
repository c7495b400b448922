"""One launch of the trimer stream at the FULL 'ex' class shape of cfg4 (Pk = 15272, Pb = Pc = 9984, n = 18: 3.17e13 flop, ~1 s),
for the DRAM-traffic capture behind bench.py's roofline.traffic:

    ncu --set full --clock-control none --import-source on -k regex:trimer_stream -c 1 -o gpurun_out/r02_trimer_class \
        python tools/ncu_trimer_class.py
    python tools/ncu_summary.py gpurun_out/r02_trimer_class.ncu-rep profiles/r02_trimer_class_ncu_full.json    (here)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qodeapplications_b200.device import Device
from qodeapplications_b200 import lib as xr

dev = Device(0)
rand = lambda *shape: torch.randn(shape, dtype=torch.float64, device=dev.torch_device)
n, Pa, Pb, Pc = 18, int(os.environ.get("XR_PA", "15272")), 9984, 9984
W, B, G, mom = rand(Pa, n * n), rand(Pb, n), rand(Pc, n), dev.zeros((2,))
dev.ctx.trimer_stream(n, Pa, Pb, Pc, 1.0, W, n * n, B, n, G, n, 0, Pa, xr.TRIMER_REDUCE, mom)
torch.cuda.synchronize()
print("done", mom.cpu().numpy())
