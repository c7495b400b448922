"""Secondary measurement: hermitian-XRCC get_xr_H on Be2/6-31G shapes (cfg1: n=18, N=11/4/8) -- the
launch-latency-bound small case -- GPU drop-in vs the NumPy/einsum restatement of the reference
(what XRbase/XR_tensor.py:49-51 configures) on the host.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
from oracle import hermitian_oracle as ho

order = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name = sys.argv[2] if len(sys.argv) > 2 else "cfg1"
ops = {0: synth.OPS_ORDER0, 1: synth.OPS_ORDER1, 2: synth.OPS_ORDER2}[order]
system = synth.make_system(name, ops=ops, with_bior=True)
charges = system["charges"]
args = ((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges])
dev = Device(0)
get_xr_H(*args, device=dev)                 # warm-up (also uploads nothing persistent: each call re-uploads)
torch.cuda.synchronize()
times = []
for _ in range(5):
    n0 = dev.ctx.launch_count()
    t0 = time.perf_counter()
    H1, H2 = get_xr_H(*args, device=dev)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    launches = dev.ctx.launch_count() - n0
t0 = time.perf_counter()
R1, R2 = ho.get_xr_H(system["symm"], system["bior"], system["densities"][:2], order, [charges, charges])
cpu = time.perf_counter() - t0
err = float(numpy.abs(H2 - R2).max() / numpy.abs(R2).max())
print(json.dumps({"what": "get_xr_H", "config": name, "xr_order": order, "gpu_seconds_e2e_best": min(times), "gpu_seconds_all": times,
                  "xr_kernel_launches": launches, "numpy_oracle_seconds": cpu, "host_cores": os.cpu_count(),
                  "max_rel_err_vs_oracle": err, "dim_H2": int(H2.shape[0])}))
