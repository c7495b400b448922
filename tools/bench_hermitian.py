"""Secondary measurement: hermitian-XRCC get_xr_H -- GPU drop-in vs the NumPy/einsum restatement of the reference
(what XRbase/XR_tensor.py:49-51 configures) on the host.     python tools/bench_hermitian.py <xr_order> <config>
cfg1 (n=18, N=11/4/8: Be2/6-31G shapes) is the launch-latency-bound small case; herm49 / herm100 (order 0) are the
larger-state-count regime.  Two GPU numbers: host densities in, host H out (everything re-uploaded every call), and
densities already resident as device tensors (what an optimiser loop that builds them on the GPU would see).
Prints one JSON line."""
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
# densities resident on the device (DeviceTensor blocks are used as they are)
from qodeapplications_b200.hermitian.tensor import DeviceTensor
resident = []
for rho in system["densities"][:2]:
    held = {}
    for key, value in rho.items():
        if isinstance(value, dict) and key not in ("n_elec", "n_states", "n_states_bra", "KetCoeffs"):
            held[key] = {sector: DeviceTensor(dev.upload(block), dev) for sector, block in value.items()}
        else:
            held[key] = value
    resident.append(held)
args_res = (args[0], resident) + args[2:]
get_xr_H(*args_res, device=dev)
torch.cuda.synchronize()
times_res = []
for _ in range(5):
    t0 = time.perf_counter()
    H1r, H2r = get_xr_H(*args_res, device=dev)
    torch.cuda.synchronize()
    times_res.append(time.perf_counter() - t0)
assert numpy.array_equal(H2r, H2)
t0 = time.perf_counter()
R1, R2 = ho.get_xr_H(system["symm"], system["bior"], system["densities"][:2], order, [charges, charges])
cpu = time.perf_counter() - t0
err = float(numpy.abs(H2 - R2).max() / numpy.abs(R2).max())
print(json.dumps({"what": "get_xr_H", "config": name, "xr_order": order, "gpu_seconds_e2e_best": min(times), "gpu_seconds_all": times,
                  "gpu_seconds_resident_densities_best": min(times_res), "density_GB": sum(b.nbytes for rho in system["densities"][:2] for k, v in rho.items() if isinstance(v, dict) for b in v.values() if hasattr(b, "nbytes")) / 1e9,
                  "xr_kernel_launches": launches, "numpy_oracle_seconds": cpu, "host_cores": os.cpu_count(),
                  "max_rel_err_vs_oracle": err, "dim_H2": int(H2.shape[0])}))
