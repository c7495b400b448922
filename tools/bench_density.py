"""Density-tensor build: GPU drop-in (general/build_density_tensors.py, results left on the device) vs the reference's own
C (oracle/_ref/libdensity_tensors_ref.so, 1 host core, as build_density_tensors.py runs it).  One JSON line per case.
    python tools/bench_density.py [be|mid]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from oracle import density_oracle as do
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_density_tensors import build_density_tensors

CASES = {"be": dict(n_orbs=9, n_core=1, n_states={0: 11, +1: 4, -1: 8}, text="Be/6-31G fragment: 18 spin orbitals, 120/16/560 configurations"),
         "mid": dict(n_orbs=12, n_core=1, n_states={0: 48, +1: 17, -1: 35}, text="24 spin orbitals, 231/22/1540 configurations, 100 states")}
name = sys.argv[1] if len(sys.argv) > 1 else "be"
case = CASES[name]
n_orbs, n_core = case["n_orbs"], case["n_core"]
z_lists = do.make_states(n_orbs, n_core, 4, case["n_states"], seed=21)
n = 2 * n_orbs
V = numpy.random.default_rng(1).standard_normal((n,) * 4)
dev = Device(0)
build_density_tensors(z_lists, n_orbs, V, n_core, device=dev, device_result=True)
torch.cuda.synchronize()
times = []
for _ in range(3):
    n0 = dev.ctx.launch_count()
    t0 = time.perf_counter()
    rho, total = build_density_tensors(z_lists, n_orbs, V, n_core, device=dev, device_result=True)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
    launches = dev.ctx.launch_count() - n0
# elements actually formed (ccaa is formed in full before it is contracted with V)
formed = 0
for op in do.OPS:
    for bra in z_lists:
        ket = bra - do.op_dchg(op)
        if ket in z_lists:
            formed += z_lists[bra].coeffs.shape[0] * z_lists[ket].coeffs.shape[0] * n ** len(op)
ref = do.reference_c()
t0 = time.perf_counter()
checked = 0
for op in do.OPS:
    for bra in z_lists:
        ket = bra - do.op_dchg(op)
        if ket in z_lists and not (name == "mid" and op == "ccaa" and bra != +1):
            want = ref.tensor(op, z_lists, bra, ket, n_orbs, n_core)
            if op != "ccaa":
                assert numpy.array_equal(dev.download(rho[op][bra, ket]), want), (op, bra, ket)
            checked += want.size
cpu = time.perf_counter() - t0
print(json.dumps({"what": "build_density_tensors (all 8 operator strings, all charge pairs), results resident on the GPU", "case": name,
                  "description": case["text"], "gpu_seconds_best": min(times), "gpu_seconds_all": times, "xr_kernel_launches": launches,
                  "tensor_elements_formed": formed, "gpu_elements_per_s": formed / min(times), "gpu_write_GBs": 8 * formed / min(times) / 1e9,
                  "cpu_reference_c_seconds": cpu, "cpu_elements_checked_bit_exact": checked, "cpu_elements_per_s": checked / cpu,
                  "cpu_cores": 1, "speedup_vs_reference_c": (checked / cpu) and (formed / min(times)) / (checked / cpu)}))
