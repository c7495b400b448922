"""Where does a get_xr_H call spend its time?  Host side (cProfile) and device side (CUPTI kernel table through
torch.profiler -- it sees every kernel of the process, ours included).   python tools/profile_hermitian.py <order> <config>"""
import cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
from qodeapplications_b200.hermitian.tensor import DeviceTensor

order = int(sys.argv[1]) if len(sys.argv) > 1 else 1
name = sys.argv[2] if len(sys.argv) > 2 else "cfg1"
ops = {0: synth.OPS_ORDER0, 1: synth.OPS_ORDER1, 2: synth.OPS_ORDER2}[order]
system = synth.make_system(name, ops=ops, with_bior=True)
charges = system["charges"]
dev = Device(0)
resident = []
for rho in system["densities"][:2]:
    held = {}
    for key, value in rho.items():
        if isinstance(value, dict) and key not in ("n_elec", "n_states", "n_states_bra", "KetCoeffs"):
            held[key] = {sector: DeviceTensor(dev.upload(block), dev) for sector, block in value.items()}
        else:
            held[key] = value
    resident.append(held)
args = ((system["symm"], system["bior"], system["nuc"]), resident, order, [charges, charges])
get_xr_H(*args, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter(); get_xr_H(*args, device=dev); torch.cuda.synchronize(); wall = time.perf_counter() - t0

prof = cProfile.Profile()
prof.enable(); get_xr_H(*args, device=dev); torch.cuda.synchronize(); prof.disable()
s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])

with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as tp:
    get_xr_H(*args, device=dev); torch.cuda.synchronize()
print(tp.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=70))
print(json.dumps({"config": name, "order": order, "wall_seconds_resident": wall}))
