"""Kernel-time table of one build_density_tensors call (CUPTI through torch.profiler).  python tools/profile_density.py [be|mid]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from oracle import density_oracle as do
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_density_tensors import build_density_tensors
case = {"be": (9, 1, {0: 11, +1: 4, -1: 8}), "mid": (12, 1, {0: 48, +1: 17, -1: 35})}[sys.argv[1] if len(sys.argv) > 1 else "mid"]
n_orbs, n_core, n_states = case
z = do.make_states(n_orbs, n_core, 4, n_states, seed=21)
V = numpy.random.default_rng(1).standard_normal((2 * n_orbs,) * 4)
dev = Device(0)
build_density_tensors(z, n_orbs, V, n_core, device=dev, device_result=True); torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as tp:
    build_density_tensors(z, n_orbs, V, n_core, device=dev, device_result=True); torch.cuda.synchronize()
print(tp.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
