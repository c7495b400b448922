"""Algorithmic flops of one get_xr_H call per bench.py hermitian workload = sum of 2*M*N*K over the pairwise contractions of
the recorded launch sequence -- a pure function of (state counts, orbital count, xr_order).  Computed here on the test-only
NumPy device stand-in (no GPU needed) and stored in profiles/hermitian_flops.json, which bench.py's reference arm reads
(the xr arm counts them from its own trace).      python tools/hermitian_flops.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from fake_xr import FakeDevice
from qodeapplications_b200.hermitian.plan import plan

table = {}
path = os.path.join(ROOT, "profiles", "hermitian_flops.json")
if os.path.exists(path):
    table = json.load(open(path))
for name in sys.argv[1:] or ["cfg2", "cfg1", "herm100"]:
    w = bench.WORKLOADS[name]
    system = bench.hermitian_system(name)
    ch = system["charges"]
    build = plan((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], w["xr_order"], [ch, ch], device=FakeDevice(), graph=False)
    flops = sum(2.0 * a[0] * a[1] * a[2] for call, a, k in build.trace if call.__name__ == "gemm_scatter")
    key = (tuple(sorted(system["n_states"].items())), system["n_orb"], w["xr_order"])
    table[str(key)] = flops
    print(name, key, flops, build.launches)
    json.dump(table, open(path, "w"), indent=1)
