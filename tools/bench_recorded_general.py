"""general-XRCC build of a SMALL system (Be2 / Be3 shapes) eager vs recorded (general/distributed.recorded_step = one CUDA graph):
    python tools/bench_recorded_general.py cfg3          -> one JSON line
The eager step is the Python host issuing a few hundred launches; the recorded step is the same launches as one graph."""
import itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_H import build_matrix_elements
from qodeapplications_b200.general.distributed import sharded_build

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
system = synth.make_system(name)
F = system["n_frag"]
dev = Device(0)
eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
dimers, trimers = list(itertools.combinations(range(F), 2)), list(itertools.combinations(range(F), 3))
build = sharded_build(eng, dimers, trimers)


def wall(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


eng.preload()
build.step()
eager = wall(build.step, 10)
ref_H2 = build.full(*dimers[0]).clone()
ref_m = {ms: build.H3_moments[ms].clone() for ms in trimers}
out = {"workload": name, "eager_ms": 1e3 * eager}
for streams in (1, 16):
    rec = build.recorded(streams=streams)
    rec.run()
    torch.cuda.synchronize()
    same = bool(torch.equal(build.full(*dimers[0]), ref_H2) and all(torch.equal(build.H3_moments[ms], ref_m[ms]) for ms in trimers))
    out["recorded_%d_streams" % streams] = {"ms": 1e3 * wall(rec.run, 50), "graph": rec.graph is not None, "streams_used": rec.n_streams,
                                            "launches": rec.launches, "bit_identical_to_eager": same,
                                            "error": rec._launcher.streams_error or rec._launcher.graph_error}
print(json.dumps(out))
