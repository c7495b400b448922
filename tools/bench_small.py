"""Secondary measurement for the small BASELINE configs (cfg1/2: Be2 shapes, cfg3: Be3 chain): general-XRCC
H1 + H2 (+ dense H3) built by the GPU drop-in vs the reference's per-element CPU path (its C kernels through
the element-level port, 1 core, sampled + extrapolated).  Prints one JSON line per config."""
import itertools, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_H import build_matrix_elements
from oracle import cpu_baseline

dev = Device(0)
for name in sys.argv[1:] or ["cfg1", "cfg3"]:
    system = synth.make_system(name)
    F = system["n_frag"]
    dimers, trimers = list(itertools.combinations(range(F), 2)), list(itertools.combinations(range(F), 3))
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    def build():
        eng.drop_caches(densities=True)
        out = [eng.H1(m) for m in range(F)] + [eng.H2(*d) for d in dimers]
        mom = [eng.H3_moments(*t) for t in trimers]
        return out, mom
    build(); torch.cuda.synchronize()
    times = []
    for _ in range(5):
        t0 = time.perf_counter(); build(); torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    cpu_baseline.prepare(system)
    counts = eng.element_counts(dimers, trimers)
    sample = cpu_baseline.make_sample(system, 400)
    secs = cpu_baseline.time_sample(sample, 1)
    cpu_full = cpu_baseline.extrapolate(secs, sample, counts)
    flops, _ = eng.algorithmic_flops(dimers, trimers)
    print(json.dumps({"config": name, "what": "general-XRCC H1+H2 (dense, downloaded) + H3 (streamed moments), host inputs re-uploaded every build",
                      "gpu_seconds_best": min(times), "gpu_seconds_all": times, "cpu_reference_seconds_extrapolated_1core": cpu_full,
                      "speedup_vs_reference_1core": cpu_full / min(times), "elements": int(sum(counts.values())),
                      "algorithmic_gflop": flops / 1e9}))
