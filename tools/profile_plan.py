"""Where the GPU time of a recorded get_xr_H goes: every call of the recorded launch sequence (hermitian/plan.py) is
re-issued between two CUDA events; times are grouped by kernel kind and operand size.
    python tools/profile_plan.py cfg1 [top]       -> one JSON line (rank-ordered groups, ms and share)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from qodeapplications_b200.device import Device
from qodeapplications_b200.hermitian.plan import plan

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
w = bench.WORKLOADS[name]
system = bench.hermitian_system(name)
ch = system["charges"]
dev = Device(0)
build = plan((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], w["xr_order"], [ch, ch], device=dev, graph=False)
build.run()
torch.cuda.synchronize()
events = []
for call, args, kwargs in build.trace:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(dev.ctx, *args, **kwargs)
    e1.record()
    events.append((call.__name__, args, e0, e1))
torch.cuda.synchronize()
groups, total = {}, 0.0
for kind, args, e0, e1 in events:
    ms = e0.elapsed_time(e1)
    total += ms
    if kind == "gemm_scatter":
        M, N, K = args[0], args[1], args[2]
        key = "gemm M=%d N=%d K=%d%s" % (M, N, K, " +tables" if args[9] is not None else "")
        nbytes = 8.0 * (M * K + N * K + M * N)
    elif kind == "permute_copy":
        shape = list(args[2])
        key = "permute %s strides %s" % (shape, list(args[3]))
        n = 1
        for s in shape:
            n *= s
        nbytes = 16.0 * n
    elif kind == "memset_zero":
        key, nbytes = "memset %d B" % args[1], float(args[1])
    else:
        key, nbytes = kind, 0.0
    g = groups.setdefault(key, {"calls": 0, "ms": 0.0, "bytes": 0.0})
    g["calls"] += 1
    g["ms"] += ms
    g["bytes"] += nbytes
# the critical path of the recorded sequence under its real data dependencies (hermitian/schedule.py): what a perfect
# multi-stream placement could reach, with these per-call times
from qodeapplications_b200.hermitian import schedule
deps = schedule.dependencies(build.trace)
finish, longest = [], 0.0
for i, (kind, args, e0, e1) in enumerate(events):
    start = max([finish[j] for j in deps[i]] + [0.0])
    finish.append(start + e0.elapsed_time(e1))
    longest = max(longest, finish[-1])
rows = sorted(groups.items(), key=lambda kv: -kv[1]["ms"])
by_kind = {}
for kind, args, e0, e1 in events:
    k = by_kind.setdefault(kind, {"calls": 0, "ms": 0.0})
    k["calls"] += 1
    k["ms"] += e0.elapsed_time(e1)
print(json.dumps({"workload": name, "calls": len(events), "total_ms_between_events": total, "critical_path_ms": longest,
                  "dependency_edges": sum(len(d) for d in deps), "by_kind": by_kind,
                  "top": [dict(what=k, calls=v["calls"], ms=round(v["ms"], 3), gbs=round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] else None)
                          for k, v in rows[:top]]}))
