#!/usr/bin/env python
"""Benchmark of the XR Hamiltonian build (BASELINE.json metric: H-build time and FP64 TFLOP/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl xr|reference] [--workload cfg4]

Workload cfg4 (the configuration north_star's target names): 4 fragments x 200 states, n = 18.  One STEP = one pass of the
build over every H1, every dimer block H2[m1][m2] materialised dense in HBM (assembled on every rank by an NCCL all-gather
when N > 1), and ONE of the four trimer blocks H3[m1][m2][m3], taken round-robin -- formed element by element and
streamed into the on-chip moment reducer (1e13 elements per trimer: they cannot be stored anywhere).  Four consecutive
steps therefore contain one full set of trimers, and `build_time_s` = 4 x the step time minus the three repeated dimer
phases is the time of one complete H build (the dimer phase is < 0.5 % of a step).  Round 1 timed whole builds per step (46 s each),
which no driver budget could hold 25 of.  Work is sharded over ranks by bra-state slabs (dimers) and by leading
pair-index slabs (trimers): total work is fixed, so scaling is "strong".  `value` = algorithmic FP64 flops of the
factored algorithm of the timed steps / device time, inputs resident in HBM; `e2e` = the same through
the public Python API from pinned host buffers, uploads and result downloads inside the timed region.

`--impl reference` times the reference's own per-element CPU path (its compiled H_contractions.c, one call per element
under the control flow of general-XRCC/build_H.py) on a bounded sample with all host cores.

Other workloads print the same JSON schema: cfg4-half / cfg3 (general path), cfg5 (streamed 1000-state dimer, scaled to
the GPUs present), cfg1 / cfg2 / herm100 (hermitian-XRCC get_xr_H).
"""
import argparse
import itertools
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "xr_h_build_fp64_tflops"
UNIT = "TFLOP/s"

WORKLOADS = {
    # name: (synth config, description)
    "cfg4": dict(kind="general", n_frag=4, n_orb=18, n_states={0: 96, +1: 34, -1: 70}, seed=4,
                 text="synthetic 4-fragment Be chain, 200 states/fragment (96/34/70), n=18 spin orbitals: "
                      "4 H1 + 6 dimer H2 (dense, 40000^2 each) + 4 trimer H3 (1.06e13 elements each, streamed)"),
    "cfg4-half": dict(kind="general", n_frag=4, n_orb=18, n_states={0: 48, +1: 17, -1: 35}, seed=4,
                      text="development size: cfg4 with 100 states/fragment"),
    "cfg3": dict(kind="general", n_frag=3, n_orb=18, n_states={0: 11, +1: 4, -1: 8}, seed=3,
                 text="Be3 chain shapes (parity-test size): 3 H1 + 3 H2 + 1 H3"),
}


def config_of(workload):
    """the `config` object of the JSON line: identical in the xr and the reference arm"""
    w = WORKLOADS[workload]
    cfg = {"workload": workload, "description": w["text"]}
    if w["kind"] == "general":
        n_tri = len(list(itertools.combinations(range(w["n_frag"]), 3)))
        cfg["step"] = ("every H1 + every dimer H2 + one of the %d trimer H3 (round-robin): %d consecutive steps = one full build"
                       % (n_tri, max(1, n_tri)))
        cfg["sharding"] = "dimers: bra-state slabs of fragment m1 + NCCL all-gather of H2; trimers: leading pair-index slabs, no collective"
        cfg["cache"] = "inputs_larger_than_L2 (4 GB of densities read, 77 GB of H2 written per step at cfg4)"
    return cfg


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="xr", choices=["xr", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the dimer-phase / gather-overlap measurements after the timed region")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sample", type=int, default=0, help="CPU sample elements per class (0 = default)")
    return ap.parse_args()


def make_system(workload):
    from qodeapplications_b200 import synth
    w = WORKLOADS[workload]
    return synth.make_system(n_frag=w["n_frag"], n_orb=w["n_orb"], n_states=w["n_states"], seed=w["seed"],
                             ops=synth.OPS_GENERAL, general_ccaa="random")


# ------------------------------------------------------------------------------------ clocks

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, power) if p > 0.5 * max(power)) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ reference arm

REFERENCE_KIND_DETAIL = ("the reference's own C kernels (general-XRCC/H_contractions.c compiled -O2 where it lies: oracle/_ref) "
                         "called once per matrix element under the per-element control flow of general-XRCC/build_H.py:42-188 "
                         "as restated in oracle/general_oracle.element_oracle (the Python driver is a port; /root/reference does "
                         "not exist on the GPU box)")


def run_reference(args):
    """CPU arm: rank 0 only; no CUDA is touched in this process."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if WORKLOADS[args.workload]["kind"] != "general":
        raise SystemExit("--impl reference is defined for the general-path workloads (cfg4, cfg4-half, cfg3)")
    from oracle import cpu_baseline
    from qodeapplications_b200.general.build_H import build_matrix_elements
    system = make_system(args.workload)
    kind = cpu_baseline.prepare(system)
    F = system["n_frag"]
    dimers = list(itertools.combinations(range(F), 2))
    trimers = list(itertools.combinations(range(F), 3))
    acct = build_matrix_elements(system["fragments"], system["symm"], system["nuc"])     # accounting only (no GPU use)
    total_flops, _ = acct.algorithmic_flops(dimers, trimers)
    counts = acct.element_counts(dimers, trimers)
    cores = os.cpu_count() or 1
    per_class = args.sample or 250 * cores
    sample = cpu_baseline.make_sample(system, per_class)
    pool = cpu_baseline.make_pool(cores) if cores > 1 else None
    for _ in range(args.warmup):
        cpu_baseline.time_sample(sample, cores, pool)
    t0 = time.perf_counter()
    acc = None
    for _ in range(args.steps):
        secs = cpu_baseline.time_sample(sample, cores, pool)
        acc = secs if acc is None else {k: acc[k] + secs[k] for k in secs}
    wall = time.perf_counter() - t0
    if pool is not None:
        pool.close()
    per_step = {k: v / args.steps for k, v in acc.items()}
    full_seconds = cpu_baseline.extrapolate(per_step, sample, counts)
    value = total_flops / full_seconds / 1e12
    n_sample = sum(len(v) for v in sample.values())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload),
        "build_time_s_extrapolated": full_seconds,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "kind_detail": REFERENCE_KIND_DETAIL,
                         "sample": "%d random charge-allowed elements of each of %d classes (5 dimer, 3 trimer kinds) of "
                                   "fragments (0,1)/(0,1,2) per step, one Python call per element into the C kernels, "
                                   "Pool(%d); whole-workload time extrapolated with exact per-class element counts"
                                   % (per_class, len(sample), cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_elements_per_step": n_sample,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------- xr arm

def pin_system(system):
    """move every density / scalar array of the general-path fragments into pinned host memory"""
    import torch
    for frag in system["fragments"]:
        for op, blocks in frag.rho.items():
            for key in list(blocks):
                t = torch.from_numpy(blocks[key]).contiguous().pin_memory()
                blocks[key] = t.numpy()


def measured_peaks():
    """driver-written peaks of this pool's B200s (HBM copy GB/s); the fallback is B200_PROFILING.md's figure"""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-written copy bandwidth)"
        except Exception:
            pass
    return 6550.0, "fallback: /opt/skills/guides/B200_PROFILING.md measured copy bandwidth (MEASURED_PEAKS.json absent)"


def init_ranks(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl xr needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)
    return rank, world, local_rank


def run_general(args):
    import numpy
    import torch
    import torch.distributed as dist
    from qodeapplications_b200.device import Device
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import sharded_build, slab_bounds

    rank, world, local_rank = init_ranks(args)
    system = make_system(args.workload)
    pin_system(system)
    dev = Device(local_rank)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    F = system["n_frag"]
    dimers = list(itertools.combinations(range(F), 2))
    trimers = list(itertools.combinations(range(F), 3))
    flops_dimers, _ = eng.algorithmic_flops(dimers, ())
    flops_trimer = {ms: eng.algorithmic_flops((), [ms])[0] for ms in trimers}
    full_flops = flops_dimers + sum(flops_trimer.values())
    counts = eng.element_counts(dimers, trimers)
    dims = [len(f.state_indices) for f in system["fragments"]]
    build = sharded_build(eng, dimers, trimers, rank, world)
    cursor = [0]

    def step(**kw):
        """one bench step: all H1, all H2 (+ gathers), the next trimer of the round-robin; returns its algorithmic flops"""
        chosen = [trimers[cursor[0] % len(trimers)]] if trimers else []
        cursor[0] += 1
        build.step(trimers=chosen, **kw)
        return flops_dimers + sum(flops_trimer[ms] for ms in chosen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        """max-over-ranks device milliseconds per call of fn (events on the launch stream, barrier on both sides)"""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev.torch_device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.preload()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = dev.ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timed_flops = 0.0
    e0.record()
    for _ in range(args.steps):
        timed_flops += step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = dev.ctx.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev.torch_device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    profile, eng.profile = eng.profile, None

    # per-kernel-class device time from the events recorded around each launch (this rank)
    per_label = {}
    for label, flops, nbytes, s, e in profile:
        t, f, b, c = per_label.get(label, (0.0, 0.0, 0.0, 0))
        per_label[label] = (t + s.elapsed_time(e), f + flops, b + nbytes, c + 1)
    tri = [v for label, v in per_label.items() if label.startswith("trimer_stream")]
    tri_ms, tri_flops, _, tri_count = (sum(x[i] for x in tri) for i in range(4)) if tri else (0.0, 0.0, 0.0, 0)

    # ---- roofline denominators, measured in THIS run: FP64 tensor pipe (DMMA probe), HBM (driver-written file)
    fp64_peak = dev.ctx.probe_fp64(0.5)
    hbm_peak, hbm_source = measured_peaks()

    # ---- the part of the build that produces a usable H, on its own: the dimer phase with and without the assemble step
    extras = {}
    if not args.no_extras:
        def dimers_only(gather, overlap=True):
            return lambda: build.step(gather=gather, trimers=[], overlap=overlap)
        compute_ms = timed(dimers_only(False), 3)
        extras["dimer_phase"] = {"compute_ms": compute_ms, "h2_bytes_written_all_ranks": sum(8.0 * (dims[a] * dims[b]) ** 2 for a, b in dimers),
                                 "note": "every H1 + all %d H2 of one build, this rank's bra slabs; zero fill of the dense blocks included" % len(dimers)}
        if world > 1:
            with_gather_ms = timed(dimers_only(True, overlap=False), 3)
            recv = build.gather_bytes()
            extras["dimer_phase"].update({
                "with_gather_ms": with_gather_ms, "gather_bytes_received_per_rank": recv,
                "gather_gbs_per_rank": recv / max(with_gather_ms - compute_ms, 1e-6) / 1e6,
                "nvlink_peak_gbs": 900.0, "limiting_collective": "NCCL all_gather_into_tensor of the H2 bra slabs (8 B/element in, 2K flop/element: gather-bound at n = 18)"})
            one = lambda **kw: (lambda: step(**kw))
            def four(**kw):          # a whole round-robin cycle, so the three variants time the same trimers
                return timed(one(**kw), len(trimers) or 1)
            extras["step_ms_by_assemble_mode"] = {"no_gather": four(gather=False), "gather_overlapped": four(gather=True, overlap=True),
                                                  "gather_blocking": four(gather=True, overlap=False)}

    # trimer moments summed over ranks (the only "exchange" the trimer path has: 24 doubles per trimer)
    for ms_ in trimers:          # make sure every trimer has a result even when steps < len(trimers)
        if ms_ not in build.H3_moments:
            build.step(gather=False, trimers=[ms_])
    moments = build.reduced_moments()

    # ---- end-to-end: pinned host inputs -> upload -> build -> results read back to pinned host
    e2e = None
    if not args.no_e2e:
        host_out = {}
        for m1, m2 in dimers:
            lo, hi, per = slab_bounds(dims[m1], rank, world)
            key = ((hi - lo) * dims[m2], dims[m1] * dims[m2])
            if key not in host_out:
                host_out[key] = torch.empty(key, dtype=torch.float64, pin_memory=True)
        side = torch.cuda.Stream(device=dev.torch_device)
        copied = [0]
        def read_back_dimers():
            # this rank's rows of the dimer blocks are final once their launches are queued: read them back on a second
            # stream while the trimer phase (99 % of the step) runs on the main one
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(side):
                side.wait_event(ready)
                for m1, m2 in dimers:
                    mine = build.my_rows(m1, m2)
                    host_out[tuple(mine.shape)].copy_(mine, non_blocking=True)
                    copied[0] += mine.numel() * 8
        def e2e_step():
            eng.drop_caches(densities=True)
            dev.h2d_bytes = dev.d2h_bytes = 0
            copied[0] = 0
            chosen = [trimers[cursor[0] % len(trimers)]] if trimers else []
            flops = step(gather=True, after_dimers=read_back_dimers)
            d2h = copied[0]
            for m in range(F):
                d2h += build.H1[m].numel() * 8
                build.H1[m].cpu()
            side.synchronize()
            for ms_ in chosen:
                d2h += build.H3_moments[ms_].numel() * 8
                build.H3_moments[ms_].cpu()
            return flops, dev.h2d_bytes, d2h
        barrier()
        e2e_steps = max(1, args.e2e_steps)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        wall0 = time.perf_counter()
        t0.record()
        e2e_flops = h2d = d2h = 0
        for _ in range(e2e_steps):
            f_, h_, d_ = e2e_step()
            e2e_flops += f_; h2d += h_; d2h += d_
        t1.record()
        barrier()
        wall = time.perf_counter() - wall0
        tms = torch.tensor([max(t0.elapsed_time(t1), 1e3 * wall)], device=dev.torch_device, dtype=torch.float64)
        byt = torch.tensor([h2d, d2h], device=dev.torch_device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            dist.all_reduce(byt)
        e2e = {"value": e2e_flops / (float(tms.item()) * 1e-3) / 1e12, "unit": UNIT,
               "h2d_bytes_per_step": int(byt[0].item() / e2e_steps), "d2h_bytes_per_step": int(byt[1].item() / e2e_steps),
               "seconds_per_step": float(tms.item()) * 1e-3 / e2e_steps, "steps": e2e_steps,
               "note": "host wall clock (max over ranks) around, every step: upload of every density from pinned host memory, the "
                       "build, download of every H1, this rank's H2 slabs and the step's H3 moments; the H2 slabs are read back "
                       "on a second stream while the trimer phase runs"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (trimer_stream_kernel): FP64 tensor pipe
    traffic, traffic_source = None, None
    for name in ("r02_trimer_traffic.json", "trimer_traffic.json"):
        tpath = os.path.join(REPO, "profiles", name)
        if os.path.exists(tpath):
            rec = json.load(open(tpath))
            traffic, traffic_source = rec.get("dram_bytes_per_launch"), "profiles/%s (%s)" % (name, rec.get("kernel", "ncu --set full"))
            break
    achieved = tri_flops / (tri_ms * 1e-3) / 1e12 if tri_ms else None
    roofline = {
        "bound": "tensor", "kernel": "trimer_stream_kernel<4,2,2> (FP64 DMMA.8x8x4 + DFMA k-tail, TMA-fed)", "achieved": achieved,
        "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic,
        "traffic_source": traffic_source,
        "peak_source": "measured in this run on this device: xr_probe_fp64 (sustained DMMA.8x8x4 issue rate, 16 warps/SM, 0.5 s, CUDA "
                       "events; = 148 SMs x 64 FMA/clk x 2 x SM clock).  MEASURED_PEAKS.json has no FP64 entry and tcgen05 has no f64 "
                       "kind, so its bf16 figure does not apply; cuBLAS DGEMM 8192^3 reaches 35.5 (profiles/r01_cublas_dgemm_peak.json)",
        "launches_timed": tri_count, "avg_launch_ms": tri_ms / tri_count if tri_count else None,
        "share_of_step": tri_ms / ms_total if ms_total else None,
        "algorithmic_flops_per_launch": tri_flops / tri_count if tri_count else None,
    }

    # ---- CPU baseline: the reference's per-element path, 1 core, bounded sample (rank 0, N=1 only)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import cpu_baseline
        kind = cpu_baseline.prepare(system)
        per_class = args.sample or 20000        # ~15 s of CPU at ~90 us per element
        sample = cpu_baseline.make_sample(system, per_class)
        secs = cpu_baseline.time_sample(sample, 1)
        full_seconds = cpu_baseline.extrapolate(secs, sample, counts)
        cpu = {"value": full_flops / full_seconds / 1e12, "unit": UNIT, "cores": 1, "kind": kind, "kind_detail": REFERENCE_KIND_DETAIL,
               "sample": "%d random charge-allowed elements of each of %d classes, one Python call per element into the "
                         "reference's C kernels (-O2), %.1f s of CPU; whole-workload time (%.3g s) extrapolated with exact "
                         "per-class element counts" % (per_class, len(sample), sum(secs.values()), full_seconds),
               "us_per_element": {"%s_%s" % k: 1e6 * v / len(sample[k]) for k, v in secs.items()},
               "host_cores_available": os.cpu_count()}
        if system["n_frag"] >= 3:
            try:
                tf, rows_done, took = cpu_baseline.factored_numpy(system["n_states"], system["n_orb"], seconds=5.0)
                cpu["factored_numpy"] = {"value": tf, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                         "sample": "the factored trimer-class algorithm of the GPU path (two GEMMs + sum of squares per "
                                                   "tile) in NumPy/BLAS on all host cores: %d rows of W x %s x %s elements in %.1f s"
                                                   % (rows_done, "P(-1)", "P(+1)", took)}
            except Exception as exc:        # a secondary figure must never cost the measured line
                cpu["factored_numpy"] = {"error": repr(exc)}

    def class_record(label, t, f, b, c):
        """per-class roofline: K = n^2 classes and streamed tiles against the FP64 tensor pipe, K = 2n dimer classes against
        the HBM write bandwidth"""
        rec = {"ms": t, "launches": c, "tflops": f / (t * 1e-3) / 1e12 if t else None,
               "gbs": b / (t * 1e-3) / 1e9 if t else None}
        hbm_bound = label in ("dimer_class_d+1", "dimer_class_d-1")
        if t:
            rec["bound"] = "hbm" if hbm_bound else "tensor"
            rec["peak"] = hbm_peak if hbm_bound else fp64_peak
            rec["frac"] = (rec["gbs"] if hbm_bound else rec["tflops"]) / rec["peak"]
        return rec

    n_tri = max(1, len(trimers))
    step_ms = ms_total / args.steps
    dimer_ms = sum(t for label, (t, f, b, c) in per_label.items() if label.startswith("dimer_class")) / args.steps
    line = {
        "metric": METRIC, "value": timed_flops / (ms_total * 1e-3) / 1e12, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.workload),
        "build_time_s": (n_tri * step_ms - (n_tri - 1) * dimer_ms) * 1e-3,
        "build_time_note": "%d consecutive steps (one full round-robin of the trimers) minus the %d repeated dimer-class phases" % (n_tri, n_tri - 1),
        "algorithmic_flops_per_step": timed_flops / args.steps, "algorithmic_flops_full_build": full_flops,
        "flops_split": {"dimer": flops_dimers, "trimer": sum(flops_trimer.values())},
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        "peaks": {"fp64_tensor_tflops": fp64_peak, "hbm_gbs": hbm_peak, "hbm_source": hbm_source},
        "kernel_classes": {label: class_record(label, *v) for label, v in sorted(per_label.items())},
        "trimer_moments": {"".join(map(str, k)): v for k, v in moments.items()},
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if "NCCL_DEBUG" not in os.environ:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout unless the caller asks for more
    if args.impl == "reference":
        run_reference(args)
    else:
        {"general": run_general}[WORKLOADS[args.workload]["kind"]](args)


if __name__ == "__main__":
    main()
