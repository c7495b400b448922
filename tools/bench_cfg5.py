"""BASELINE configs[4]: synthetic dimer stress (48 spin orbitals/fragment, 1000 states/fragment), sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_cfg5.py --scale 1.0 --steps 2 --warmup 1
    python tools/bench_cfg5.py --scale 0.25            # a quarter of the states per fragment: fits one GPU

The dimer block H2[0][1] has (1000*1000)^2 = 1e12 elements (8 TB) and is never stored: every rank draws ITS bra slab of the
densities on the device (~110 GB per GPU at full size, synth.make_device_slab_fragments), builds its rows of the class
factors, the fragment-2 factor slabs are all-gathered over NCCL, and xr_gemm_reduce streams the slab of the block through
the FP64 tensor pipe into per-class (sum, sum of squares).  One step = all five charge-transfer classes, factor build and
exchange included.  The moments are checked against the Gram-matrix identity
    sum_ab (A B^T)_ab^2 = <A^T A, B^T B>,   sum_ab (A B^T)_ab = (sum_a A_a) . (sum_b B_b)
evaluated with torch.matmul (cuBLAS) -- an independent path used only as the checker; the factor builders themselves are
checked against the reference at small sizes by tests/test_general_gpu.py.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

import numpy
import torch
import torch.distributed as dist

from qodeapplications_b200 import synth
from qodeapplications_b200.device import Device
from qodeapplications_b200.general.build_H import build_matrix_elements
from qodeapplications_b200.general.distributed import balanced_shard


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of configs[4]'s states per fragment")
    ap.add_argument("--n-orb", type=int, default=48)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = Device(local)

    full = synth.CONFIGS["cfg5"]["n_states"]
    n_states = {chg: max(2, int(round(n * args.scale))) for chg, n in full.items()}
    n = args.n_orb
    dim = sum(n_states.values())
    t0 = time.time()
    symm, nuc = synth.make_integrals(2, n, numpy.random.default_rng(5))
    mine = balanced_shard(n_states, rank, world)       # an equal share of every charge sector of both fragments
    held = {0: mine, 1: mine}
    frags = synth.make_device_slab_fragments(2, n, n_states, held, dev.torch_device, seed=5)
    torch.cuda.synchronize()
    density_bytes = sum(t.numel() * 8 for f in frags for blocks in f.rho.values() for t in blocks.values())
    setup_s = time.time() - t0
    eng = build_matrix_elements(frags, symm, nuc, device=dev, held=held)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        m = eng.H2_moments_device(0, 1, shard=(rank, world))
        if world > 1:
            dist.all_reduce(m)
        return m

    # ---- check: Gram identity on this rank's slab of every class
    check = None
    if not args.no_check:
        expected = torch.zeros((5, 2), dtype=torch.float64, device=dev.torch_device)
        def inspect(d1, A, B, P1, P2, K):
            a, b = A[:P1, :K], B[:P2, :K]
            expected[d1 + 2, 1] = ((a.T @ a) * (b.T @ b)).sum()
            expected[d1 + 2, 0] = a.sum(dim=0) @ b.sum(dim=0)
        got = eng.H2_moments_device(0, 1, shard=(rank, world), inspect=inspect)
        err_sq = float(((got[:, 1] - expected[:, 1]).abs() / expected[:, 1].clamp_min(1e-300)).max())
        scale = float(got[:, 1].sum().sqrt())                # |sum| is bounded by sqrt(count * sumsq); compare on that scale
        err_sum = float((got[:, 0] - expected[:, 0]).abs().max()) / max(scale, 1e-300)
        errs = torch.tensor([err_sq, err_sum], dtype=torch.float64, device=dev.torch_device)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        check = {"identity": "sum C^2 = <A^T A, B^T B>; sum C = (sum_a A_a).(sum_b B_b), cuBLAS as checker",
                 "max_rel_err_sumsq": float(errs[0]), "max_err_sum_over_norm": float(errs[1])}
        del expected, got

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = dev.ctx.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        moments = step()
    stop.record()
    barrier()
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev.torch_device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = dev.ctx.launch_count() - launches0
    seconds = float(ms[0]) / 1e3 / args.steps

    # one more (untimed) step with an event before every class's stream: where the step's time goes on this rank
    marks, eng.profile = [], []
    def mark(d1, A, B, P1, P2, K):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((d1, e))
    begin = torch.cuda.Event(enable_timing=True)
    begin.record()
    eng.H2_moments_device(0, 1, shard=(rank, world), inspect=mark)
    torch.cuda.synchronize()
    phases, last = {}, begin
    for (d1, e), (label, flops, nbytes, e0, e1) in zip(marks, eng.profile):
        phases["d%+d" % d1] = {"factors_and_exchange_ms": last.elapsed_time(e), "stream_ms": e0.elapsed_time(e1),
                               "stream_tflops": flops / e0.elapsed_time(e1) / 1e9}
        last = e1
    eng.profile = None

    # algorithmic flops of the streamed block (all ranks) + the factor contractions
    pairs = lambda d: sum(n_states[c] * n_states[c - d] for c in n_states if c - d in n_states)
    P0, P1, P2 = pairs(0), pairs(1), pairs(2)
    stream_flops = 2.0 * (P0 * P0 * (n * n + 2) + 2 * P1 * P1 * (2 * n) + 2 * P2 * P2 * (n * n))
    factor_flops = 2.0 * (P0 * (n * n) * (n * n + 1) + 2 * P2 * (n * n) * (n * n) + 2 * 2 * P1 * (n * n * n * n + n * n))
    peak_mem = torch.cuda.max_memory_allocated() / 1e9
    if rank == 0:
        out = {
            "metric": "streamed dimer H2 build throughput (FP64)", "value": (stream_flops + factor_flops) / seconds / 1e12,
            "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": seconds * 1e3,
            "higher_is_better": True, "scaling": "strong", "dtype": "f64", "data": "synthetic (drawn on device, per-rank bra slabs)",
            "config": {"workload": "cfg5" if args.scale == 1.0 else "cfg5 x %g states" % args.scale, "n_orb": n,
                       "n_states": {str(k): v for k, v in n_states.items()}, "dim_per_fragment": dim,
                       "block_elements": float(dim) ** 4, "pairs_per_class": {"0": P0, "+-1": P1, "+-2": P2}},
            "stream_tflops_alg": stream_flops / 1e12, "factor_tflops_alg": factor_flops / 1e12,
            "moments": {"sum": float(moments[:, 0].sum()), "sumsq": float(moments[:, 1].sum())},
            "check": check, "gpu_launches": launches, "density_GB_per_gpu": density_bytes / 1e9, "peak_mem_GB": peak_mem,
            "setup_s": setup_s, "rank0_phases": phases,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
