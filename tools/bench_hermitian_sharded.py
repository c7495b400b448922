"""Row-sharded hermitian get_xr_H on N GPUs of one box (hermitian/distributed.py): one process per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_hermitian_sharded.py <xr_order> <config>

times get_xr_H(..., shard=(rank, world)) with densities resident in HBM (each rank holds only its bra slab of fragment 0),
max over ranks, and checks rank 0's assembled H2 against the unsharded build on rank 0.  XR_FAKE=1 runs the same script
over gloo on the TEST-ONLY NumPy device stand-in (a dry run of the host logic, no GPU).  Prints one JSON line on rank 0."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy, torch
import torch.distributed as dist
from qodeapplications_b200 import synth
from qodeapplications_b200.hermitian.get_xr_result import get_xr_H

order = int(sys.argv[1]) if len(sys.argv) > 1 else 0
name = sys.argv[2] if len(sys.argv) > 2 else "herm49"
fake = os.environ.get("XR_FAKE") == "1"
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if fake:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_xr import FakeDevice
    dev = FakeDevice()
    sync = lambda: None
else:
    from qodeapplications_b200.device import Device
    torch.cuda.set_device(local)
    dev = Device(local)
    sync = torch.cuda.synchronize
if world > 1:
    dist.init_process_group("gloo" if fake else "nccl", rank=rank, world_size=world)

ops = {0: synth.OPS_ORDER0, 1: synth.OPS_ORDER1, 2: synth.OPS_ORDER2}[order]
system = synth.make_system(name, ops=ops, with_bior=True)
charges = system["charges"]
args = ((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges])


def timed(together=True, **kw):
    get_xr_H(*args, device=dev, **kw)          # warm-up
    best = 1e30
    for _ in range(3):
        if world > 1 and together:
            dist.barrier()
        sync()
        t0 = time.perf_counter()
        H = get_xr_H(*args, device=dev, **kw)
        sync()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=None if fake else dev.torch_device)
        if world > 1 and together:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best, H


t_shard, (H1, H2) = timed(shard=(rank, world))
if rank == 0:
    t_one, (R1, R2) = timed(together=False)
    err = float(numpy.abs(H2 - R2).max() / numpy.abs(R2).max())
    print(json.dumps({"what": "get_xr_H row-sharded", "config": name, "xr_order": order, "n_gpus": world, "fake_device": fake,
                      "seconds_sharded_max_over_ranks": t_shard, "seconds_unsharded_rank0": t_one, "dim_H2": int(H2.shape[0]),
                      "max_rel_diff_sharded_vs_unsharded": err}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
