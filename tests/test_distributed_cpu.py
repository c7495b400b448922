"""world_size-2 run of the sharded build over gloo on CPU (host logic of the multi-GPU path: bra-state
slabs, in-place all-gather of H2, summed trimer moments).  Uses the TEST-ONLY NumPy device stand-in."""
import itertools
import os
import socket
import sys
import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _device(rank, world, port, gpu):
    """process group + device of one rank: gloo on the test-only NumPy stand-in, or NCCL on GPU `rank` (test_distributed_gpu.py)"""
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if gpu:
        from qodeapplications_b200.device import Device
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        return lambda: Device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fake_xr import FakeDevice
    return FakeDevice


def _np(t):
    return t.detach().cpu().numpy()


def _worker(rank, world, port, out_dir, gpu=False):
    make_device = _device(rank, world, port, gpu)
    from qodeapplications_b200 import synth
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import sharded_build
    system = synth.make_system("toy3")
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=make_device())
    dimers = list(itertools.combinations(range(3), 2))
    build = sharded_build(eng, dimers, [(0, 1, 2)], rank, world)
    build.step(gather=True)
    moments = build.reduced_moments()
    payload = {"H2_%d%d" % k: _np(build.full(*k)) for k in dimers}
    payload["moments"] = numpy.array(moments[(0, 1, 2)])
    payload["assemble"] = numpy.array(build.assemble)
    # inputs uploaded once and broadcast: an engine whose densities arrived that way builds the same blocks
    eng_b = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=make_device())
    eng_b.preload_distributed(rank, world)
    payload["h2d_broadcast"] = numpy.array(eng_b.dev.h2d_bytes)
    payload["H2_01_broadcast_inputs"] = _np(eng_b.H2_device(0, 1))
    payload["h2d_everything"] = numpy.array(sum(numpy.asarray(b).nbytes for f in system["fragments"] for op in ("a", "c", "aa", "cc", "ca", "caa", "cca")
                                                for b in f.rho[op].values()))
    # both assemble modes, asked for by name, and steps repeated on the same buffers (the copy-engine gather's closing
    # barrier is what keeps a rank from rewriting a slab a peer still reads); on CPU / gloo "ce" falls back to the collective
    for mode in ("nccl", "ce"):
        again = sharded_build(eng, dimers, [(0, 1, 2)], rank, world, assemble=mode)
        for _ in range(3):
            again.step(gather=True)
        for k in dimers:
            payload["H2_%d%d_%s" % (k + (mode,))] = _np(again.full(*k))
        payload["assemble_" + mode] = numpy.array(again.assemble)
    # streamed dimer with the factor exchange (all-gather of the fragment-2 factor slabs)
    t = eng.H2_moments_device(0, 2, shard=(rank, world))
    dist.all_reduce(t)
    payload["dimer_moments"] = _np(t).sum(axis=0)
    # the same with each rank holding only ITS bra slab of the densities (the cfg5 input layout)
    from qodeapplications_b200.general.distributed import slab_bounds
    n_states = synth.CONFIGS["toy3"]["n_states"]
    held, frags = {}, list(system["fragments"])
    for m in (0, 2):
        lo, hi = slab_bounds(len(frags[m].state_indices), rank, world)[:2]
        held[m] = (lo, hi)
        frags[m] = synth.slab_fragment(frags[m], (lo, hi), n_states)
    eng_slab = build_matrix_elements(frags, system["symm"], system["nuc"], device=make_device(), held=held)
    t = eng_slab.H2_moments_device(0, 2, shard=(rank, world))
    dist.all_reduce(t)
    payload["dimer_moments_held"] = _np(t).sum(axis=0)
    # ... and with the balanced per-charge-sector shard as the held range
    from qodeapplications_b200.general.distributed import balanced_shard
    mine = balanced_shard(n_states, rank, world)
    frags = list(system["fragments"])
    for m in (0, 2):
        frags[m] = synth.slab_fragment(frags[m], mine, n_states)
    eng_sect = build_matrix_elements(frags, system["symm"], system["nuc"], device=make_device(), held={0: mine, 2: mine})
    t = eng_sect.H2_moments_device(0, 2, shard=(rank, world))
    dist.all_reduce(t)
    payload["dimer_moments_sector"] = _np(t).sum(axis=0)
    numpy.savez(os.path.join(out_dir, "rank%d.npz" % rank), **payload)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_build_matches_reference(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = numpy.load(os.path.join(GOLDEN, "general_toy3.npz"))
    ref3 = numpy.zeros(tuple(g["H3_012_shape"]))
    ref3[g["H3_012_rows"], g["H3_012_cols"]] = g["H3_012_vals"]
    for rank in range(world):
        out = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        for m1, m2 in itertools.combinations(range(3), 2):          # every rank holds the assembled H2
            ref = g["H2_%d%d" % (m1, m2)]
            assert numpy.abs(out["H2_%d%d" % (m1, m2)] - ref).max() <= 1e-10 * numpy.abs(ref).max()
            for mode in ("nccl", "ce"):
                assert numpy.array_equal(out["H2_%d%d_%s" % (m1, m2, mode)], out["H2_%d%d" % (m1, m2)])
        ref = g["H2_01"]
        assert numpy.abs(out["H2_01_broadcast_inputs"] - ref).max() <= 1e-10 * numpy.abs(ref).max()
        assert 0 < int(out["h2d_broadcast"]) < 0.75 * int(out["h2d_everything"]) + 4096        # this rank uploaded about half
        assert abs(out["moments"][1] - (ref3 ** 2).sum()) <= 1e-10 * (ref3 ** 2).sum()
        assert abs(out["moments"][0] - ref3.sum()) <= 1e-9 * numpy.abs(ref3).sum()
        ref2 = g["H2_02"]
        assert abs(out["dimer_moments"][1] - (ref2 ** 2).sum()) <= 1e-11 * (ref2 ** 2).sum()
        assert abs(out["dimer_moments"][0] - ref2.sum()) <= 1e-10 * numpy.abs(ref2).sum()
        assert numpy.allclose(out["dimer_moments_held"], out["dimer_moments"], rtol=1e-12, atol=0)
        assert numpy.allclose(out["dimer_moments_sector"], out["dimer_moments"], rtol=1e-12, atol=0)


def _hermitian_worker(rank, world, port, out_dir, gpu=False):
    make_device = _device(rank, world, port, gpu)
    from qodeapplications_b200 import synth
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    payload = {}
    for order, ops in ((0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)):
        system = synth.make_system("toy", ops=ops, with_bior=True)
        charges = system["charges"]
        H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges],
                          device=make_device(), shard=(rank, world))
        payload["H1_0_order%d" % order], payload["H1_1_order%d" % order], payload["H2_order%d" % order] = H1[0], H1[1], H2
    numpy.savez(os.path.join(out_dir, "rank%d.npz" % rank), **payload)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_get_xr_H_matches_reference(tmp_path, world):
    """hermitian path, one process per rank over gloo: every rank builds the rows of its slab of fragment 0's bra states
    (3 ranks over 7 states: slabs cut through charge sectors, the last one is short) and ends up with the reference's
    H1 and H2 after the all-gathers"""
    mp.spawn(_hermitian_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for order in (0, 1, 2):
        g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
        for rank in range(world):
            out = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
            for key in ("H1_0", "H1_1", "H2"):
                got, ref = out["%s_order%d" % (key, order)], g[key]
                assert got.shape == ref.shape
                assert numpy.abs(got - ref).max() <= (1e-9 if order else 1e-10) * numpy.abs(ref).max(), (order, rank, key)
