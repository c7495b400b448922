"""The multi-rank builds of tests/test_distributed_cpu.py over NCCL on real GPUs (one process per GPU): needs a box with
at least two of them (`gpurun --gpus 2`), skipped otherwise.  Same workers, same checks against the reference's golden vectors.
The file sorts last so that the single-GPU tier has finished before two processes share the box."""
import itertools
import os
import numpy
import pytest
import torch
import torch.multiprocessing as mp

import test_distributed_cpu as cpu

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box (NCCL, one process per GPU)")]


@pytest.mark.timeout(600)
def test_two_rank_sharded_build_over_nccl(tmp_path):
    world = 2
    mp.spawn(cpu._worker, args=(world, cpu._free_port(), str(tmp_path), True), nprocs=world, join=True)
    g = numpy.load(os.path.join(cpu.GOLDEN, "general_toy3.npz"))
    ref3 = numpy.zeros(tuple(g["H3_012_shape"]))
    ref3[g["H3_012_rows"], g["H3_012_cols"]] = g["H3_012_vals"]
    for rank in range(world):
        out = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        for m1, m2 in itertools.combinations(range(3), 2):
            ref = g["H2_%d%d" % (m1, m2)]
            assert numpy.abs(out["H2_%d%d" % (m1, m2)] - ref).max() <= 1e-10 * numpy.abs(ref).max()
            for mode in ("nccl", "ce"):       # both assemble modes by name, three steps each on the same buffers
                assert numpy.array_equal(out["H2_%d%d_%s" % (m1, m2, mode)], out["H2_%d%d" % (m1, m2)])
        assert str(out["assemble_nccl"]) == "nccl"
        print("assemble modes used on this box:", str(out["assemble"]), str(out["assemble_ce"]))
        assert abs(out["moments"][1] - (ref3 ** 2).sum()) <= 1e-10 * (ref3 ** 2).sum()
        ref2 = g["H2_02"]
        assert abs(out["dimer_moments"][1] - (ref2 ** 2).sum()) <= 1e-11 * (ref2 ** 2).sum()
        assert numpy.allclose(out["dimer_moments_held"], out["dimer_moments"], rtol=1e-12, atol=0)
        assert numpy.allclose(out["dimer_moments_sector"], out["dimer_moments"], rtol=1e-12, atol=0)


@pytest.mark.timeout(600)
def test_row_sharded_get_xr_H_over_nccl(tmp_path):
    world = 2
    mp.spawn(cpu._hermitian_worker, args=(world, cpu._free_port(), str(tmp_path), True), nprocs=world, join=True)
    for order in (0, 1, 2):
        g = numpy.load(os.path.join(cpu.GOLDEN, "hermitian_toy_order%d.npz" % order))
        for rank in range(world):
            out = numpy.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
            for key in ("H1_0", "H1_1", "H2"):
                got, ref = out["%s_order%d" % (key, order)], g[key]
                assert got.shape == ref.shape
                assert numpy.abs(got - ref).max() <= 1e-10 * numpy.abs(ref).max(), (order, rank, key)
