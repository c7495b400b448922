"""Density-tensor build (general-XRCC/density_tensors.c + build_density_tensors.py): the oracle restatement against the
reference's own outputs (tests/golden/density_tensors.npz from oracle/gen_golden_density.py, and the reference C compiled
into oracle/_ref when present), and the product's host logic on the TEST-ONLY NumPy device stand-in."""
import os
import sys
import numpy
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLDEN = numpy.load(os.path.join(HERE, "golden", "density_tensors.npz"))

from oracle import density_oracle as do
from gen_golden_density import CASE, inputs
from fake_xr import FakeDevice

REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libdensity_tensors_ref.so")


def _sectors(z_lists):
    for op in do.OPS:
        for bra in z_lists:
            ket = bra - do.op_dchg(op)
            if ket in z_lists:
                yield op, bra, ket


def test_oracle_restatement_is_bit_identical_to_the_reference():
    z_lists, V = inputs()
    ref = do.reference_c() if os.path.exists(REF_SO) else None
    for op, bra, ket in _sectors(z_lists):
        mine = do.tensor(op, z_lists, bra, ket, CASE["n_orbs"], CASE["n_core"])
        if op != "ccaa":
            assert numpy.array_equal(mine, GOLDEN["%s_%d_%d" % (op, bra, ket)]), op
        else:       # build_density_tensors.py:125-133 keeps sum V[p,q,r,s] ccaa[p,q,s,r]
            want = GOLDEN["ccaa_%d_%d" % (bra, ket)]
            got = numpy.einsum("pqrs,ijpqsr->ij", V, mine)
            assert numpy.abs(got - want).max() <= 1e-13 * numpy.abs(want).max()
        if ref is not None:
            assert numpy.array_equal(mine, ref.tensor(op, z_lists, bra, ket, CASE["n_orbs"], CASE["n_core"])), op


def test_config_enumeration_matches_find_config_index():
    for n_elec, n_orbs, n_core in ((4, 4, 1), (3, 4, 1), (5, 4, 1), (2, 3, 0), (6, 5, 2)):
        configs = do.all_configs(n_elec, n_orbs, n_core)
        assert len(configs) == len(set(map(tuple, configs)))
        from qodeapplications_b200.general.build_density_tensors import _check_enumeration, _masks
        _check_enumeration(configs, n_orbs, n_core)
        if len(configs) > 2:
            with pytest.raises(ValueError):
                _check_enumeration(configs[::-1], n_orbs, n_core)
            with pytest.raises(ValueError):
                _check_enumeration(configs[:-1], n_orbs, n_core)
        masks = _masks(configs)
        assert all(bin(int(m)).count("1") == n_elec for m in masks)


def test_build_density_tensors_host_logic():
    from qodeapplications_b200.general.build_density_tensors import build_density_tensors
    z_lists, V = inputs()
    densities, total = build_density_tensors(z_lists, CASE["n_orbs"], V, CASE["n_core"], device=FakeDevice())
    assert total == int(GOLDEN["total_size"])
    seen = 0
    for op, blocks in densities.items():
        for (bra, ket), nested in blocks.items():
            want = GOLDEN["%s_%d_%d" % (op, bra, ket)]
            got = numpy.array(nested)
            seen += 1
            if op == "ccaa":
                assert numpy.abs(got - want).max() <= 1e-13 * numpy.abs(want).max()
            else:
                assert numpy.array_equal(got, want), (op, bra, ket)
                assert isinstance(nested[0][0], numpy.ndarray) and nested[0][0].shape == (2 * CASE["n_orbs"],) * len(op)
    assert seen == len([k for k in GOLDEN.files if k != "total_size"])
