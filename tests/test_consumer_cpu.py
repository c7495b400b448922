"""Consumer of the build (supersystem-matrix expansion, general-XRCC/hamiltonian.py): the oracle restatement against the
reference's own outputs (tests/golden/supersystem_hmat.npz, oracle/gen_golden_hmat.py), and the product's host logic
(offset tables, transition mask plumbing) on the TEST-ONLY NumPy device stand-in."""
import os
import sys
import numpy
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLDEN = os.path.join(HERE, "golden")

from oracle import general_oracle as go
from gen_golden_hmat import random_blocks, toy3_blocks
from fake_xr import FakeDevice

CASES = ["rand4x3", "rand3x4", "herm3x4", "herm2x5"]


def _case(tag):
    g = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))
    seed, F, spf, trimers = (int(x) for x in g[tag + "_meta"])
    return random_blocks(seed, F, spf, bool(trimers)), F, spf, g[tag]


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_braket_loops(tag):
    H, F, spf, ref = _case(tag)
    assert numpy.abs(go.supersystem_matrix([spf] * F, H) - ref).max() <= 1e-13 * numpy.abs(ref).max()


def test_oracle_matches_reference_on_toy3_blocks():
    H, spf = toy3_blocks()
    ref = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))["toy3"]
    assert numpy.abs(go.supersystem_matrix([spf] * 3, H) - ref).max() <= 1e-13 * numpy.abs(ref).max()


@pytest.mark.parametrize("tag", CASES)
def test_braket_loops_host_logic(tag):
    from qodeapplications_b200.general.hamiltonian import braket_loops
    H, F, spf, ref = _case(tag)
    Hmat = numpy.zeros((spf ** F, spf ** F))
    braket_loops(Hmat, F, spf, H, device=FakeDevice())
    assert numpy.abs(Hmat - ref).max() <= 1e-13 * numpy.abs(ref).max()
    braket_loops(Hmat, F, spf, H, device=FakeDevice())            # the reference accumulates: Hmat[I,J] += ...
    assert numpy.abs(Hmat - 2 * ref).max() <= 1e-13 * numpy.abs(ref).max()


def test_unequal_fragment_dimensions_host_logic():
    from qodeapplications_b200.general.hamiltonian import supersystem_matrix
    rng = numpy.random.default_rng(5)
    dims = [2, 4, 3]
    H1 = [rng.standard_normal((d, d)) for d in dims]
    H2 = [[rng.standard_normal((dims[M] * dims[N],) * 2) if M < N else None for N in range(3)] for M in range(3)]
    H3 = [[[rng.standard_normal((24, 24)) if (M, N, O) == (0, 1, 2) else None for O in range(3)] for N in range(3)] for M in range(3)]
    big = supersystem_matrix(dims, FakeDevice()).add_all((H1, H2, H3))
    ref = go.supersystem_matrix(dims, (H1, H2, H3))
    assert numpy.abs(big.matrix.numpy() - ref).max() <= 1e-13 * numpy.abs(ref).max()
    with pytest.raises(ValueError):
        big.add((1, 0), H2[0][1])
    with pytest.raises(ValueError):
        big.add((0, 1), H2[0][2])


@pytest.mark.parametrize("name", ["toy3", "toy5", "toyh3"])
def test_matrix_free_operator_host_logic(name):
    """y = Hmat.v from the class factors (never forming H2/H3) equals the reference's braket_loops matrix"""
    from qodeapplications_b200 import synth
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.operator import xr_operator
    system = synth.make_system(name)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=FakeDevice())
    op = xr_operator(eng)
    if name == "toy3":
        ref = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))["toy3"]
    else:       # five charge states / unequal orbital counts: the reference's blocks from the fixture, expanded by the oracle restatement
        g = numpy.load(os.path.join(GOLDEN, "general_%s.npz" % name))
        H3 = numpy.zeros(tuple(g["H3_012_shape"]))
        H3[g["H3_012_rows"], g["H3_012_cols"]] = g["H3_012_vals"]
        H = ([g["H1_%d" % m] for m in range(3)], [[g["H2_%d%d" % (M, N)] if M < N else None for N in range(3)] for M in range(3)],
             [[[H3 if (M, N, O) == (0, 1, 2) else None for O in range(3)] for N in range(3)] for M in range(3)])
        ref = go.supersystem_matrix(op.dims, H)
    got = op.dense().numpy()
    assert numpy.abs(got - ref).max() <= 1e-10 * numpy.abs(ref).max()
    # several vectors at once (trailing axis) and accumulation into a given output
    import torch
    rng = numpy.random.default_rng(3)
    v = rng.standard_normal(tuple(op.dims) + (2,))
    y = op.apply(torch.from_numpy(v)).numpy()
    D = ref.shape[0]
    assert numpy.abs(y.reshape(D, 2) - ref @ v.reshape(D, 2)).max() <= 1e-10 * numpy.abs(ref).max() * numpy.abs(v).max() * D ** 0.5


def test_hermitian_full_matrix_host_logic():
    """H2 + H1 (x) 1 + 1 (x) H1 restricted to charge-diagonal monomer blocks: hermitian-XRCC/mains/workflow.py:214-226 restated"""
    from qodeapplications_b200.hermitian.full_matrix import full_matrix
    rng = numpy.random.default_rng(6)
    charges = [[0, 1, -1], [0, -1]]
    state_dict = [{0: 3, 1: 2, -1: 2}, {0: 2, -1: 3}]
    n = [sum(state_dict[m][c] for c in charges[m]) for m in (0, 1)]
    H1 = [rng.standard_normal((n[0], n[0])), rng.standard_normal((n[1], n[1]))]
    H2 = rng.standard_normal((n[0] * n[1], n[0] * n[1]))
    slices = []
    for m in (0, 1):
        at, sl = 0, {}
        for c in charges[m]:
            sl[c] = slice(at, at + state_dict[m][c])
            at += state_dict[m][c]
        slices.append(sl)
    ref = H2.copy().reshape(n[0], n[1], n[0], n[1])
    for c0 in charges[0]:
        for c1 in charges[1]:
            ref[slices[0][c0], slices[1][c1], slices[0][c0], slices[1][c1]] += \
                numpy.einsum("ij,kl->ikjl", H1[0][slices[0][c0], slices[0][c0]], numpy.eye(state_dict[1][c1])) + \
                numpy.einsum("ij,kl->ikjl", numpy.eye(state_dict[0][c0]), H1[1][slices[1][c1], slices[1][c1]])
    ref = ref.reshape(n[0] * n[1], n[0] * n[1])
    got = full_matrix(H1, H2, state_dict, charges, device=FakeDevice())
    assert numpy.abs(got - ref).max() <= 1e-14 * numpy.abs(ref).max()
