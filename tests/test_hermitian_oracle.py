"""The hermitian-path CPU oracle against vectors produced by the REFERENCE's own code
(oracle/gen_golden.py: get_xr_H and diagrammatic_expansion.blocks run unmodified through the qode shim)."""
import os
import numpy
import pytest

from qodeapplications_b200 import synth
from oracle import hermitian_oracle as ho

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=1e-10):
    scale = max(numpy.abs(b).max(), 1e-300)
    assert numpy.abs(numpy.asarray(a) - b).max() <= tol * scale, (numpy.abs(a - b).max(), scale)


@pytest.mark.parametrize("fixture,n_labels,ops,name", [("hermitian_toy_blocks.npz", 32, synth.OPS_ORDER1, "toy"),
                                                       ("hermitian_toy_blocks2.npz", 28, synth.OPS_ORDER2, "toy"),
                                                       ("hermitian_toy4_blocks34.npz", 56, synth.OPS_ORDER4, "toy4")])
def test_every_diagram_block_matches_reference(fixture, n_labels, ops, name):
    toy = synth.make_system(name, ops=ops, with_bior=True)       # the fixture was generated from exactly this draw
    g = numpy.load(os.path.join(GOLDEN, fixture))
    dens, symm = toy["densities"], toy["symm"]
    ints = ho.integrals(symm.S, symm.T, symm.U, symm.V)
    seen = set()
    for key in g.files:
        if key == "input_sha256":
            continue
        parts = key.split("|")
        label = parts[0]
        seen.add(label)
        if len(parts) == 3:
            m = int(parts[1])
            ci, cj = (int(x) for x in parts[2].split(","))
            _close(ho.monomer_block(label, dens, ints, m, ci, cj), g[key])
        else:
            ci0, ci1, cj0, cj1 = (int(x) for x in parts[1].split(","))
            got = ho.dimer_block(label, dens, ints, (0, 1), ((ci0, cj0), (ci1, cj1)))
            assert got is not None, key
            _close(got, g[key])
    assert len(seen) == n_labels       # orders 0-1: 3 one-fragment + 29 two-fragment diagrams; order 2: 28


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)])
def test_get_xr_H_matches_reference(order, ops):
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    system = synth.make_system("toy", ops=ops, with_bior=True)
    charges = system["charges"]
    H1, H2 = ho.get_xr_H(system["symm"], system["bior"], system["densities"], order, [charges, charges])
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"], 1e-9 if order else 1e-10)
