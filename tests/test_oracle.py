"""The CPU oracle against the golden vectors the REFERENCE produced (oracle/gen_golden.py),
and the two C libraries (reference-compiled vs our restatement) against each other."""
import itertools
import os
import numpy
import pytest

from qodeapplications_b200 import synth
from oracle import general_oracle as go

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_SO = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libH_contractions_ref.so")


def _load(name):
    return numpy.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _h3_dense(g, key):
    shape = tuple(g[key + "_shape"])
    H = numpy.zeros(shape)
    H[g[key + "_rows"], g[key + "_cols"]] = g[key + "_vals"]
    return H


def _close(a, b, tol=1e-10):
    scale = max(numpy.abs(b).max(), 1e-300)
    assert numpy.abs(a - b).max() <= tol * scale, (numpy.abs(a - b).max(), scale)


@pytest.mark.parametrize("name", ["toy", "toy3", "toy5", "toyh", "toyh3"])
def test_block_oracle_matches_reference_golden(name):
    g = _load("general_%s.npz" % name)
    system = synth.make_system(name)
    frags, ints, nuc = system["fragments"], system["symm"], system["nuc"]
    F = system["n_frag"]
    for m in range(F):
        _close(go.block_monomer(frags, ints, nuc, m), g["H1_%d" % m])
    for m1, m2 in itertools.combinations(range(F), 2):
        _close(go.block_dimer(frags, ints, nuc, m1, m2), g["H2_%d%d" % (m1, m2)])
    for ms in itertools.combinations(range(F), 3):
        ref = _h3_dense(g, "H3_%d%d%d" % ms)
        assert numpy.count_nonzero(ref) > 1000
        _close(go.block_trimer(frags, ints, ms), ref)


@pytest.mark.parametrize("which", ["port", "ref"])
def test_element_oracle_matches_reference_golden(which):
    if which == "ref" and not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    g = _load("general_toy3.npz")
    system = synth.make_system("toy3")
    frags = system["fragments"]
    eo = go.element_oracle(frags, system["symm"], system["nuc"], go.c_contractions(which))
    st = [f.state_indices for f in frags]
    H1 = numpy.array([[eo.monomer(1, I, J) for J in st[1]] for I in st[1]])
    _close(H1, g["H1_1"])
    basis = [(a, b) for a in st[0] for b in st[2]]
    H2 = numpy.array([[eo.dimer((0, 2), I, J) for J in basis] for I in basis], dtype=float)
    _close(H2, g["H2_02"])
    ref3 = _h3_dense(g, "H3_012")
    basis3 = [(a, b, c) for a in st[0] for b in st[1] for c in st[2]]
    rng = numpy.random.default_rng(0)
    rows = rng.choice(len(basis3), size=24, replace=False)
    for i in rows:
        row = numpy.array([eo.trimer((0, 1, 2), basis3[i], J) for J in basis3], dtype=float)
        _close(row, ref3[i], tol=1e-10 * max(1., numpy.abs(ref3).max() / max(numpy.abs(ref3[i]).max(), 1e-30)))


def test_c_port_matches_reference_c():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    ref, port = go.c_contractions("ref"), go.c_contractions("port")
    rng = numpy.random.default_rng(5)
    n1, n2, n3 = 5, 4, 3
    r = lambda *s: rng.standard_normal(s)
    cases = [
        ("monomer", (n1, r(n1, n1), r(n1, n1, n1, n1), r(n1, n1), r(n1, n1, n1, n1))),
        ("monomer_1e", (n1, r(n1, n1), r(n1, n1))),
        ("monomer_2e", (n1, r(n1, n1, n1, n1), r(n1, n1, n1, n1))),
        ("monomer_extPot", (n1, r(n1, n1), r(n1, n1))),
        ("dimer_2min2pls", (n1, n2, r(n1, n1), r(n2, n2), r(n1, n1, n2, n2))),
        ("dimer_1min1pls_1e", (n1, n2, r(n1), r(n2), r(n1, n2))),
        ("dimer_1min1pls_2e", (n1, n2, r(n1), r(n1, n1, n1), r(n2), r(n2, n2, n2), r(n1, n1, n1, n2), r(n1, n2, n2, n2))),
        ("dimer_ExEx", (n1, n2, r(n1, n1), r(n2, n2), r(n1, n2, n1, n2))),
        ("trimer_2min1pls1pls", (n1, n2, n3, r(n1, n1), r(n2), r(n3), r(n1, n1, n2, n3))),
        ("trimer_2pls1min1min", (n1, n2, n3, r(n1, n1), r(n2), r(n3), r(n2, n3, n1, n1))),
        ("trimer_Ex1min1pls", (n1, n2, n3, r(n1, n1), r(n2), r(n3), r(n1, n2, n1, n3))),
    ]
    for name, args in cases:
        a, b = getattr(ref, name)(*args), getattr(port, name)(*args)
        assert abs(a - b) <= 1e-12 * max(1., abs(a)), (name, a, b)


def test_trimer_class_moments():
    rng = numpy.random.default_rng(3)
    W, b, c = rng.standard_normal((7, 4, 4)), rng.standard_normal((5, 4)), rng.standard_normal((6, 4))
    T = numpy.einsum("ars,br,cs->abc", W, b, c)
    total, sumsq = go.trimer_class_moments(W, b, c)
    assert abs(total - T.sum()) < 1e-10 * abs(T).sum()
    assert abs(sumsq - (T * T).sum()) < 1e-10 * (T * T).sum()
