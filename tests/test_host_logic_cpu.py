"""Host-side orchestration of the product (class bookkeeping, sign folding, offset tables, diagram
planning) run on the TEST-ONLY NumPy device stand-in (tests/fake_xr.py) and compared with the vectors
the reference produced.  The CUDA kernels are not involved here; they are checked by the -m gpu tests."""
import itertools
import os
import re
import numpy
import pytest
import torch

from qodeapplications_b200 import synth
from fake_xr import FakeDevice

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=1e-10):
    a, b = numpy.asarray(a), numpy.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(numpy.abs(b).max(), 1e-300)
    assert numpy.abs(a - b).max() <= tol * scale, (numpy.abs(a - b).max(), scale)


@pytest.mark.parametrize("name", ["toy", "toy3", "toy5", "toyh", "toyh3"])
def test_general_blocks_host_logic(name):
    from qodeapplications_b200.general.build_H import build_matrix_elements
    g = numpy.load(os.path.join(GOLDEN, "general_%s.npz" % name))
    system = synth.make_system(name)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=FakeDevice())
    F = system["n_frag"]
    for m in range(F):
        _close(eng.H1(m), g["H1_%d" % m])
    for m1, m2 in itertools.combinations(range(F), 2):
        _close(eng.H2(m1, m2), g["H2_%d%d" % (m1, m2)])
    for ms in itertools.combinations(range(F), 3):
        key = "H3_%d%d%d" % ms
        ref = numpy.zeros(tuple(g[key + "_shape"]))
        ref[g[key + "_rows"], g[key + "_cols"]] = g[key + "_vals"]
        _close(eng.H3(*ms), ref)
        s, q = eng.H3_moments(*ms)
        assert abs(q - (ref ** 2).sum()) <= 1e-10 * (ref ** 2).sum()
        parts = [eng.H3_moments(*ms, shard=(r, 2)) for r in range(2)]
        assert abs(parts[0][1] + parts[1][1] - q) <= 1e-10 * q


@pytest.mark.parametrize("name", ["toy3", "toy5", "toyh3"])
def test_general_trimer_consumers_host_logic(name):
    """H3_elements (sampled-element consumer) and H3_sparse (screened compaction) against the dense block"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    system = synth.make_system(name)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=FakeDevice())
    dense = eng.H3(0, 1, 2)
    st = [f.state_indices for f in system["fragments"]]
    dims = [len(x) for x in st]
    D = int(numpy.prod(dims))
    rng = numpy.random.default_rng(5)
    flat = numpy.concatenate([rng.integers(D * D, size=150), rng.choice(numpy.flatnonzero(dense), size=150)])
    I = [tuple(st[k][x] for k, x in enumerate(numpy.unravel_index(int(e) // D, dims))) for e in flat]
    J = [tuple(st[k][x] for k, x in enumerate(numpy.unravel_index(int(e) % D, dims))) for e in flat]
    got = eng.H3_elements(0, 1, 2, I, J)
    _close(got, dense.reshape(-1)[flat], 1e-13)
    tau = float(numpy.median(numpy.abs(dense[dense != 0])))
    idx, val = eng.H3_sparse(0, 1, 2, tau, capacity=7)          # a tiny list: every class overflows once and is re-run
    keep = numpy.flatnonzero(numpy.abs(dense.reshape(-1)) > tau)
    assert numpy.array_equal(idx, keep)
    _close(val, dense.reshape(-1)[keep], 1e-13)
    parts = [eng.H3_sparse(0, 1, 2, tau, shard=(r, 2)) for r in range(2)]
    assert numpy.array_equal(numpy.sort(numpy.concatenate([q[0] for q in parts])), keep)
    counts = eng.H3_sparse(0, 1, 2, tau, count_only=True)
    assert sum(counts.values()) == len(keep)
    some = sorted(counts)[:5]
    assert len(eng.H3_sparse(0, 1, 2, tau, classes=some)[0]) == sum(counts[c] for c in some)
    with pytest.raises(MemoryError):
        eng.H3_sparse(0, 1, 2, 0.0, capacity=1, max_elements=3)


def test_general_bra_slabs_host_logic():
    from qodeapplications_b200.general.build_H import build_matrix_elements
    system = synth.make_system("toy")
    dev = FakeDevice()
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    full = eng.H2(0, 1)
    dim1 = len(system["fragments"][0].state_indices)
    parts = [dev.download(eng.H2_device(0, 1, bra_range=r)) for r in ((0, 3), (3, 4), (4, dim1))]
    _close(numpy.concatenate(parts, axis=0), full, 1e-13)    # (BLAS rounding differs with the slab size; the GPU test asks for equality)


def test_general_H1_of_held_bra_slabs_host_logic():
    """H1_device honours held=: each slab holder builds the rows of its bra states (ccaa scalar slab included) and the
    slabs add up to the full block; asking for states outside the slab raises instead of reading out of bounds."""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import balanced_shard
    system = synth.make_system("toy5")
    n_states = synth.CONFIGS["toy5"]["n_states"]
    frags = system["fragments"]
    full = build_matrix_elements(frags, system["symm"], system["nuc"], device=FakeDevice()).H1(1)
    dim = len(frags[1].state_indices)
    for slabs in ([(0, 2), (2, dim)], [balanced_shard(n_states, r, 3) for r in range(3)]):
        total = numpy.zeros_like(full)
        for held in slabs:
            mine = list(frags)
            mine[1] = synth.slab_fragment(frags[1], held, n_states)
            eng = build_matrix_elements(mine, system["symm"], system["nuc"], device=FakeDevice(), held={1: held})
            part = eng.H1(1)
            total += part
        _close(total, full, 1e-13)
    eng = build_matrix_elements(mine, system["symm"], system["nuc"], device=FakeDevice(), held={1: slabs[-1]})
    with pytest.raises(ValueError):
        eng._rho_rows(1, "ca", 0, 0, 0, n_states[0])


@pytest.fixture(scope="module")
def toy1():
    return synth.make_system("toy", ops=synth.OPS_ORDER2, with_bior=True)


def _hermitian_blocks(system, dev, family):
    from qodeapplications_b200.hermitian import diagrammatic_expansion
    from qodeapplications_b200.hermitian.diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams
    from qodeapplications_b200.hermitian.precontract import precontract
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    from qodeapplications_b200.hermitian.util import struct, timer
    symm, dens = system["symm"], system["densities"][:2]
    store, contractor = DeviceStore(dev), Contractor(dev)
    cache = precontract(dens, symm.S, timer(), store=store, contractor=contractor)
    ints, diagrams = {"S": (symm.S, S_diagrams), "ST": (struct(S=symm.S, T=symm.T), ST_diagrams),
                      "SU": (struct(S=symm.S, U=symm.U), SU_diagrams), "SV": (struct(S=symm.S, V=symm.V), SV_diagrams)}[family]
    return diagrammatic_expansion.blocks(densities=dens, integrals=ints, diagrams=diagrams, contract_cache=cache,
                                         timings=timer(), precon_timings=timer())


@pytest.mark.parametrize("fixture,ops,name", [("hermitian_toy_blocks.npz", synth.OPS_ORDER1, "toy"),
                                              ("hermitian_toy_blocks2.npz", synth.OPS_ORDER2, "toy"),
                                              ("hermitian_toy4_blocks34.npz", synth.OPS_ORDER4, "toy4")])
def test_hermitian_every_diagram_block_host_logic(fixture, ops, name):
    """blocks[subsystem][charges][label] -- the reference's own access pattern -- for all 116 diagrams of orders 0-4"""
    g = numpy.load(os.path.join(GOLDEN, fixture))
    toy1 = synth.make_system(name, ops=ops, with_bior=True)       # the fixture was generated from exactly this draw
    dev = FakeDevice()
    def family_of(label):
        rest = label.replace("s01", "").replace("s10", "")
        return "S" if rest == "" else "S" + rest[0].upper()
    cache = {}
    for key in g.files:
        if key == "input_sha256":
            continue
        parts = key.split("|")
        label = parts[0]
        fam = family_of(label)
        if fam not in cache:
            cache[fam] = _hermitian_blocks(toy1, dev, fam)
        blk = cache[fam]
        if len(parts) == 3:
            m = int(parts[1])
            ci, cj = (int(x) for x in parts[2].split(","))
            _close(blk[(m,)][((ci, cj),)][label], g[key])
        else:
            ci0, ci1, cj0, cj1 = (int(x) for x in parts[1].split(","))
            _close(blk[(0, 1)][((ci0, cj0), (ci1, cj1))][label], g[key])


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)])
def test_hermitian_get_xr_H_host_logic(order, ops):
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    system = synth.make_system("toy", ops=ops, with_bior=True)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges],
                      device=FakeDevice())
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"], 1e-9 if order else 1e-10)


@pytest.mark.parametrize("order", [0, 1])
def test_hermitian_get_xr_S_host_logic(order):
    """get_xr_S (get_xr_result.py:357-422): the overlap matrix the reference builds (charge-blocked, from its own modules)
    in the final ordering it meant to return"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_S
    from oracle import hermitian_oracle as ho
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_S.npz"))
    system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    charges = system["charges"]
    S2 = get_xr_S((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges],
                  device=FakeDevice())
    _close(S2, ho.reorder(g["S2_blocked_order%d" % order], system["densities"], [charges, charges]))
    if order == 0:
        assert numpy.array_equal(S2, numpy.eye(S2.shape[0]))
    with pytest.raises(NotImplementedError):
        get_xr_S((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], 2, [charges, charges], device=FakeDevice())


def test_hermitian_dimer_matrix_blocked_ordering(toy1):
    """XR_term.dimer_matrix keeps the reference's charge-blocked ordering by default"""
    from qodeapplications_b200.hermitian import XR_term
    from qodeapplications_b200.hermitian.util import timer
    from oracle import hermitian_oracle as ho
    dev = FakeDevice()
    blk = _hermitian_blocks(toy1, dev, "SV")
    charges = [(a, b) for a in toy1["charges"] for b in toy1["charges"]]
    got = XR_term.dimer_matrix(blk, {1: ["v0000"], 2: ["v0101", "v0001", "v0100", "v0011"]}, (0, 1), charges, timer())
    symm = toy1["symm"]
    ref = ho.dimer_matrix(toy1["densities"], ho.integrals(symm.S, V=symm.V), {1: ["v0000"], 2: ["v0101", "v0001", "v0100", "v0011"]}, charges)
    _close(got, ref)


@pytest.mark.parametrize("which", ["bra", "ket"])
def test_hermitian_det_variants_host_logic(which):
    """get_xr_H(..., bra_det=True) / (ket_det=True): traced two-fragment diagrams, u100 special processing, vector result
    (StateSpaceOptimizer/state_gradients.py:173,183) against the reference's own output"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    g = numpy.load(os.path.join(GOLDEN, "hermitian_det_%s.npz" % which))
    system = synth.make_det_system(which)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"], 0, [charges, charges],
                      bra_det=(which == "bra"), ket_det=(which == "ket"), device=FakeDevice())
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    assert numpy.count_nonzero(g["H2"]) > 20
    _close(H2, g["H2"])


def test_det_variants_beyond_order_0_raise_like_the_reference():
    """the reference passes bra_det / ket_det into its order-1 and order-2 branches but raises inside them
    (tests/golden/reference_det_orders.json = what the unmodified reference did, recorded by
    oracle/check_reference_det_orders.py); the drop-in refuses the same combinations and serves order 0."""
    import json
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    record = json.load(open(os.path.join(GOLDEN, "reference_det_orders.json")))
    for which in ("bra", "ket"):
        assert record["%s_det order 0" % which]["ok"]
        system = synth.make_det_system(which)
        ch = system["charges"]
        for order in (1, 2):
            assert not record["%s_det order %d" % (which, order)]["ok"]
            with pytest.raises(NotImplementedError):
                get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"], order, [ch, ch],
                         bra_det=(which == "bra"), ket_det=(which == "ket"), device=FakeDevice())


@pytest.mark.skipif(not os.path.isdir("/root/reference/hermitian-XRCC"), reason="needs the reference tree (build container only)")
def test_reference_det_orders_fixture_is_current():
    import json
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "check_reference_det_orders.py")
    out = subprocess.run([sys.executable, script], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert json.loads(out) == json.load(open(os.path.join(GOLDEN, "reference_det_orders.json")))


@pytest.mark.parametrize("name", ["toy", "toy5"])
def test_general_streamed_dimer_moments_host_logic(name):
    """H2_moments (the consumer for dimer blocks that cannot be stored) equals the moments of the dense block"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    g = numpy.load(os.path.join(GOLDEN, "general_%s.npz" % name))
    system = synth.make_system(name)
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=FakeDevice())
    ref = g["H2_01"]
    s, q = eng.H2_moments(0, 1)
    assert abs(q - (ref ** 2).sum()) <= 1e-11 * (ref ** 2).sum()
    assert abs(s - ref.sum()) <= 1e-10 * numpy.abs(ref).sum()


def test_tensornet_backend_function_table():
    """hermitian/meta_backend.py: the table XRbase/meta_backend.py:27-72 hands to tensornet.primitive_tensor_factory"""
    from qodeapplications_b200.hermitian import meta_backend as mb
    from qodeapplications_b200.hermitian import tensor as xt
    dev = FakeDevice()
    wrap = lambda a: mb.xr_wrapper(xt.DeviceTensor(dev.upload(a), dev))
    rng = numpy.random.default_rng(2)
    A, B, S = rng.standard_normal((3, 4, 5, 6)), rng.standard_normal((2, 6)), rng.standard_normal((5, 7))
    f = mb.xr_functions
    out = f.contract((wrap(A), 0, 1, "p", "q"), 2.0, (wrap(B), 2, "q"), (wrap(S), "p", 3))
    assert f.shape(out) == (3, 4, 2, 7)
    assert numpy.allclose(out.host(), 2.0 * numpy.einsum("abpq,cq,pd->abcd", A, B, S), rtol=1e-13, atol=1e-13)
    t = f.mult(-0.5, wrap(A))
    assert numpy.array_equal(t.host(), -0.5 * A)
    c = f.copy_data(t)
    f.increment(c, wrap(A))
    assert numpy.allclose(c.host(), 0.5 * A) and numpy.array_equal(t.host(), -0.5 * A)
    assert f.element(wrap(A), (1, 2, 3, 4)) == A[1, 2, 3, 4]
    s = f.contract((wrap(B), "x", "y"), (wrap(B), "x", "y"))
    assert abs(f.scalar_value(s) - (B * B).sum()) < 1e-12 and f.shape(s) == ()
    assert isinstance(f.str(s), str)


def test_xr_tensor_sums_of_products():
    """XR_tensor: `temp = zeros(); temp += A(...); temp -= A(permuted)` as frag-states/compress_frags.py:91-99 writes it"""
    from qodeapplications_b200.hermitian import XR_tensor
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    dev = FakeDevice()
    engine = (DeviceStore(dev), Contractor(dev))
    rng = numpy.random.default_rng(4)
    a, b = rng.standard_normal((3, 3, 4)), rng.standard_normal((4, 5))
    A, B = XR_tensor.init(a), XR_tensor.init(b)
    temp = XR_tensor.zeros()
    temp += A(0, 1, 2)
    temp -= A(1, 0, 2)
    assert numpy.allclose(XR_tensor.raw(temp, engine), a - a.transpose(1, 0, 2), atol=1e-15)
    expr = 0.5 * (A(0, 1, "p") @ B("p", 2)) + A(1, 0, "p") @ B("p", 2) - 2.0 * (A(0, 1, "q") @ B("q", 2))
    want = 0.5 * numpy.einsum("abp,pc->abc", a, b) + numpy.einsum("bap,pc->abc", a, b) - 2.0 * numpy.einsum("abq,qc->abc", a, b)
    assert numpy.allclose(XR_tensor.raw(expr, engine), want, rtol=1e-13, atol=1e-13)
    with pytest.raises(ValueError):
        XR_tensor.raw(A(0, 1, 2) + B(0, 1), engine)


RAGGED = [{0: 2, +1: 1, -1: 0}, {0: 1, +1: 1, -1: 1}, {0: 3, +1: 0, -1: 0}, {0: 2, +1: 0, -1: 2}, {0: 1}]


@pytest.mark.parametrize("n_states", RAGGED, ids=lambda d: "-".join("%+d:%d" % kv for kv in d.items()))
def test_general_ragged_and_empty_sectors_host_logic(n_states):
    """charge sectors with no states, one state, or a missing charge: every H1/H2/H3 element against the element-level port"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from oracle import general_oracle as go
    system = synth.make_system(n_frag=3, n_orb=4, n_states=n_states, seed=77)
    fr, ints, nuc = system["fragments"], system["symm"], system["nuc"]
    eng = build_matrix_elements(fr, ints, nuc, device=FakeDevice())
    for m in range(3):
        _close(eng.H1(m), go.block_monomer(fr, ints, nuc, m), 1e-12)
    for m1, m2 in itertools.combinations(range(3), 2):
        _close(eng.H2(m1, m2), go.block_dimer(fr, ints, nuc, m1, m2), 1e-12)
    ref = go.block_trimer(fr, ints, (0, 1, 2))
    _close(eng.H3(0, 1, 2), ref, 1e-12)
    s, q = eng.H3_moments(0, 1, 2)
    assert abs(q - (ref ** 2).sum()) <= 1e-10 * (ref ** 2).sum() and abs(s - ref.sum()) <= 1e-10 * max(numpy.abs(ref).sum(), 1.0)


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1)])
@pytest.mark.parametrize("n_states", RAGGED[:2] + RAGGED[3:4], ids=lambda d: "-".join("%+d:%d" % kv for kv in d.items()))
def test_hermitian_ragged_and_empty_sectors_host_logic(order, ops, n_states):
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from oracle import hermitian_oracle as ho
    system = synth.make_system(n_frag=2, n_orb=4, n_states=n_states, seed=78, ops=ops, with_bior=True)
    ch = system["charges"]
    args = (system["densities"][:2], order, [ch, ch])
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), *args, device=FakeDevice())
    R1, R2 = ho.get_xr_H(system["symm"], system["bior"], *args)
    _close(H1[0], R1[0])
    _close(H1[1], R1[1])
    _close(H2, R2, 1e-9 if order else 1e-10)


def factored_system(name="toy4", ops=synth.OPS_ORDER4, min_rank=5, seed=5):
    """(system with every density of >= min_rank operators given as a LAZY XR_tensor expression -- an antisymmetrised sum of
    outer products of a 'core' matrix and a 'valence' density, the form of frag-states/decomps.py:111-166 --, the same
    system with those expressions written out densely by numpy.einsum for the oracle)"""
    from qodeapplications_b200.hermitian import XR_tensor
    lazy = synth.make_system(name, ops=ops, with_bior=True)
    dense = dict(lazy, densities=[dict(rho) for rho in lazy["densities"]])
    rng = numpy.random.default_rng(seed)
    letters = "abcdefgh"
    for m, rho in enumerate(lazy["densities"]):
        for op in ops:
            k = len(op)
            if k < min_rank:
                continue
            lazy_op, dense_op = {}, {}
            for (ci, cj), t in rho[op].items():
                Ni, Nj, n = t.shape[0], t.shape[1], t.shape[2]
                core = rng.standard_normal((n, n)) / n
                val = rng.standard_normal((Ni, Nj) + (n,) * (k - 2)) * n ** (-(k - 2) / 2)
                C, V = XR_tensor.init(core), XR_tensor.init(val)
                # core on the first and last orbital axis, minus the same with the first two orbital axes exchanged
                base = C(2, k + 1) @ V(0, 1, *range(3, k + 1))
                lazy_op[(ci, cj)] = base - base(0, 1, 3, 2, *range(4, k + 2))
                orb = letters[:k]
                full = numpy.einsum("%s%s,ij%s->ij%s" % (orb[0], orb[-1], orb[1:-1], orb), core, val)
                dense_op[(ci, cj)] = full - full.swapaxes(2, 3)
            lazy["densities"][m][op], dense["densities"][m][op] = lazy_op, dense_op
    return lazy, dense


def test_hermitian_factored_densities_host_logic():
    """densities of 5..8 operators handed over as lazy sums of outer products (never formed): every diagram block of
    S-orders 1-4 that reads one, and get_xr_H at order 2, against the oracle on the densely written-out tensors"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from qodeapplications_b200.hermitian.tensor import FactoredTensor
    from qodeapplications_b200.hermitian.diagrams import specs
    from oracle import hermitian_oracle as ho
    lazy, dense = factored_system()
    dev = FakeDevice()
    charges = lazy["charges"]
    symm = lazy["symm"]
    checked = 0
    for fam, ints in (("ST", ho.integrals(symm.S, T=symm.T)), ("SU", ho.integrals(symm.S, U=symm.U)), ("SV", ho.integrals(symm.S, V=symm.V))):
        blk = _hermitian_blocks(lazy, dev, fam)
        labels = [l for l, (_, _, operands) in specs.TWO_FRAGMENT.items()
                  if l in blk._diagrams.catalog[2] and any(len(re.match(r"[ca]+", name).group(0)) >= 5 for name, _ in operands if name[0] in "ca")]
        assert labels
        for label in labels[::3]:                       # every third one keeps the CPU suite short
            for ci0, ci1, cj0, cj1 in itertools.product(charges, repeat=4):
                if ci0 + ci1 != cj0 + cj1:
                    continue
                chgs = ((ci0, cj0), (ci1, cj1))
                ref = ho.dimer_block(label, dense["densities"], ints, (0, 1), chgs)
                got = blk[(0, 1)][chgs][label]
                if ref is None:
                    assert got is None
                    continue
                _close(got, ref)
                checked += 1
    assert checked > 50
    # nothing of the size of a high-rank density was ever allocated on the device
    biggest = max(numpy.prod(t.shape) for rho in dense["densities"][:2] for op in rho if isinstance(rho[op], dict) and len(op) >= 7
                  for t in rho[op].values() if hasattr(t, "shape"))
    assert dev.largest_allocation < biggest
    lazy2, dense2 = factored_system("toy", synth.OPS_ORDER2, min_rank=5, seed=6)
    args = (2, [lazy2["charges"], lazy2["charges"]])
    H1, H2 = get_xr_H((lazy2["symm"], lazy2["bior"], lazy2["nuc"]), lazy2["densities"][:2], *args, device=FakeDevice())
    R1, R2 = ho.get_xr_H(dense2["symm"], dense2["bior"], dense2["densities"][:2], *args)
    _close(H1[0], R1[0])
    _close(H2, R2, 1e-9)


def test_xr_tensor_slicing_and_lazy_reindexing():
    """the rest of the tensornet surface the reference's callers use on this seam: slices of primitive tensors
    (StateSpaceOptimizer/orb_grads.py:67), re-indexing of products and sums without evaluation (frag-states/decomps.py:71-108)"""
    from qodeapplications_b200.hermitian import XR_tensor
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    dev = FakeDevice()
    engine = (DeviceStore(dev), Contractor(dev))
    rng = numpy.random.default_rng(8)
    v, m, d = rng.standard_normal((4, 4, 6, 4)), rng.standard_normal((4, 4, 4, 3)), rng.standard_normal((4, 4))
    V, M, Dm = XR_tensor.init(v), XR_tensor.init(m), XR_tensor.init(d)
    sl = slice(1, 5)
    got = XR_tensor.raw(V[:, :, sl, :]("p", "q", 1, "s") @ M("p", "q", "s", 0), engine)
    _close(got, numpy.einsum("pqbs,pqsa->ab", v[:, :, sl, :], m), 1e-13)
    # a product re-indexed twice, then antisymmetrised, then contracted: nothing is evaluated before raw()
    outer = Dm(0, 3) @ M(1, 2, 4, 5)                                   # [a, b, c, d, e, f] = d[a, d] m[b, c, e, f]
    dense = numpy.einsum("ad,bcef->abcdef", d, m)
    swapped = outer(1, 0, 2, 3, 4, 5)
    _close(XR_tensor.raw(swapped, engine), dense.transpose(1, 0, 2, 3, 4, 5), 1e-15)
    anti = outer - swapped
    w = rng.standard_normal((4, 4))
    fresh = FakeDevice()
    got = XR_tensor.raw(anti(0, "x", "y", 1, 2, 3) @ XR_tensor.init(w)("x", "y"), (DeviceStore(fresh), Contractor(fresh)))
    _close(got, numpy.einsum("axydef,xy->adef", dense - dense.transpose(1, 0, 2, 3, 4, 5), w), 1e-13)
    assert fresh.largest_allocation < dense.size          # the rank-6 tensor itself was never formed
    with pytest.raises(ValueError):
        outer(0, 1, 2)


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_general_randomised_systems_host_logic(seed):
    """seeded random systems -- 2 or 3 fragments, equal or different orbital counts, any subset of the charges -2..2 with 0..3
    states each -- against the element-level port of build_H.py (the same loop run over 190 draws, and 54 hermitian ones at orders 0-2, found no mismatch)"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from oracle import general_oracle as go
    rng = numpy.random.default_rng(seed)
    for _ in range(4):
        F = int(rng.integers(2, 4))
        n_orb = [int(rng.integers(2, 6)) for _ in range(F)] if rng.random() < 0.5 else int(rng.integers(2, 6))
        n_states = {c: int(rng.integers(0, 4)) for c in [0] + [c for c in (1, -1, 2, -2) if rng.random() < 0.6]}
        n_states[0] = max(n_states[0], 1)
        system = synth.make_system(n_frag=F, n_orb=n_orb, n_states=n_states, seed=int(rng.integers(1 << 30)))
        fr, ints, nuc = system["fragments"], system["symm"], system["nuc"]
        eng = build_matrix_elements(fr, ints, nuc, device=FakeDevice())
        for m in range(F):
            _close(eng.H1(m), go.block_monomer(fr, ints, nuc, m))
        for m1, m2 in itertools.combinations(range(F), 2):
            ref = go.block_dimer(fr, ints, nuc, m1, m2)
            _close(eng.H2(m1, m2), ref)
            assert abs(eng.H2_moments(m1, m2)[1] - (ref ** 2).sum()) <= 1e-9 * max((ref ** 2).sum(), 1e-30)
        if F == 3:
            ref = go.block_trimer(fr, ints, (0, 1, 2))
            _close(eng.H3(0, 1, 2), ref)
            s, q = eng.H3_moments(0, 1, 2)
            assert abs(q - (ref ** 2).sum()) <= 1e-9 * max((ref ** 2).sum(), 1e-30)
            assert abs(s - ref.sum()) <= 1e-9 * max(numpy.abs(ref).sum(), 1e-30)


REFERENCE_HERMITIAN = "/root/reference/hermitian-XRCC"


def _reference_get_xr_result(monkeypatch):
    """the reference's own get_xr_result.py, loaded unmodified from where it lies, on top of THIS package's modules under
    their top-level names (hermitian.install_aliases) and the qode stand-in of oracle/qode_shim; returns (module, restore)"""
    import importlib.util
    import sys
    from qodeapplications_b200 import hermitian
    from qodeapplications_b200.hermitian import tensor
    shim = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "qode_shim")
    before = dict(sys.modules)
    monkeypatch.syspath_prepend(os.path.abspath(shim))
    monkeypatch.setattr(tensor, "_default_device", FakeDevice())
    def restore():
        for name in list(sys.modules):
            if name not in before:
                del sys.modules[name]
        sys.modules.update(before)
    try:
        for name in [n for n in sys.modules if n == "qode" or n.startswith("qode.") or n == "excitonic"]:
            del sys.modules[name]
        hermitian.install_aliases()
        spec = importlib.util.spec_from_file_location("reference_get_xr_result", os.path.join(REFERENCE_HERMITIAN, "get_xr_result.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        assert ref.XR_term is hermitian.XR_term and ref.diagrammatic_expansion is hermitian.diagrammatic_expansion
        assert ref.precontract is hermitian.precontract.precontract and ref.D is hermitian.diagram_lists
    except Exception:
        restore()
        raise
    return ref, restore


@pytest.mark.skipif(not os.path.isdir(REFERENCE_HERMITIAN), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)])
def test_reference_driver_runs_unmodified_on_these_modules(order, ops, monkeypatch):
    """The drop-in claim itself: the reference's OWN get_xr_H drives this package's XR_term / diagrammatic_expansion /
    precontract / diagrams / diagram_lists -- its call signatures, its charge-blocked matrices, its Python reorder loop --
    and must reproduce the golden H1/H2 that the all-reference run produced."""
    ref, restore = _reference_get_xr_result(monkeypatch)
    try:
        system = synth.make_system("toy", ops=ops, with_bior=True)
        charges = system["charges"]
        H1, H2 = ref.get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges])
    finally:
        restore()
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"], 1e-9 if order else 1e-10)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_HERMITIAN), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("which", ["bra", "ket"])
def test_reference_driver_det_variants_on_these_modules(which, monkeypatch):
    """the same for get_xr_H(bra_det=True) / (ket_det=True) as StateSpaceOptimizer/state_gradients.py:173,183 calls it"""
    ref, restore = _reference_get_xr_result(monkeypatch)
    try:
        system = synth.make_det_system(which)
        charges = system["charges"]
        H1, H2 = ref.get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"], 0, [charges, charges],
                              bra_det=(which == "bra"), ket_det=(which == "ket"))
    finally:
        restore()
    g = numpy.load(os.path.join(GOLDEN, "hermitian_det_%s.npz" % which))
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"])


def _xr_ccsd_recipe(order, D, dimer, monomer_in_dimer, inverse):
    """H2 the way hermitian-XRCC/mains/xr_ccsd.py:100-185 assembles it at integer S-order `order` (any order, not only the
    0-2 get_xr_H hard-codes): dimer(kind, active) -> matrix of the diagram lists `active` with integrals `kind`."""
    S2 = dimer("S", {0: D.S0[0], 2: sum((D.S2[n] for n in range(1, order + 1)), [])})
    S2H2 = dimer("ST_bior", {1: [], 2: D.ST2[order]}) + dimer("SU_bior", {1: [], 2: D.SU2[order]}) + dimer("SV_bior", {1: [], 2: D.SV2[order]})
    S2H2 = S2H2 + dimer("ST_symm", {1: D.ST1[0], 2: sum((D.ST2[o] for o in range(order)), [])})
    S2H2 = S2H2 + dimer("SU_symm", {1: D.SU1[0], 2: sum((D.SU2[o] for o in range(order)), [])})
    S2H2 = S2H2 + dimer("SV_diff", {1: [], 2: D.SV2[order - 1]})
    S2H2 = S2H2 + dimer("SV_symm", {1: D.SV1[0], 2: sum((D.SV2[o] for o in range(order - 1)), [])})
    S2H2 = S2H2 + dimer("ST_symm", {})              # an empty selection is legal (xr_ccsd.py leaves unused families blank)
    return inverse(S2) @ S2H2 - monomer_in_dimer()


@pytest.mark.parametrize("order", [4])
def test_hermitian_high_order_assembly_in_the_style_of_xr_ccsd(order):
    """mains/xr_ccsd.py drives blocks + XR_term.dimer_matrix itself with the diagram lists of S-orders 3 and 4 (8-operator
    densities): the same recipe through this package and through the oracle"""
    from qodeapplications_b200.hermitian import XR_term, diagrammatic_expansion, diagram_lists as D
    from qodeapplications_b200.hermitian.diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams
    from qodeapplications_b200.hermitian.precontract import precontract
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    from qodeapplications_b200.hermitian.util import struct, timer
    from oracle import hermitian_oracle as ho
    system = synth.make_system("toy4", ops=synth.OPS_ORDER4, with_bior=True)
    symm, bior, dens = system["symm"], system["bior"], system["densities"][:2]
    charges = [(a, b) for a in system["charges"] for b in system["charges"]]
    dev = FakeDevice()
    cache = precontract(dens, symm.S, timer(), store=DeviceStore(dev), contractor=Contractor(dev))
    families = {"S": (symm.S, S_diagrams, ho.integrals(symm.S)),
                "ST_symm": (struct(S=symm.S, T=symm.T), ST_diagrams, ho.integrals(symm.S, T=symm.T)),
                "SU_symm": (struct(S=symm.S, U=symm.U), SU_diagrams, ho.integrals(symm.S, U=symm.U)),
                "SV_symm": (struct(S=symm.S, V=symm.V), SV_diagrams, ho.integrals(symm.S, V=symm.V)),
                "ST_bior": (struct(S=symm.S, T=bior.T), ST_diagrams, ho.integrals(symm.S, T=bior.T)),
                "SU_bior": (struct(S=symm.S, U=bior.U), SU_diagrams, ho.integrals(symm.S, U=bior.U)),
                "SV_bior": (struct(S=symm.S, V=bior.V), SV_diagrams, ho.integrals(symm.S, V=bior.V)),
                "SV_diff": (struct(S=symm.S, V=bior.V_diff), SV_diagrams, ho.integrals(symm.S, V=bior.V_diff))}
    blocks = {k: diagrammatic_expansion.blocks(densities=dens, integrals=ints, diagrams=diagrams, contract_cache=cache, timings=timer(),
                                               precon_timings=timer()) for k, (ints, diagrams, _) in families.items()}
    ours = _xr_ccsd_recipe(order, D,
                           lambda kind, active: XR_term.dimer_matrix(blocks[kind], active, (0, 1), charges, timer()),
                           lambda: sum(XR_term.dimer_matrix(blocks[k], {1: lst[0]}, (0, 1), charges, timer())
                                       for k, lst in (("ST_symm", D.ST1), ("SU_symm", D.SU1), ("SV_symm", D.SV1))),
                           numpy.linalg.inv)
    ref = _xr_ccsd_recipe(order, D,
                          lambda kind, active: ho.dimer_matrix(dens, families[kind][2], active, charges),
                          lambda: sum(ho.dimer_matrix(dens, families[k][2], {1: lst[0]}, charges)
                                      for k, lst in (("ST_symm", D.ST1), ("SU_symm", D.SU1), ("SV_symm", D.SV1))),
                          numpy.linalg.inv)
    _close(ours, ref, 1e-9)


@pytest.mark.parametrize("order", ["proper", 0, 1, 2, 3, 4])
def test_hermitian_xr_ccsd_build_H_host_logic(order):
    """hermitian/xr_ccsd.build_H (the Hamiltonian build of mains/xr_ccsd.py:69-213, any S-order): "proper", 1 and 2 against
    the reference's golden get_xr_H output (xr_order 0, 1, 2 -- the same matrices), 0, 3 and 4 against the oracle running
    the script's recipe"""
    from qodeapplications_b200.hermitian.xr_ccsd import build_H
    from qodeapplications_b200.hermitian import diagram_lists as D
    from oracle import hermitian_oracle as ho
    golden = {"proper": 0, 1: 1, 2: 2}.get(order)
    name, ops = ("toy", {0: synth.OPS_ORDER0, 1: synth.OPS_ORDER1, 2: synth.OPS_ORDER2}[golden]) if golden is not None else ("toy4", synth.OPS_ORDER4)
    system = synth.make_system(name, ops=ops, with_bior=True)
    symm, bior, dens, ch = system["symm"], system["bior"], system["densities"][:2], system["charges"]
    H1, H2 = build_H((symm, bior, system["nuc"]), dens, order, ch, device=FakeDevice())
    if golden is not None:
        g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % golden))
        _close(H1[0], g["H1_0"])
        _close(H1[1], g["H1_1"])
        _close(H2, g["H2"], 1e-9 if golden else 1e-10)
        return
    charges = [(a, b) for a in ch for b in ch]
    kinds = {"S": ho.integrals(symm.S), "ST_symm": ho.integrals(symm.S, T=symm.T), "SU_symm": ho.integrals(symm.S, U=symm.U),
             "SV_symm": ho.integrals(symm.S, V=symm.V), "ST_bior": ho.integrals(symm.S, T=bior.T), "SU_bior": ho.integrals(symm.S, U=bior.U),
             "SV_bior": ho.integrals(symm.S, V=bior.V), "SV_diff": ho.integrals(symm.S, V=bior.V_diff)}
    dimer = lambda kind, active: ho.dimer_matrix(dens, kinds[kind], active, charges)
    monomers = lambda: sum(dimer(k, {1: lst[0]}) for k, lst in (("ST_symm", D.ST1), ("SU_symm", D.SU1), ("SV_symm", D.SV1)))
    if order == 0:      # xr_ccsd.py:120-134 at order 0: S = 1, zeroth-order diagrams (monomer ones included) with biorthogonal integrals
        blocked = (dimer("ST_bior", {1: D.ST1[0], 2: D.ST2[0]}) + dimer("SU_bior", {1: D.SU1[0], 2: D.SU2[0]})
                   + dimer("SV_bior", {1: D.SV1[0], 2: D.SV2[0]}) - monomers())
    else:
        blocked = _xr_ccsd_recipe(order, D, dimer, monomers, numpy.linalg.inv)
    _close(H2, ho.reorder(blocked, dens, [ch, ch]), 1e-9)
    for m in (0, 1):
        _close(H1[m], sum(ho.monomer_matrix(dens, kinds[k], lst[0], m, ch) for k, lst in (("ST_symm", D.ST1), ("SU_symm", D.SU1), ("SV_symm", D.SV1))))


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)])
def test_hermitian_plan_replays_get_xr_H_on_new_densities(order, ops):
    """hermitian/plan.py: the launch sequence recorded on one set of densities, replayed on another, equals a fresh
    get_xr_H on the new densities (and the golden reference result on the original ones)"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from qodeapplications_b200.hermitian.plan import plan
    system = synth.make_system("toy", ops=ops, with_bior=True)
    ch = system["charges"]
    ints = (system["symm"], system["bior"], system["nuc"])
    dev = FakeDevice()
    build = plan(ints, system["densities"][:2], order, [ch, ch], device=dev)
    assert build.launches > 10
    H1, H2 = build()
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    _close(H1[0], g["H1_0"])
    _close(H2, g["H2"], 1e-9 if order else 1e-10)
    # new densities of the same shapes (another seed): the replay must follow them
    other = synth.make_system("toy", ops=ops, with_bior=True, seed=77)["densities"][:2]
    R1, R2 = build(other)
    E1, E2 = get_xr_H(ints, other, order, [ch, ch], device=FakeDevice())
    _close(R2, E2, 1e-12)
    _close(R1[1], E1[1], 1e-12)
    assert numpy.abs(R2 - H2).max() > 1e-3 * numpy.abs(H2).max()
    again = build(other)[1]
    assert numpy.array_equal(again, R2)


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1)])
def test_hermitian_schedule_dependencies_allow_any_valid_order(order, ops):
    """hermitian/schedule.py: the dependencies worked out from the recorded arguments are sufficient -- the recorded
    get_xr_H re-issued in a very different order that respects them (always the LAST ready call first) gives bit-identical
    matrices -- and they are not trivial (most calls do not depend on their predecessor)"""
    from qodeapplications_b200.hermitian.plan import plan
    from qodeapplications_b200.hermitian import schedule
    system = synth.make_system("toy", ops=ops, with_bior=True)
    ch = system["charges"]
    dev = FakeDevice()
    build = plan((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [ch, ch], device=dev, verify=False)
    H1, H2 = build()
    deps = schedule.dependencies(build.trace)
    n = len(build.trace)
    assert sum(1 for i, d in enumerate(deps) if i and (i - 1) not in d) > n // 3
    users = [[] for _ in range(n)]
    missing = [len(d) for d in deps]
    for i, d in enumerate(deps):
        for j in d:
            users[j].append(i)
    ready = [i for i in range(n) if missing[i] == 0]
    issued = []
    while ready:
        i = ready.pop()                      # the most recently readied / highest index first: far from program order
        issued.append(i)
        call, args, kwargs = build.trace[i]
        call(dev.ctx, *args, **kwargs)
        for u in users[i]:
            missing[u] -= 1
            if missing[u] == 0:
                ready.append(u)
                ready.sort()
    assert len(issued) == n and issued != sorted(issued)
    assert numpy.array_equal(build.H2.host(), H2) and numpy.array_equal(build.H1[0].host(), H1[0])
    stream_of, cross = schedule.assign_streams(deps, 6)
    assert len(set(stream_of)) > 1 and all(stream_of[j] != stream_of[i] for i, c in enumerate(cross) for j in c)


def test_general_recorded_step_replays_and_follows_new_densities():
    """general/distributed.recorded_step: a recorded build step replayed on new densities equals a fresh engine's blocks, and
    the dependencies computed from its recorded arguments (raw-pointer offset tables and trimer streams included) allow a
    very different issue order bit for bit"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import sharded_build
    from qodeapplications_b200 import schedule
    system = synth.make_system("toy3")
    dev = FakeDevice()
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    dimers = list(itertools.combinations(range(3), 2))
    build = sharded_build(eng, dimers, [(0, 1, 2)])
    rec = build.recorded()
    rec.run()
    g = numpy.load(os.path.join(GOLDEN, "general_toy3.npz"))
    _close(build.full(0, 1).numpy(), g["H2_01"])
    _close(build.H1[1].numpy(), g["H1_1"])
    other = synth.make_system("toy3", seed=99)
    rec.update(other["fragments"])
    rec.run()
    fresh = build_matrix_elements(other["fragments"], system["symm"], system["nuc"], device=FakeDevice())
    assert numpy.array_equal(build.full(0, 2).numpy(), fresh.H2(0, 2))
    assert numpy.array_equal(build.H1[2].numpy(), fresh.H1(2))
    assert numpy.array_equal(build.H3_moments[(0, 1, 2)].numpy(), fresh.H3_moments(0, 1, 2, per_class=True))
    # any order that respects the computed dependencies gives the same bits
    known = [t for t in rec._alive if t.dtype == torch.int64] + list(eng._idx_dev.values())
    deps = schedule.dependencies(rec.trace, known)
    n = len(deps)
    assert sum(1 for i, d in enumerate(deps) if i and (i - 1) not in d) > n // 2
    users, missing = [[] for _ in range(n)], [len(d) for d in deps]
    for i, d in enumerate(deps):
        for j in d:
            users[j].append(i)
    ready, issued = [i for i in range(n) if missing[i] == 0], 0
    before = (build.full(0, 2).numpy().copy(), build.H3_moments[(0, 1, 2)].numpy().copy())
    while ready:
        i = ready.pop()
        call, args, kwargs = rec.trace[i]
        call(dev.ctx, *args, **kwargs)
        issued += 1
        for u in users[i]:
            missing[u] -= 1
            if missing[u] == 0:
                ready.append(u)
                ready.sort()
    assert issued == n
    assert numpy.array_equal(build.full(0, 2).numpy(), before[0]) and numpy.array_equal(build.H3_moments[(0, 1, 2)].numpy(), before[1])


def test_schedule_treats_unknown_calls_and_unsized_operands_as_barriers():
    """schedule.py is conservative by construction: a call it has no footprint for, or an offset table it cannot read,
    depends on everything before it and everything after depends on it"""
    from qodeapplications_b200 import schedule
    dev = FakeDevice()
    a, b, c, d = (dev.zeros((4, 4)) for _ in range(4))

    def named(name):
        def call(ctx, *args, **kwargs):
            return None
        call.__name__ = name
        return call
    trace = [(named("copy2d_scaled"), (b, 4, a, 4, 4, 4, 1.0), {}),                  # b <- a
             (named("copy2d_scaled"), (d, 4, c, 4, 4, 4, 1.0), {}),                  # d <- c: independent of call 0
             (named("something_new"), (a,), {}),                                     # unknown: a barrier
             (named("copy2d_scaled"), (c, 4, d, 4, 4, 4, 1.0), {}),                  # after the barrier
             (named("gemm_scatter"), (4, 4, 4, 1.0, a, 4, b, 4, c, 12345, 0, None, False), {}),   # raw-pointer table nobody knows
             (named("copy2d_scaled"), (b, 4, a, 4, 4, 4, 1.0), {})]
    deps = schedule.dependencies(trace)
    assert deps[0] == [] and deps[1] == []
    assert deps[2] == [0, 1]
    assert deps[3] == [2]
    assert deps[4] == [3]                  # (the barrier 2 is implied through 3)
    assert deps[5] == [4]
    stream_of, cross = schedule.assign_streams(deps, 3)
    assert stream_of[0] != stream_of[1]
    assert all(stream_of[j] != stream_of[i] for i, cs in enumerate(cross) for j in cs)


def test_hermitian_plan_with_device_resident_densities_updated_in_place():
    """densities already on the device (DeviceTensor blocks, what a density builder on the GPU hands over) ARE the input
    slots of the plan: the caller rewrites them in place and calls build() with no arguments"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from qodeapplications_b200.hermitian.plan import plan
    from qodeapplications_b200.hermitian.tensor import DeviceTensor
    system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    other = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True, seed=5)["densities"][:2]
    ch = system["charges"]
    ints = (system["symm"], system["bior"], system["nuc"])
    dev = FakeDevice()
    resident = []
    for rho in system["densities"][:2]:
        resident.append({key: ({sector: DeviceTensor(dev.upload(block), dev) for sector, block in value.items()}
                               if isinstance(value, dict) and key not in ("n_elec", "n_states", "n_states_bra") else value)
                         for key, value in rho.items()})
    build = plan(ints, resident, 1, [ch, ch], device=dev)
    H1, H2 = build()
    E1, E2 = get_xr_H(ints, system["densities"][:2], 1, [ch, ch], device=FakeDevice())
    _close(H2, E2, 1e-12)
    for m, rho in enumerate(other):                       # new values written into the caller's own device buffers
        for key, value in rho.items():
            if isinstance(value, dict) and key not in ("n_elec", "n_states", "n_states_bra"):
                for sector, block in value.items():
                    resident[m][key][sector].buf.copy_(torch.from_numpy(numpy.ascontiguousarray(block)))
    R1, R2 = build()
    F1, F2 = get_xr_H(ints, other, 1, [ch, ch], device=FakeDevice())
    _close(R2, F2, 1e-12)
    _close(R1[0], F1[0], 1e-12)
    with pytest.raises(ValueError):                       # a block of another shape is not this plan's input
        bad = [dict(rho) for rho in other]
        bad[0]["ca"] = {sector: numpy.zeros((1,) + block.shape) for sector, block in other[0]["ca"].items()}
        build(bad)
