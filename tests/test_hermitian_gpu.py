"""GPU parity tests for the hermitian-XRCC path: diagram blocks, XR_term matrices and get_xr_H computed
by libxr_b200.so against the reference's golden vectors (toy) and the NumPy oracle (larger shapes).
Tolerance (BASELINE.json north_star, SURVEY section 7 hard part 7), at every xr_order: PER ELEMENT
|diff| <= 1e-10 * max(|H_ref|, 1e-3 * rms of the block/matrix)."""
import os
import numpy
import pytest

from qodeapplications_b200 import synth
from oracle import hermitian_oracle as ho

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, tol=None):
    """H blocks (tol=None): the per-element bar of north_star.  Kernel unit tests pass their own, much tighter, tolerance
    relative to the largest element."""
    a, b = numpy.asarray(a), numpy.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if tol is None:
        floor = 1e-3 * max(float(numpy.sqrt(numpy.mean(b * b))), 1e-300) if b.size else 0.0
        excess = numpy.abs(a - b) - 1e-10 * numpy.maximum(numpy.abs(b), floor)
        assert b.size == 0 or excess.max() <= 0, (float(numpy.abs(a - b).max()), floor)
    else:
        assert b.size == 0 or numpy.abs(a - b).max() <= tol * max(numpy.abs(b).max(), 1e-300), (numpy.abs(a - b).max(), tol)


@pytest.fixture(scope="module")
def dev():
    from qodeapplications_b200.device import Device
    return Device(0)


def _blocks(system, dev, family, which="symm"):
    from qodeapplications_b200.hermitian import diagrammatic_expansion
    from qodeapplications_b200.hermitian.diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams
    from qodeapplications_b200.hermitian.precontract import precontract
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    from qodeapplications_b200.hermitian.util import struct, timer
    ints_set, dens = system[which], system["densities"][:2]
    S = system["symm"].S
    store, contractor = DeviceStore(dev), Contractor(dev)
    cache = precontract(dens, S, timer(), store=store, contractor=contractor)
    ints, diagrams = {"S": (S, S_diagrams), "ST": (struct(S=S, T=ints_set.T), ST_diagrams),
                      "SU": (struct(S=S, U=ints_set.U), SU_diagrams), "SV": (struct(S=S, V=ints_set.V), SV_diagrams)}[family]
    return diagrammatic_expansion.blocks(densities=dens, integrals=ints, diagrams=diagrams, contract_cache=cache,
                                         timings=timer(), precon_timings=timer())


@pytest.mark.parametrize("fixture,n_blocks,ops,name", [("hermitian_toy_blocks.npz", 221, synth.OPS_ORDER1, "toy"),
                                                       ("hermitian_toy_blocks2.npz", 165, synth.OPS_ORDER2, "toy"),
                                                       ("hermitian_toy4_blocks34.npz", 328, synth.OPS_ORDER4, "toy4")])
def test_every_diagram_block_matches_reference_golden(dev, fixture, n_blocks, ops, name):
    g = numpy.load(os.path.join(GOLDEN, fixture))
    toy = synth.make_system(name, ops=ops, with_bior=True)       # the fixture was generated from exactly this draw
    def family_of(label):
        rest = label.replace("s01", "").replace("s10", "")
        return "S" if rest == "" else "S" + rest[0].upper()
    cache, count = {}, 0
    for key in g.files:
        if key == "input_sha256":
            continue
        parts = key.split("|")
        label = parts[0]
        fam = family_of(label)
        if fam not in cache:
            cache[fam] = _blocks(toy, dev, fam)
        if len(parts) == 3:
            ci, cj = (int(x) for x in parts[2].split(","))
            _close(cache[fam][(int(parts[1]),)][((ci, cj),)][label], g[key])
        else:
            ci0, ci1, cj0, cj1 = (int(x) for x in parts[1].split(","))
            _close(cache[fam][(0, 1)][((ci0, cj0), (ci1, cj1))][label], g[key])
        count += 1
    assert count == n_blocks


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1), (2, synth.OPS_ORDER2)])
def test_get_xr_H_matches_reference_golden(dev, order, ops):
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    system = synth.make_system("toy", ops=ops, with_bior=True)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges], device=dev)
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"])


@pytest.mark.parametrize("order", [0, 1])
def test_get_xr_S_matches_reference_golden(dev, order):
    """get_xr_S (get_xr_result.py:357-422, orbital solver): golden = the reference's own charge-blocked matrix"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_S
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_S.npz"))
    system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    charges = system["charges"]
    S2 = get_xr_S((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges], device=dev)
    _close(S2, ho.reorder(g["S2_blocked_order%d" % order], system["densities"], [charges, charges]))


def test_cfg1_order0_against_oracle(dev):
    """Be2 / 6-31G shapes (n = 18, N = 11/4/8): get_xr_H(order 0) against the NumPy oracle"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    system = synth.make_system("cfg1", ops=synth.OPS_ORDER0, with_bior=True)
    charges = system["charges"]
    args = (system["densities"][:2], 0, [charges, charges])
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), *args, device=dev)
    R1, R2 = ho.get_xr_H(system["symm"], system["bior"], *args)
    _close(H1[0], R1[0])
    _close(H1[1], R1[1])
    _close(H2, R2)


def test_mid_order1_against_oracle(dev):
    """n = 8, N = 5/3/4: get_xr_H(order 1) (rank-5 densities, S2 inverse) against the NumPy oracle"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    system = synth.make_system("mid", ops=synth.OPS_ORDER1, with_bior=True)
    charges = system["charges"]
    args = (system["densities"][:2], 1, [charges, charges])
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), *args, device=dev)
    R1, R2 = ho.get_xr_H(system["symm"], system["bior"], *args)
    _close(H1[0], R1[0])
    _close(H2, R2)


def test_mid_order2_against_oracle(dev):
    """n = 8, N = 5/3/4: get_xr_H(order 2) (rank-6 densities, three-factor diagrams) against the NumPy oracle"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    system = synth.make_system("mid", ops=synth.OPS_ORDER2, with_bior=True)
    charges = system["charges"]
    args = (system["densities"][:2], 2, [charges, charges])
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), *args, device=dev)
    R1, R2 = ho.get_xr_H(system["symm"], system["bior"], *args)
    _close(H2, R2)


def test_xr_tensor_expression_syntax(dev):
    """the tensornet-style seam: raw( A(0,1,"p","q") @ B(2,"q") @ C("p",3) ) evaluated on the GPU"""
    from qodeapplications_b200.hermitian import XR_tensor
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceStore
    rng = numpy.random.default_rng(3)
    a, b, c = rng.standard_normal((3, 4, 5, 6)), rng.standard_normal((7, 6)), rng.standard_normal((5, 2))
    A, B, C = XR_tensor.init(a), XR_tensor.init(b), XR_tensor.init(c)
    engine = (DeviceStore(dev), Contractor(dev))
    got = XR_tensor.raw(-2.5 * (A(0, 1, "p", "q") @ B(2, "q") @ C("p", 3)), engine)
    _close(got, -2.5 * numpy.einsum("ijpq,kq,pl->ijkl", a, b, c), 1e-13)
    got = XR_tensor.raw(A(1, 0, "p", 2) @ C("p", 3), engine)           # free labels sorted ascending in the result
    _close(got, numpy.einsum("jipq,pl->ijql", a, c), 1e-13)
    _close(XR_tensor.raw(3.0 * A(3, 1, 0, 2), engine), 3.0 * a.transpose(2, 1, 3, 0), 1e-15)


def test_dimer_matrix_blocked_and_final_orderings(dev):
    from qodeapplications_b200.hermitian import XR_term
    from qodeapplications_b200.hermitian.util import timer
    toy = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    blk = _blocks(toy, dev, "SV")
    charges = [(a, b) for a in toy["charges"] for b in toy["charges"]]
    active = {1: ["v0000"], 2: ["v0101", "v0001", "v0100", "v0011", "s01v0101"]}
    symm = toy["symm"]
    ref = ho.dimer_matrix(toy["densities"], ho.integrals(symm.S, V=symm.V), active, charges)
    _close(XR_term.dimer_matrix(blk, active, (0, 1), charges, timer()), ref)
    final = XR_term.dimer_matrix(blk, active, (0, 1), charges, timer(), ordering="final")
    _close(final, ho.reorder(ref, toy["densities"], [toy["charges"], toy["charges"]]))


def test_permute_copy(dev):
    rng = numpy.random.default_rng(0)
    a = rng.standard_normal((3, 4, 5, 6))
    dA = dev.upload(a)
    out = dev.empty((6, 4, 3, 5))
    strides = [5 * 6 * 4, 5 * 6, 6, 1]            # element strides of a's axes 0..3
    order = [3, 1, 0, 2]
    dev.ctx.permute_copy(out, dA, [a.shape[k] for k in order], [strides[k] for k in order], -2.0)
    assert numpy.array_equal(dev.download(out), -2.0 * a.transpose(order))


@pytest.mark.parametrize("which", ["bra", "ket"])
def test_det_variants_match_reference_golden(dev, which):
    """get_xr_H(bra_det=True) / (ket_det=True) at xr_order 0 against the reference's own output"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    g = numpy.load(os.path.join(GOLDEN, "hermitian_det_%s.npz" % which))
    system = synth.make_det_system(which)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"], 0, [charges, charges],
                      bra_det=(which == "bra"), ket_det=(which == "ket"), device=dev)
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"])


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1)])
@pytest.mark.parametrize("n_states", [{0: 2, +1: 1, -1: 0}, {0: 1, +1: 1, -1: 1}, {0: 2, +1: 0, -1: 2}],
                         ids=lambda d: "-".join("%+d:%d" % kv for kv in d.items()))
def test_ragged_and_empty_charge_sectors(dev, order, ops, n_states):
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    system = synth.make_system(n_frag=2, n_orb=4, n_states=n_states, seed=78, ops=ops, with_bior=True)
    ch = system["charges"]
    args = (system["densities"][:2], order, [ch, ch])
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), *args, device=dev)
    R1, R2 = ho.get_xr_H(system["symm"], system["bior"], *args)
    _close(H1[0], R1[0])
    _close(H1[1], R1[1])
    _close(H2, R2)


@pytest.mark.parametrize("order,ops", [(0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1)])
def test_row_sharded_get_xr_H_single_rank(dev, order, ops):
    """shard=(0, 1): the row-sharded code path (slab buffers, slab rows of S2inv) without a collective"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order%d.npz" % order))
    system = synth.make_system("toy", ops=ops, with_bior=True)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], order, [charges, charges], device=dev,
                      shard=(0, 1))
    _close(H1[0], g["H1_0"])
    _close(H1[1], g["H1_1"])
    _close(H2, g["H2"])


@pytest.mark.parametrize("rank", [0, 1, 2])
def test_row_sharded_get_xr_H_own_rows(dev, rank, monkeypatch):
    """each of 3 ranks' slabs of the order-0 build (bra offsets in the spectator deltas, a short last slab) computed on
    this GPU with the all-gather stubbed out: the rank's own rows must be the reference's"""
    import torch.distributed as dist
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from qodeapplications_b200.general.distributed import slab_bounds
    monkeypatch.setattr(dist, "all_gather_into_tensor", lambda out, inp, group=None: None)
    g = numpy.load(os.path.join(GOLDEN, "hermitian_toy_order0.npz"))
    system = synth.make_system("toy", ops=synth.OPS_ORDER0, with_bior=True)
    charges = system["charges"]
    H1, H2 = get_xr_H((system["symm"], system["bior"], system["nuc"]), system["densities"][:2], 0, [charges, charges], device=dev,
                      shard=(rank, 3))
    dim = g["H1_0"].shape[0]
    lo, hi, per = slab_bounds(dim, rank, 3)
    _close(H1[0][lo:hi], g["H1_0"][lo:hi])
    _close(H2[lo * dim:hi * dim], g["H2"][lo * dim:hi * dim])


def test_factored_densities(dev):
    """densities of 5..8 operators given as lazy sums of outer products (frag-states/decomps.py form): a sample of the
    S-order 1-4 diagram blocks that read them and get_xr_H at order 2, against the oracle on the written-out tensors"""
    import itertools
    from test_host_logic_cpu import factored_system, _hermitian_blocks
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    lazy, dense = factored_system()
    charges, symm = lazy["charges"], lazy["symm"]
    for fam, ints, labels in (("ST", ho.integrals(symm.S, T=symm.T), ["s01t00", "s01s01s10t10", "s01s01s01s10t00"]),
                              ("SV", ho.integrals(symm.S, V=symm.V), ["s01v0000", "s01s10v0101", "s01s01s01s10v1111", "s01s01s10s10v0011"])):
        blk = _hermitian_blocks(lazy, dev, fam)
        for label in labels:
            checked = 0
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
            assert checked
    lazy2, dense2 = factored_system("toy", synth.OPS_ORDER2, min_rank=5, seed=6)
    args = (2, [lazy2["charges"], lazy2["charges"]])
    H1, H2 = get_xr_H((lazy2["symm"], lazy2["bior"], lazy2["nuc"]), lazy2["densities"][:2], *args, device=dev)
    R1, R2 = ho.get_xr_H(dense2["symm"], dense2["bior"], dense2["densities"][:2], *args)
    _close(H1[0], R1[0])
    _close(H2, R2)


@pytest.mark.parametrize("name,order,ops", [("toy", 0, synth.OPS_ORDER0), ("toy", 2, synth.OPS_ORDER2), ("mid", 1, synth.OPS_ORDER1)])
def test_plan_replays_get_xr_H_as_one_cuda_graph(dev, name, order, ops):
    """hermitian/plan.py on the GPU: the recorded launch sequence is captured into a CUDA graph; replayed on new densities
    it equals a fresh get_xr_H on them (bit for bit: the same kernels on the same buffers) and the oracle"""
    from qodeapplications_b200.hermitian.get_xr_result import get_xr_H
    from qodeapplications_b200.hermitian.plan import plan
    system = synth.make_system(name, ops=ops, with_bior=True)
    ch = system["charges"]
    ints = (system["symm"], system["bior"], system["nuc"])
    build = plan(ints, system["densities"][:2], order, [ch, ch], device=dev)
    assert build.graph is not None, getattr(build, "graph_error", None)
    assert build.n_streams > 1, getattr(build, "streams_error", None)        # independent diagram GEMMs on several streams
    H1, H2 = build()
    assert build.n_streams > 1, "the multi-stream graph did not reproduce the eager build"
    E1, E2 = get_xr_H(ints, system["densities"][:2], order, [ch, ch], device=dev)
    assert numpy.array_equal(H2, E2) and numpy.array_equal(H1[0], E1[0])
    other = synth.make_system(name, ops=ops, with_bior=True, seed=78)["densities"][:2]
    R1, R2 = build(other)
    F1, F2 = get_xr_H(ints, other, order, [ch, ch], device=dev)
    assert numpy.array_equal(R2, F2) and numpy.array_equal(R1[1], F1[1])
    O1, O2 = ho.get_xr_H(system["symm"], system["bior"], other, order, [ch, ch])
    _close(R2, O2)
    _close(R1[0], O1[0])
    assert numpy.array_equal(build(other)[1], R2)
    chain = plan(ints, other, order, [ch, ch], device=dev, streams=1)
    assert chain.n_streams == 1 and numpy.array_equal(chain()[1], R2)


def test_device_inverse_converges_and_falls_back(dev):
    """get_xr_result.precise_inverse: Newton-Schulz from the identity + double-double polish on the device; a matrix far
    from the identity leaves a large residual and checked_inverse then takes the LAPACK starting guess"""
    from qodeapplications_b200.hermitian.get_xr_result import precise_inverse, checked_inverse, INVERSE_RESIDUAL_TOL
    rng = numpy.random.default_rng(9)
    n = 300
    noise = rng.standard_normal((n, n))
    S = numpy.eye(n) + 0.3 * (noise + noise.T) / numpy.linalg.norm(noise + noise.T, 2)
    X, res2 = precise_inverse(S, dev)
    assert float(res2.host()) <= INVERSE_RESIDUAL_TOL
    assert numpy.abs(X.host() @ S - numpy.eye(n)).max() <= 1e-14
    far = 5.0 * numpy.eye(n) + noise / numpy.sqrt(n)            # |I - S| > 1: the iteration from the identity diverges
    _, res2 = precise_inverse(far, dev)
    assert not float(res2.host()) <= INVERSE_RESIDUAL_TOL
    Y = checked_inverse(far, dev)
    assert numpy.abs(Y.host() @ far - numpy.eye(n)).max() <= 1e-13


def test_contractor_streams_large_interleaved_operands_without_reordering(dev):
    """a large operand whose free index sits between contracted ones (rho[i,j,a,b,c,d,e] contracted over a,b,d,e) goes
    through xr_gemm_stream where it lies: no xr_permute_copy of it is launched, and the result equals the einsum"""
    from qodeapplications_b200.hermitian.tensor import Contractor, DeviceTensor
    rng = numpy.random.default_rng(12)
    n = 12
    rho, V = rng.standard_normal((3, 4, n, n, n, n, n)), rng.standard_normal((n, n, n, n))
    W = rng.standard_normal((5, n, n, n))
    R, dV, dW = (DeviceTensor(dev.upload(x), dev) for x in (rho, V, W))
    C = Contractor(dev)
    C.STREAM_MIN_ELEMENTS = 1 << 18
    dev.begin_trace()
    out = C.contract(R, ["i", "j", "a", "b", "c", "d", "e"], dV, ["a", "b", "d", "e"], ["c", "i", "j"])
    rho4 = DeviceTensor(R.buf[0], dev)            # [j, a, b, c, d, e] viewed as [i=j, j=a, p, x, s, r]
    out2 = C.contract(dW, ["f", "p", "s", "r"], rho4, ["i", "j", "p", "x", "s", "r"], ["x", "f", "i", "j"])
    trace, _ = dev.end_trace()
    names = [call.__name__ for call, a, k in trace]
    assert names.count("gemm_stream") == 2
    big = [a for call, a, k in trace if call.__name__ == "permute_copy" and numpy.prod(a[2]) >= rho[0].size]
    assert not big, "the large operand was re-ordered"
    _close(out.host(), numpy.einsum("ijabcde,abde->cij", rho, V), 1e-12)
    _close(out2.host(), numpy.einsum("fpsr,ijpxsr->xfij", W, rho[0]), 1e-12)
