"""GPU parity tests for the general-XRCC path: every value computed by libxr_b200.so is compared
with the CPU oracle (and with the golden vectors the reference itself produced).
Tolerance (BASELINE.json north_star, SURVEY section 7 hard part 7): PER ELEMENT |diff| <= 1e-10 * max(|H_ref|, 1e-3 * rms of the
block); elements that are structurally zero in the reference must be exactly zero."""
import itertools
import os
import numpy
import pytest
import torch

from qodeapplications_b200 import synth
from oracle import general_oracle as go

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-10


def _close(a, b, tol=None):
    """H blocks (tol=None): the per-element bar of north_star.  Kernel unit tests pass their own, much tighter, tolerance
    relative to the largest element."""
    a, b = numpy.asarray(a), numpy.asarray(b)
    assert a.shape == b.shape
    if tol is None:
        floor = 1e-3 * max(float(numpy.sqrt(numpy.mean(b * b))), 1e-300) if b.size else 0.0
        excess = numpy.abs(a - b) - 1e-10 * numpy.maximum(numpy.abs(b), floor)
        assert b.size == 0 or excess.max() <= 0, (float(numpy.abs(a - b).max()), floor)
    else:
        assert b.size == 0 or numpy.abs(a - b).max() <= tol * max(numpy.abs(b).max(), 1e-300), (numpy.abs(a - b).max(), tol)
    assert numpy.all(a[b == 0] == 0), "structural zeros must stay exactly zero"


@pytest.fixture(scope="module")
def dev():
    from qodeapplications_b200.device import Device
    return Device(0)


def _engine(system, dev):
    from qodeapplications_b200.general.build_H import build_matrix_elements
    return build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)


# ------------------------------------------------------------------------------- kernels

@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (64, 64, 16), (130, 70, 37), (257, 300, 324), (1000, 33, 18),
                                   (129, 1, 325), (300, 520, 36),
                                   (130, 3, 5000), (64, 1, 104976), (300, 18, 5832), (1584, 1, 20000)])     # split-K path
@pytest.mark.parametrize("aligned", [True, False])
def test_gemm_scatter_plain(dev, M, N, K, aligned):
    rng = numpy.random.default_rng(M * 1000 + N * 10 + K)
    lda = K + (K % 2 if aligned else 1 - K % 2)      # even (vector path) or odd (8-byte path)
    ldb = lda + (2 if aligned else 0)
    A, B = rng.standard_normal((M, lda)), rng.standard_normal((N, ldb))
    dA, dB = dev.upload(A), dev.upload(B)
    ldc = N + 3
    C0 = rng.standard_normal((M, ldc))
    dC = dev.upload(C0)
    ref = 0.75 * A[:, :K] @ B[:, :K].T
    dev.ctx.gemm_scatter(M, N, K, 0.75, dA, lda, dB, ldb, dC, None, ldc, None, False)
    out = dev.download(dC)
    _close(out[:, :N], ref, 1e-13 * max(1, K))
    assert numpy.array_equal(out[:, N:], C0[:, N:])       # padding columns untouched
    dev.ctx.gemm_scatter(M, N, K, -0.5, dA, lda, dB, ldb, dC, None, ldc, None, True)
    _close(dev.download(dC)[:, :N], ref - 0.5 * A[:, :K] @ B[:, :K].T, 1e-13 * max(1, K))


@pytest.mark.parametrize("E1,E2,EK1,EK2,N", [(5, 18, 36, 36, 1), (3, 1, 1, 5000, 1), (7, 18, 324, 324, 1), (2, 70, 16, 100, 8),
                                             (11, 18, 18, 324, 18), (130, 1, 1, 2600, 32), (4, 6, 36, 36, 3), (1, 100, 40, 18, 5)])
def test_gemm_stream(dev, E1, E2, EK1, EK2, N):
    """xr_gemm_stream: A addressed through a 4-D tensor map as [E1][EK1][E2][EK2] (free index between contracted ones),
    tall-skinny output, K split over the GPU; against numpy on the same strided view"""
    rng = numpy.random.default_rng(E1 + E2 + EK1 + EK2 + N)
    T = rng.standard_normal((E1, EK1, E2, EK2))
    K = EK1 * EK2
    ldb = K + 2
    B = rng.standard_normal((N, ldb))
    ref = -0.5 * numpy.einsum("aubv,nuv->abn", T, B[:, :K].reshape(N, EK1, EK2)).reshape(E1 * E2, N)
    dT, dB = dev.upload(T), dev.upload(B)
    M = E1 * E2
    offM = numpy.arange(M, dtype=numpy.int64) * 3                 # output transposed-ish: C[n, m] with padding
    offN = numpy.arange(N, dtype=numpy.int64) * (3 * M + 1)
    C0 = rng.standard_normal((N * (3 * M + 1) + 5,))
    dC = dev.upload(C0)
    ok = dev.ctx.gemm_stream(E1, EK1 * E2 * EK2, E2, EK2, EK1, E2 * EK2, EK2, N, -0.5, dT, dB, ldb, dC,
                             dev.upload(offM, numpy.int64), 0, dev.upload(offN, numpy.int64), True)
    assert ok
    want = C0.copy()
    want[(offM[:, None] + offN[None, :]).reshape(-1)] += ref.reshape(-1)
    got = dev.download(dC)
    assert numpy.abs(got - want).max() <= 1e-13 * K ** 0.5 * max(1.0, numpy.abs(ref).max())
    # plain output, overwrite
    dP = dev.empty((M, N))
    assert dev.ctx.gemm_stream(E1, EK1 * E2 * EK2, E2, EK2, EK1, E2 * EK2, EK2, N, -0.5, dT, dB, ldb, dP, None, N, None, False)
    assert numpy.abs(dev.download(dP) - ref).max() <= 1e-13 * K ** 0.5 * max(1.0, numpy.abs(ref).max())
    # bit-reproducible
    dQ = dev.empty((M, N))
    dev.ctx.gemm_stream(E1, EK1 * E2 * EK2, E2, EK2, EK1, E2 * EK2, EK2, N, -0.5, dT, dB, ldb, dQ, None, N, None, False)
    assert numpy.array_equal(dev.download(dQ), dev.download(dP))


def test_gemm_stream_declines_what_it_cannot_address(dev):
    rng = numpy.random.default_rng(1)
    T, B = dev.upload(rng.standard_normal((4, 3, 5, 7))), dev.upload(rng.standard_normal((2, 22)))
    C = dev.zeros((20, 2))
    assert dev.ctx.gemm_stream(4, 105, 5, 7, 3, 35, 7, 2, 1.0, T, B, 22, C, None, 2, None, False) is False       # odd strides
    assert dev.ctx.gemm_stream(4, 106, 5, 8, 3, 40, 8, 40, 1.0, T, B, 24, C, None, 40, None, False) is False     # N > 32
    assert numpy.all(dev.download(C) == 0)


def test_gemm_scatter_offset_tables(dev):
    """C[i0,i1,j0,j1] <- A[(i0,j0),k] B[(i1,j1),k]: the [ikjl] shuffle every dimer diagram needs."""
    rng = numpy.random.default_rng(7)
    n0i, n0j, n1i, n1j, K = 5, 4, 3, 6, 50
    A, B = rng.standard_normal((n0i * n0j, K)), rng.standard_normal((n1i * n1j, K))
    offM = (numpy.arange(n0i)[:, None] * (n1i * n0j * n1j) + numpy.arange(n0j)[None, :] * n1j).reshape(-1)
    offN = (numpy.arange(n1i)[:, None] * (n0j * n1j) + numpy.arange(n1j)[None, :]).reshape(-1)
    dC = dev.zeros((n0i, n1i, n0j, n1j))
    dev.ctx.gemm_scatter(n0i * n0j, n1i * n1j, K, 1.0, dev.upload(A), K, dev.upload(B), K, dC,
                         dev.upload(offM, numpy.int64), 0, dev.upload(offN, numpy.int64), False)
    ref = numpy.einsum("ijx,klx->ikjl", A.reshape(n0i, n0j, K), B.reshape(n1i, n1j, K))
    _close(dev.download(dC), ref, 1e-13)


@pytest.mark.parametrize("n,Pa,Pb,Pc,pad", [(5, 3, 4, 7, 0), (6, 20, 33, 130, 0), (18, 9, 17, 257, 0), (18, 40, 50, 1000, 0),
                                            (48, 5, 18, 140, 0),
                                            # several work items per persistent CTA (items = ceil(Pa/8)*ceil(Pb/16) > 148), with
                                            # fewer gamma tiles per item than ring slots (1, 3) and more (6): the producer
                                            # stages item i+1's W/beta rows while item i streams
                                            (6, 400, 200, 40, 0), (18, 130, 330, 300, 0), (18, 70, 530, 700, 0),
                                            # odd leading dimension of W: rows are not 16-byte aligned, no TMA staging
                                            (18, 130, 330, 300, 1), (6, 400, 200, 40, 3), (20, 9, 40, 129, 0), (48, 30, 100, 140, 0),
                                            # every (DMMA k-steps, DFMA tail) cover the dispatcher instantiates: n = 4*KS + TAIL
                                            (9, 12, 20, 150, 0), (10, 9, 33, 140, 0), (12, 17, 18, 129, 0), (13, 8, 16, 128, 0),
                                            (13, 20, 40, 130, 1), (14, 10, 30, 200, 0), (16, 24, 40, 260, 0), (17, 9, 17, 130, 0),
                                            (19, 9, 17, 130, 0), (24, 9, 20, 140, 0), (30, 10, 18, 130, 0), (36, 9, 17, 129, 0),
                                            (44, 8, 16, 200, 0), (26, 9, 17, 130, 0), (28, 8, 16, 129, 0), (34, 9, 17, 130, 0),
                                            (42, 9, 20, 140, 0)])
def test_trimer_stream(dev, n, Pa, Pb, Pc, pad):
    from qodeapplications_b200 import lib as xr
    rng = numpy.random.default_rng(n + Pa + Pb + Pc)
    ldw = n * n + pad
    Wp = rng.standard_normal((Pa, ldw))
    W, beta, gamma = numpy.ascontiguousarray(Wp[:, :n * n]), rng.standard_normal((Pb, n)), rng.standard_normal((Pc, n))
    ref = -1.5 * numpy.einsum("ars,br,cs->abc", W.reshape(Pa, n, n), beta, gamma, optimize=True)
    dW, dB, dG = dev.upload(Wp), dev.upload(beta), dev.upload(gamma)
    # materialise with a permuted layout [c, a, b]
    offA = numpy.arange(Pa, dtype=numpy.int64) * Pb
    offB = numpy.arange(Pb, dtype=numpy.int64)
    offC = numpy.arange(Pc, dtype=numpy.int64) * (Pa * Pb)
    dC = dev.zeros((Pc, Pa, Pb))
    dev.ctx.trimer_stream(n, Pa, Pb, Pc, -1.5, dW, ldw, dB, n, dG, n, 0, Pa, xr.TRIMER_MATERIALIZE, None, dC,
                          dev.upload(offA, numpy.int64), dev.upload(offB, numpy.int64), dev.upload(offC, numpy.int64))
    _close(dev.download(dC), ref.transpose(2, 0, 1), 1e-13 * n)
    # reduce, in two shards (the second one starts at an a that is not a multiple of the 8-row item height)
    mom = dev.zeros((2,))
    split = Pa // 2 + (1 if Pa > 20 else 0)
    for lo, hi in ((0, split), (split, Pa)):
        dev.ctx.trimer_stream(n, Pa, Pb, Pc, -1.5, dW, ldw, dB, n, dG, n, lo, hi, xr.TRIMER_REDUCE, mom, None, None, None, None)
    got = dev.download(mom)
    assert abs(got[0] - ref.sum()) <= 1e-11 * numpy.abs(ref).sum()
    assert abs(got[1] - (ref ** 2).sum()) <= 1e-12 * (ref ** 2).sum()
    total, sumsq = go.trimer_class_moments(W.reshape(Pa, n, n), beta, gamma)
    assert abs(got[1] - 2.25 * sumsq) <= 1e-11 * 2.25 * sumsq
    # bit-reproducible: the same launch twice gives the same moments
    mom2 = dev.zeros((2,))
    for lo, hi in ((0, split), (split, Pa)):
        dev.ctx.trimer_stream(n, Pa, Pb, Pc, -1.5, dW, ldw, dB, n, dG, n, lo, hi, xr.TRIMER_REDUCE, mom2, None, None, None, None)
    assert numpy.array_equal(dev.download(mom2), got)


@pytest.mark.parametrize("n,Pa,Pb,Pc", [(5, 3, 4, 7), (6, 400, 200, 40), (18, 130, 330, 300), (18, 70, 530, 700), (17, 9, 17, 130),
                                        (20, 9, 40, 129), (24, 9, 20, 140), (48, 30, 100, 140)])
def test_trimer_sample_and_threshold_consumers(dev, n, Pa, Pb, Pc):
    """xr_trimer_sample / xr_trimer_threshold: elements picked out of, and screened from, the streamed tiles"""
    rng = numpy.random.default_rng(n * 7 + Pa + Pb + Pc)
    W, beta, gamma = rng.standard_normal((Pa, n * n)), rng.standard_normal((Pb, n)), rng.standard_normal((Pc, n))
    ref = 0.5 * numpy.einsum("ars,br,cs->abc", W.reshape(Pa, n, n), beta, gamma, optimize=True)
    ldw = n * n + (n * n) % 2
    Wp = numpy.zeros((Pa, ldw))
    Wp[:, :n * n] = W
    dW, dB, dG = dev.upload(Wp), dev.upload(beta), dev.upload(gamma)
    count = min(500, Pa * Pb * Pc)
    abc = numpy.stack([rng.integers(Pa, size=count), rng.integers(Pb, size=count), rng.integers(Pc, size=count)], axis=1).astype(numpy.int64)
    abc[0], abc[-1] = (0, 0, 0), (Pa - 1, Pb - 1, Pc - 1)
    abc[1] = abc[2]                                             # a repeated request
    out = dev.empty((count,))
    dev.ctx.trimer_sample(n, Pa, Pb, Pc, 0.5, dW, ldw, dB, n, dG, n, numpy.ascontiguousarray(abc), out)
    _close(dev.download(out), ref[abc[:, 0], abc[:, 1], abc[:, 2]], 1e-13 * n)
    # screening: keep ~3 % of the elements, in a permuted layout [c, a, b], first with a list that is too short
    tau = float(numpy.quantile(numpy.abs(ref), 0.97))
    offA = numpy.arange(Pa, dtype=numpy.int64) * Pb
    offB = numpy.arange(Pb, dtype=numpy.int64)
    offC = numpy.arange(Pc, dtype=numpy.int64) * (Pa * Pb)
    tables = [dev.upload(o, numpy.int64) for o in (offA, offB, offC)]
    keep = numpy.flatnonzero(numpy.abs(ref.transpose(2, 0, 1).reshape(-1)) > tau)
    cnt = dev.zeros((1,), dtype=torch.int64)
    for cap in (max(1, len(keep) // 3), len(keep) + 5):
        idx, val = dev.empty((cap,), dtype=torch.int64), dev.empty((cap,))
        dev.ctx.trimer_threshold(n, Pa, Pb, Pc, 0.5, dW, ldw, dB, n, dG, n, 0, Pa, tau, tables[0], tables[1], tables[2], cap, idx, val, cnt)
        assert int(cnt.cpu()[0]) == len(keep)
    got_idx, got_val = idx.cpu().numpy()[:len(keep)], val.cpu().numpy()[:len(keep)]
    order = numpy.argsort(got_idx)
    assert numpy.array_equal(got_idx[order], keep)
    _close(got_val[order], ref.transpose(2, 0, 1).reshape(-1)[keep], 1e-13 * n)
    # a sub-range of a: the two halves partition the list
    half = Pa // 2
    total = 0
    for lo, hi in ((0, half), (half, Pa)):
        dev.ctx.trimer_threshold(n, Pa, Pb, Pc, 0.5, dW, ldw, dB, n, dG, n, lo, hi, tau, tables[0], tables[1], tables[2], 0, None, None, cnt)
        total += int(cnt.cpu()[0])
    assert total == len(keep)


def test_legacy_scalar_abi(dev):
    """The 11 H_contractions.c symbols, same ABI, against the CPU oracle's C library."""
    from qodeapplications_b200.general.H_contractions import import_C
    contract = import_C("H_contractions", flags="-O2")
    ref_path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libH_contractions_ref.so")
    ref = go.c_contractions("ref" if os.path.exists(ref_path) else "port")
    rng = numpy.random.default_rng(5)
    r = lambda *s: rng.standard_normal(s)
    for n1, n2, n3 in ((5, 4, 3), (18, 18, 18)):
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
            fn = getattr(contract, name)
            fn.return_type(float)
            a, b = fn(*args), getattr(ref, name)(*args)
            terms = max(numpy.abs(args[-1]).sum(), 1.0)
            assert abs(a - b) <= 1e-13 * terms, (name, a, b)


# ------------------------------------------------------------------------------ H blocks

@pytest.mark.parametrize("name", ["toy", "toy3", "toy5", "toyh", "toyh3"])
def test_blocks_match_reference_golden(dev, name):
    g = numpy.load(os.path.join(GOLDEN, "general_%s.npz" % name))
    system = synth.make_system(name)
    eng = _engine(system, dev)
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


def test_element_accessors_keep_reference_signatures(dev):
    """monomer/dimer/trimer(fragments, I, J) exactly as general-XRCC/test_H.py:51-58 calls them."""
    g = numpy.load(os.path.join(GOLDEN, "general_toy3.npz"))
    system = synth.make_system("toy3")
    eng = _engine(system, dev)
    st = [f.state_indices for f in system["fragments"]]
    rng = numpy.random.default_rng(1)
    H1 = g["H1_2"]
    for _ in range(10):
        i, j = rng.integers(len(st[2]), size=2)
        assert abs(eng.monomer(2, st[2][i], st[2][j]) - H1[i, j]) <= TOL * numpy.abs(H1).max()
    basis = [(a, b) for a in st[0] for b in st[1]]
    H2 = g["H2_01"]
    for _ in range(50):
        i, j = rng.integers(len(basis), size=2)
        assert abs(eng.dimer((0, 1), basis[i], basis[j]) - H2[i, j]) <= TOL * numpy.abs(H2).max()
    basis3 = [(a, b, c) for a in st[0] for b in st[1] for c in st[2]]
    ref = numpy.zeros(tuple(g["H3_012_shape"]))
    ref[g["H3_012_rows"], g["H3_012_cols"]] = g["H3_012_vals"]
    for _ in range(200):
        i, j = rng.integers(len(basis3), size=2)
        assert abs(eng.trimer((0, 1, 2), basis3[i], basis3[j]) - ref[i, j]) <= TOL * numpy.abs(ref).max()


def test_cfg1_dimer_against_block_oracle(dev):
    """Be2 / 6-31G shapes (n=18, N=11/4/8): full H1 and H2 against the NumPy block oracle."""
    system = synth.make_system("cfg1")
    eng = _engine(system, dev)
    frags, ints, nuc = system["fragments"], system["symm"], system["nuc"]
    for m in range(2):
        _close(eng.H1(m), go.block_monomer(frags, ints, nuc, m))
    _close(eng.H2(0, 1), go.block_dimer(frags, ints, nuc, 0, 1))


def test_cfg1_sampled_against_reference_c(dev):
    """The same H2 against the reference's own compiled C, element by element on a sample."""
    ref_path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libH_contractions_ref.so")
    system = synth.make_system("cfg1")
    eng = _engine(system, dev)
    frags = system["fragments"]
    eo = go.element_oracle(frags, system["symm"], system["nuc"], go.c_contractions("ref" if os.path.exists(ref_path) else "port"))
    H2 = eng.H2(0, 1)
    basis = [(a, b) for a in frags[0].state_indices for b in frags[1].state_indices]
    rng = numpy.random.default_rng(2)
    floor = 1e-3 * float(numpy.sqrt(numpy.mean(H2 * H2)))
    for _ in range(400):
        i, j = rng.integers(len(basis), size=2)
        ref = eo.dimer((0, 1), basis[i], basis[j]) or 0.0
        assert abs(H2[i, j] - ref) <= TOL * max(abs(ref), floor)


def test_bra_slab_sharding_reassembles_H2(dev):
    """Rows built per bra slab (what each rank of a multi-GPU run computes) concatenate to H2."""
    system = synth.make_system("toy")
    eng = _engine(system, dev)
    full = eng.H2(0, 1)
    dim1 = len(system["fragments"][0].state_indices)
    parts = []
    for lo, hi in ((0, 2), (2, 3), (3, dim1)):
        parts.append(dev.download(eng.H2_device(0, 1, bra_range=(lo, hi))))
    assert numpy.array_equal(numpy.concatenate(parts, axis=0), full)


def test_cfg3_trimer_moments_and_sampled_elements(dev):
    """Be3 chain shapes: streamed moments of H3 equal those of the materialised block, the shards add
    up, and sampled elements equal the reference C evaluated through the element oracle."""
    ref_path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libH_contractions_ref.so")
    system = synth.make_system("cfg3")
    eng = _engine(system, dev)
    H3 = eng.H3_device(0, 1, 2)
    total, sumsq = float(H3.sum()), float((H3 * H3).sum())
    s, q = eng.H3_moments(0, 1, 2)
    assert abs(q - sumsq) <= 1e-11 * sumsq
    assert abs(s - total) <= 1e-9 * float(H3.abs().sum())
    parts = [eng.H3_moments(0, 1, 2, shard=(r, 3)) for r in range(3)]
    assert abs(sum(p[1] for p in parts) - sumsq) <= 1e-11 * sumsq
    frags = system["fragments"]
    eo = go.element_oracle(frags, system["symm"], system["nuc"], go.c_contractions("ref" if os.path.exists(ref_path) else "port"))
    st = [f.state_indices for f in frags]
    dims = [len(s_) for s_ in st]
    rng = numpy.random.default_rng(4)
    floor = 1e-3 * float((H3 * H3).mean().sqrt())
    D = dims[0] * dims[1] * dims[2]
    nz = torch.nonzero(H3.reshape(-1))[:, 0]
    picks = nz[torch.randint(len(nz), (300,), device=nz.device)].cpu().numpy()
    picks = numpy.concatenate([picks, rng.integers(D * D, size=100)])
    vals = H3.reshape(-1)[torch.from_numpy(picks).to(H3.device)].cpu().numpy()
    for flat, val in zip(picks, vals):
        i, j = divmod(int(flat), D)
        I = numpy.unravel_index(i, dims)
        J = numpy.unravel_index(j, dims)
        ref = eo.trimer((0, 1, 2), tuple(st[k][I[k]] for k in range(3)), tuple(st[k][J[k]] for k in range(3)))
        assert abs(val - (ref or 0.0)) <= TOL * max(abs(ref or 0.0), floor)


CFG4_STATES = {0: 96, +1: 34, -1: 70}


@pytest.fixture(scope="module")
def cfg4_three(dev):
    """three fragments of the headline configuration (200 states each, n = 18): one dimer and one trimer at FULL size"""
    system = synth.make_system(n_frag=3, n_orb=18, n_states=CFG4_STATES, seed=4, ops=synth.OPS_GENERAL, general_ccaa="random")
    ref_path = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libH_contractions_ref.so")
    eo = go.element_oracle(system["fragments"], system["symm"], system["nuc"],
                           go.c_contractions("ref" if os.path.exists(ref_path) else "port"))
    return system, _engine(system, dev), eo


@pytest.mark.timeout(900)
def test_cfg4_size_dimer_sampled_against_reference_c(dev, cfg4_three):
    """One dense 40 000 x 40 000 H2 of the benchmark configuration, built on the GPU exactly as bench.py builds it; 600 elements
    (400 drawn from the charge-allowed classes, 200 anywhere) against the reference's compiled C, per element."""
    system, eng, eo = cfg4_three
    frags = system["fragments"]
    H2 = eng.H2_device(0, 1)
    D = H2.shape[0]
    assert D == 40000
    floor = 1e-3 * float((H2 * H2).mean().sqrt())
    st = [f.state_indices for f in frags]
    rng = numpy.random.default_rng(44)
    gen = torch.Generator(device=H2.device).manual_seed(44)
    # charge-allowed elements: rows/columns drawn uniformly, kept when non-zero on the device
    cand = torch.randint(D * D, (4000,), device=H2.device, generator=gen)
    vals = H2.reshape(-1)[cand]
    nz = cand[vals != 0][:400]
    picks = numpy.concatenate([nz.cpu().numpy(), rng.integers(D * D, size=200)])
    got = H2.reshape(-1)[torch.from_numpy(picks).to(H2.device)].cpu().numpy()
    d2 = len(st[1])
    worst = 0.0
    for flat, val in zip(picks, got):
        i, j = divmod(int(flat), D)
        I = (st[0][i // d2], st[1][i % d2])
        J = (st[0][j // d2], st[1][j % d2])
        ref = eo.dimer((0, 1), I, J) or 0.0
        assert abs(val - ref) <= TOL * max(abs(ref), floor), (I, J, val, ref)
        if ref == 0.0:
            assert val == 0.0
        else:
            worst = max(worst, abs(val - ref) / max(abs(ref), floor))
    assert len(nz) == 400
    del H2


@pytest.mark.timeout(900)
def test_cfg4_size_trimer_class_moments_against_gram_identity(dev, cfg4_three):
    """Every charge-transfer class of a full-size trimer (1.06e13 elements) streamed through xr_trimer_stream: the sum of
    squares of the streamed tiles equals the exact Gram-matrix value of the class factors (O((Pa+Pb+Pc) n^4) on the host),
    the first moment equals the factor-sum value, and two ranks' shards add up to the whole."""
    system, eng, eo = cfg4_three
    ms = (0, 1, 2)
    got = eng.H3_moments(*ms, per_class=True)
    halves = sum(eng.H3_moments(*ms, shard=(r, 2), per_class=True) for r in range(2))
    seen = 0
    for idx, cl in enumerate(eng._trimer_classes(ms)):
        fac = eng._trimer_factors(ms, cl)
        if fac is None:
            continue
        n = fac["n"]
        W = dev.download(fac["W"])[:, :n * n].reshape(-1, n, n)
        beta, gamma = dev.download(fac["beta"])[:, :n], dev.download(fac["gamma"])[:, :n]
        total, sumsq = go.trimer_class_moments(W, beta, gamma)
        assert abs(got[idx, 1] - sumsq) <= 1e-11 * sumsq, (idx, cl["kind"])
        assert abs(got[idx, 0] - fac["alpha"] * total) <= 1e-9 * (sumsq * W.shape[0] * beta.shape[0] * gamma.shape[0]) ** 0.5
        assert abs(halves[idx, 1] - got[idx, 1]) <= 1e-12 * got[idx, 1]
        seen += 1
    assert seen == 12


@pytest.mark.parametrize("name", ["toy3", "toyh3", "cfg3"])
def test_trimer_consumers_against_dense_block(dev, name):
    """H3_elements / H3_sparse (block level, all 12 classes, signs and offsets) against the materialised H3"""
    system = synth.make_system(name)
    eng = _engine(system, dev)
    dense = eng.H3_device(0, 1, 2).reshape(-1)
    st = [f.state_indices for f in system["fragments"]]
    dims = [len(x) for x in st]
    D = int(numpy.prod(dims))
    rng = numpy.random.default_rng(6)
    nz = torch.nonzero(dense)[:, 0]
    flat = numpy.concatenate([rng.integers(D * D, size=200), nz[torch.randint(len(nz), (300,), device=nz.device)].cpu().numpy()])
    I = [tuple(st[k][x] for k, x in enumerate(numpy.unravel_index(int(e) // D, dims))) for e in flat]
    J = [tuple(st[k][x] for k, x in enumerate(numpy.unravel_index(int(e) % D, dims))) for e in flat]
    want = dense[torch.from_numpy(flat).to(dense.device)].cpu().numpy()
    assert numpy.array_equal(eng.H3_elements(0, 1, 2, I, J), want)          # the same tile arithmetic: bit-identical
    tau = float(dense.abs().max()) * 0.5
    idx, val = eng.H3_sparse(0, 1, 2, tau, capacity=64)
    keep = torch.nonzero(dense.abs() > tau)[:, 0]
    assert numpy.array_equal(idx, keep.cpu().numpy())
    assert numpy.array_equal(val, dense[keep].cpu().numpy())


@pytest.mark.timeout(900)
def test_cfg4_size_trimer_elements_against_reference_c(dev, cfg4_three):
    """Elements of a full-size trimer block (8e6 x 8e6: it cannot exist) picked out of the streamed tiles by the
    sampled-element consumer, against the reference's compiled trimer_* C functions, per element."""
    system, eng, eo = cfg4_three
    frags = system["fragments"]
    st = [f.state_indices for f in frags]
    rng = numpy.random.default_rng(45)
    chg = {c: [s for s in st[0] if s[0] == c] for c in CFG4_STATES}
    I, J = [], []
    patterns = [p for kind in ((-2, 1, 1), (2, -1, -1), (0, -1, 1)) for p in sorted(set(itertools.permutations(kind)))]
    while len(I) < 360:
        pat = patterns[len(I) % len(patterns)]
        bra = [int(rng.choice(list(CFG4_STATES))) for _ in range(3)]
        ket = [b - d for b, d in zip(bra, pat)]
        if not all(k in CFG4_STATES for k in ket):
            continue
        I.append(tuple((c, int(rng.integers(CFG4_STATES[c]))) for c in bra))
        J.append(tuple((c, int(rng.integers(CFG4_STATES[c]))) for c in ket))
    for _ in range(40):                                           # charge-forbidden elements: exactly zero
        I.append(tuple((0, int(rng.integers(96))) for _ in range(3)))
        J.append(tuple((0, int(rng.integers(96))) for _ in range(3)))
    got = eng.H3_elements(0, 1, 2, I, J)
    ref = numpy.array([eo.trimer((0, 1, 2), i, j) or 0.0 for i, j in zip(I, J)])
    assert numpy.count_nonzero(ref) >= 360
    floor = 1e-3 * float(numpy.sqrt(numpy.mean(ref[ref != 0] ** 2)))
    assert (numpy.abs(got - ref) <= TOL * numpy.maximum(numpy.abs(ref), floor)).all(), float(numpy.abs(got - ref).max())
    assert numpy.all(got[ref == 0] == 0)


def test_gemm_dd_newton_polish_of_an_inverse(dev):
    """xr_gemm_dd: residual I - M X and the update X + X R against numpy.longdouble (what the host polish used)"""
    rng = numpy.random.default_rng(12)
    n = 203
    M = numpy.eye(n) + 0.05 * rng.standard_normal((n, n))
    X = numpy.linalg.inv(M)
    ML, XL = M.astype(numpy.longdouble), X.astype(numpy.longdouble)
    RL = numpy.eye(n, dtype=numpy.longdouble) - ML @ XL
    dM, dX = dev.upload(M), dev.upload(X)
    R, out = dev.empty((n, n)), dev.empty((n, n))
    dev.ctx.gemm_dd(n, n, n, dM, n, dX, n, None, 0, -1.0, R, n)
    got = dev.download(R)
    assert numpy.abs(got - RL.astype(numpy.float64)).max() <= 1e-19 * n * numpy.abs(X).max()     # plain FP64 would be ~1e-16
    assert numpy.abs(got).max() < 1e-13
    dev.ctx.gemm_dd(n, n, n, dX, n, R, n, dX, n, +1.0, out, n)
    polished = (XL + XL @ RL).astype(numpy.float64)
    assert numpy.abs(dev.download(out) - polished).max() <= 2e-16 * numpy.abs(X).max()
    # rectangular, with leading dimensions and a C0 operand
    A, B, C0 = rng.standard_normal((37, 60)), rng.standard_normal((50, 29)), rng.standard_normal((37, 31))
    o = dev.zeros((37, 33))
    dev.ctx.gemm_dd(37, 29, 50, dev.upload(A), 60, dev.upload(B), 29, dev.upload(C0), 31, -1.0, o, 33)
    ref = (C0[:, :29].astype(numpy.longdouble) - A[:, :50].astype(numpy.longdouble) @ B.astype(numpy.longdouble)).astype(numpy.float64)
    res = dev.download(o)
    assert numpy.abs(res[:, :29] - ref).max() <= 1e-15 * numpy.abs(ref).max()
    assert numpy.all(res[:, 29:] == 0)


# ------------------------------------------------------------------------------- streamed dimers (cfg5 path)

@pytest.mark.parametrize("M,N,K", [(1, 1, 2), (7, 5, 3), (64, 64, 16), (130, 70, 37), (257, 300, 326), (1000, 333, 36),
                                   (3000, 2100, 2306)])
def test_gemm_reduce(dev, M, N, K):
    rng = numpy.random.default_rng(M + 7 * N + 13 * K)
    lda = K + K % 2
    ldb = lda + 2
    A, B = rng.standard_normal((M, lda)), rng.standard_normal((N, ldb))
    C = -0.5 * A[:, :K] @ B[:, :K].T
    moments = dev.upload(numpy.array([[1.0, 2.0], [0.0, 0.0]]))
    dev.ctx.gemm_reduce(M, N, K, -0.5, dev.upload(A), lda, dev.upload(B), ldb, moments)
    out = dev.download(moments)
    assert abs(out[0, 1] - 2.0 - (C * C).sum()) <= 1e-12 * (C * C).sum()
    assert abs(out[0, 0] - 1.0 - C.sum()) <= 1e-12 * numpy.abs(C).sum()
    assert numpy.array_equal(out[1], [0.0, 0.0])
    again = dev.upload(numpy.array([[1.0, 2.0], [0.0, 0.0]]))
    dev.ctx.gemm_reduce(M, N, K, -0.5, dev.upload(A), lda, dev.upload(B), ldb, again)
    assert numpy.array_equal(dev.download(again), out), "the reduction must be deterministic"


@pytest.mark.parametrize("name", ["toy", "toy5", "toyh", "cfg1"])
def test_streamed_dimer_moments_equal_dense_block(dev, name):
    system = synth.make_system(name)
    eng = _engine(system, dev)
    H2 = eng.H2_device(0, 1)
    per_class = dev.download(eng.H2_moments_device(0, 1))
    sumsq, total = float((H2 * H2).sum()), float(H2.sum())
    assert abs(per_class[:, 1].sum() - sumsq) <= 1e-12 * sumsq
    assert abs(per_class[:, 0].sum() - total) <= 1e-11 * float(H2.abs().sum())


def test_streamed_dimer_from_device_resident_bra_slabs(dev):
    """Each 'rank' holds only its bra slab of the densities, already on the device (the cfg5 input layout);
    the slabs' class moments add up to those of the whole block."""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import slab_bounds
    name = "mid"
    system = synth.make_system(name)
    n_states = synth.CONFIGS[name]["n_states"]
    whole = dev.download(_engine(system, dev).H2_moments_device(0, 1))
    dim = len(system["fragments"][0].state_indices)
    world, acc = 3, numpy.zeros((5, 2))
    for rank in range(world):
        lo, hi = slab_bounds(dim, rank, world)[:2]
        frags = list(system["fragments"])
        frags[0] = synth.slab_fragment(frags[0], (lo, hi), n_states)
        for op in frags[0].rho:
            for key, block in frags[0].rho[op].items():
                frags[0].rho[op][key] = dev.upload(block)          # device-resident input
        eng = build_matrix_elements(frags, system["symm"], system["nuc"], device=dev, held={0: (lo, hi)})
        part = dev.zeros((5, 2))
        for d1, c1, c2, A, B, K, ld in eng._dimer_class_factors(0, 1, (lo, hi), None):
            dev.ctx.gemm_reduce(c1.P, c2.P, K, 1.0, A, ld, B, ld, part.data_ptr() + 16 * (d1 + 2))
        acc += dev.download(part)
        with pytest.raises(ValueError):
            eng.H2_device(0, 1)                                     # needs bra states this engine does not hold
    assert numpy.allclose(acc, whole, rtol=1e-12, atol=1e-12 * numpy.abs(whole).max())


RAGGED = [{0: 2, +1: 1, -1: 0}, {0: 1, +1: 1, -1: 1}, {0: 3, +1: 0, -1: 0}, {0: 2, +1: 0, -1: 2}, {0: 1}]


@pytest.mark.parametrize("n_states", RAGGED, ids=lambda d: "-".join("%+d:%d" % kv for kv in d.items()))
def test_ragged_and_empty_charge_sectors(dev, n_states):
    """charge sectors with no states, one state, or a missing charge: zero-size factor matrices reach the library"""
    system = synth.make_system(n_frag=3, n_orb=4, n_states=n_states, seed=77)
    fr, ints, nuc = system["fragments"], system["symm"], system["nuc"]
    eng = _engine(system, dev)
    for m in range(3):
        _close(eng.H1(m), go.block_monomer(fr, ints, nuc, m), 1e-12)
    for m1, m2 in itertools.combinations(range(3), 2):
        _close(eng.H2(m1, m2), go.block_dimer(fr, ints, nuc, m1, m2), 1e-12)
    ref = go.block_trimer(fr, ints, (0, 1, 2))
    _close(eng.H3(0, 1, 2), ref, 1e-12)
    s, q = eng.H3_moments(0, 1, 2)
    assert abs(q - (ref ** 2).sum()) <= 1e-10 * (ref ** 2).sum() and abs(s - ref.sum()) <= 1e-10 * max(numpy.abs(ref).sum(), 1.0)


def test_zero_size_arguments_are_no_ops(dev):
    """every block-level entry point returns XR_OK without touching memory when an extent is zero (null buffers allowed)"""
    from qodeapplications_b200 import lib as xr
    c = dev.ctx
    before = c.launch_count()
    c.gemm_scatter(0, 5, 3, 1.0, None, 3, None, 3, None, None, 5)
    c.gemm_scatter(5, 0, 3, 1.0, None, 3, None, 3, None, None, 1)
    c.gemm_reduce(0, 4, 2, 1.0, None, 2, None, 2, None)
    c.copy2d_scaled(None, 4, None, 4, 0, 4)
    c.permute_copy(None, None, (3, 0, 2), (0, 2, 1))
    c.scatter_const(None, None, 0, 1.0)
    c.trimer_stream(4, 3, 0, 5, 1.0, None, 16, None, 4, None, 4, 0, 3, xr.TRIMER_REDUCE, None)
    c.trimer_stream(4, 3, 2, 5, 1.0, None, 16, None, 4, None, 4, 2, 2, xr.TRIMER_REDUCE, None)
    c.embed_add(None, None, 1, 0, 0, 1, None, None)
    assert c.launch_count() == before


@pytest.mark.parametrize("name", ["toy3", "cfg3"])
def test_recorded_step_is_one_cuda_graph(dev, name):
    """general/distributed.recorded_step on the GPU: the recorded build step as a CUDA graph (single- and multi-stream) gives
    the eager step's blocks bit for bit, and follows new densities"""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.distributed import sharded_build
    system = synth.make_system(name)
    eng = _engine(system, dev)
    dimers = list(itertools.combinations(range(3), 2))
    build = sharded_build(eng, dimers, [(0, 1, 2)])
    build.step()
    want = [build.full(*k).clone() for k in dimers] + [build.H3_moments[(0, 1, 2)].clone(), build.H1[0].clone()]
    for streams in (1, 16):
        rec = build.recorded(streams=streams)
        assert rec.graph is not None and rec.n_streams == streams, (rec._launcher.graph_error, rec._launcher.streams_error)
        for _ in range(2):
            rec.run()
        got = [build.full(*k) for k in dimers] + [build.H3_moments[(0, 1, 2)], build.H1[0]]
        assert all(torch.equal(a, b) for a, b in zip(got, want)), streams
    other = synth.make_system(name, seed=99)
    rec.update(other["fragments"])
    rec.run()
    fresh = build_matrix_elements(other["fragments"], system["symm"], system["nuc"], device=dev)
    assert numpy.array_equal(dev.download(build.full(0, 2)), fresh.H2(0, 2))
    assert numpy.array_equal(dev.download(build.H3_moments[(0, 1, 2)]), fresh.H3_moments(0, 1, 2, per_class=True))
