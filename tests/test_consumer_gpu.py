"""GPU parity of the consumer (xr_embed_add behind general/hamiltonian.py) against the reference's own braket_loops
outputs (tests/golden/supersystem_hmat.npz) and the oracle restatement."""
import itertools
import os
import sys
import numpy
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLDEN = os.path.join(HERE, "golden")

from oracle import general_oracle as go
from gen_golden_hmat import random_blocks, toy3_blocks
from qodeapplications_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from qodeapplications_b200.device import Device
    return Device(0)


@pytest.mark.parametrize("tag", ["rand4x3", "rand3x4", "herm3x4", "herm2x5"])
def test_braket_loops_matches_reference(dev, tag):
    from qodeapplications_b200.general.hamiltonian import braket_loops
    g = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))
    seed, F, spf, trimers = (int(x) for x in g[tag + "_meta"])
    H = random_blocks(seed, F, spf, bool(trimers))
    Hmat = numpy.zeros((spf ** F, spf ** F))
    braket_loops(Hmat, F, spf, H, device=dev)
    ref = g[tag]
    assert numpy.abs(Hmat - ref).max() <= 1e-14 * numpy.abs(ref).max()
    assert numpy.all(Hmat[ref == 0] == 0)
    on_device = dev.zeros((spf ** F, spf ** F))                   # CUDA tensor in, nothing copied to the host
    braket_loops(on_device, F, spf, H, device=dev)
    assert numpy.array_equal(dev.download(on_device), Hmat)


def test_build_then_expand_stays_on_device(dev):
    """toy3: blocks built by the B200 engine go straight into the supersystem matrix (never on the host) and the result
    equals the reference's braket_loops applied to the reference's own blocks."""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.hamiltonian import supersystem_matrix
    system = synth.make_system("toy3")
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    F = 3
    dims = [len(f.state_indices) for f in system["fragments"]]
    H1 = [eng.H1_device(m) for m in range(F)]
    H2 = [[eng.H2_device(M, N) if M < N else None for N in range(F)] for M in range(F)]
    H3 = [[[eng.H3_device(M, N, O) if (M, N, O) == (0, 1, 2) else None for O in range(F)] for N in range(F)] for M in range(F)]
    h2d = dev.h2d_bytes
    big = supersystem_matrix(dims, dev).add_all((H1, H2, H3))
    assert dev.h2d_bytes - h2d < 64 * 1024                          # offset tables only
    ref = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))["toy3"]
    out = dev.download(big.matrix)
    assert numpy.abs(out - ref).max() <= 1e-10 * numpy.abs(ref).max()


def test_unequal_dims_and_cfg3_size(dev):
    from qodeapplications_b200.general.hamiltonian import supersystem_matrix
    rng = numpy.random.default_rng(8)
    dims = [5, 3, 4, 2]
    F = len(dims)
    d = lambda *ms: int(numpy.prod([dims[m] for m in ms]))
    H1 = [rng.standard_normal((d(m), d(m))) for m in range(F)]
    H2 = [[rng.standard_normal((d(M, N),) * 2) if M < N else None for N in range(F)] for M in range(F)]
    H3 = [[[rng.standard_normal((d(M, N, O),) * 2) if M < N < O else None for O in range(F)] for N in range(F)] for M in range(F)]
    big = supersystem_matrix(dims, dev).add_all((H1, H2, H3))
    ref = go.supersystem_matrix(dims, (H1, H2, H3))
    assert numpy.abs(dev.download(big.matrix) - ref).max() <= 1e-14 * numpy.abs(ref).max()
    # Be3-chain size (23 states per fragment, 12167^2 matrix = 1.2 GB): H.v through the expanded matrix equals the sum of
    # the block actions (torch einsum as the checker)
    spf, F = 23, 3
    H1 = [torch.randn((spf, spf), dtype=torch.float64, device=dev.torch_device) for _ in range(F)]
    H2 = [[torch.randn((spf ** 2, spf ** 2), dtype=torch.float64, device=dev.torch_device) if M < N else None for N in range(F)]
          for M in range(F)]
    big = supersystem_matrix([spf] * F, dev)
    for M in range(F):
        big.add((M,), H1[M])
    for M, N in itertools.combinations(range(F), 2):
        big.add((M, N), H2[M][N])
    v = torch.randn((spf,) * F, dtype=torch.float64, device=dev.torch_device)
    y = torch.einsum("ai,ijk->ajk", H1[0], v) + torch.einsum("bj,ijk->ibk", H1[1], v) + torch.einsum("ck,ijk->ijc", H1[2], v)
    y = y + torch.einsum("abij,ijk->abk", H2[0][1].reshape((spf,) * 4), v) + torch.einsum("acik,ijk->ajc", H2[0][2].reshape((spf,) * 4), v)
    y = y + torch.einsum("bcjk,ijk->ibc", H2[1][2].reshape((spf,) * 4), v)
    got = (big.matrix @ v.reshape(-1)).reshape((spf,) * F)
    assert float((got - y).abs().max()) <= 1e-11 * float(y.abs().max())


def test_matrix_free_operator_against_expanded_matrix(dev):
    """cfg3 (Be3 chain shapes, 23 states/fragment): y = Hmat.v from the class factors equals the action of the matrix
    expanded from the dense blocks (xr_embed_add), and the toy3 operator equals the reference's braket_loops matrix."""
    from qodeapplications_b200.general.build_H import build_matrix_elements
    from qodeapplications_b200.general.hamiltonian import supersystem_matrix
    from qodeapplications_b200.general.operator import xr_operator
    system = synth.make_system("toy3")
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    ref = numpy.load(os.path.join(GOLDEN, "supersystem_hmat.npz"))["toy3"]
    got = dev.download(xr_operator(eng).dense())
    assert numpy.abs(got - ref).max() <= 1e-10 * numpy.abs(ref).max()

    system = synth.make_system("cfg3")
    eng = build_matrix_elements(system["fragments"], system["symm"], system["nuc"], device=dev)
    F = 3
    dims = [len(f.state_indices) for f in system["fragments"]]
    H1 = [eng.H1_device(m) for m in range(F)]
    H2 = [[eng.H2_device(M, N) if M < N else None for N in range(F)] for M in range(F)]
    H3 = [[[eng.H3_device(M, N, O) if (M, N, O) == (0, 1, 2) else None for O in range(F)] for N in range(F)] for M in range(F)]
    big = supersystem_matrix(dims, dev).add_all((H1, H2, H3)).matrix
    v = torch.randn(tuple(dims) + (3,), dtype=torch.float64, device=dev.torch_device)
    op = xr_operator(eng)
    y = op.apply(v)
    want = (big @ v.reshape(-1, 3)).reshape(y.shape)
    assert float((y - want).abs().max()) <= 1e-10 * float(want.abs().max())
    # sum of all trimer elements two ways: the operator on the all-ones vector vs the streamed moments of the tile kernel
    ones = torch.ones(tuple(dims), dtype=torch.float64, device=dev.torch_device)
    total = float(xr_operator(eng, monomers=False, dimers=False).apply(ones).sum())
    streamed, sumsq = eng.H3_moments(0, 1, 2)
    assert abs(total - streamed) <= 1e-9 * (sumsq * float(numpy.prod(dims)) ** 2) ** 0.5
