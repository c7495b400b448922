"""GPU parity of the density-tensor build (xr_density_tensor / the eight legacy density_tensors.c symbols) against the
reference's own outputs: bit-exact, because the kernel keeps the reference's summation order and rounding."""
import ctypes
import os
import sys
import numpy
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
GOLDEN = numpy.load(os.path.join(HERE, "golden", "density_tensors.npz"))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libdensity_tensors_ref.so")

from oracle import density_oracle as do
from oracle import general_oracle as go
from gen_golden_density import CASE, inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from qodeapplications_b200.device import Device
    return Device(0)


def test_build_density_tensors_matches_reference_bit_for_bit(dev):
    from qodeapplications_b200.general.build_density_tensors import build_density_tensors
    z_lists, V = inputs()
    densities, total = build_density_tensors(z_lists, CASE["n_orbs"], V, CASE["n_core"], device=dev)
    assert total == int(GOLDEN["total_size"])
    for op, blocks in densities.items():
        for (bra, ket), nested in blocks.items():
            want = GOLDEN["%s_%d_%d" % (op, bra, ket)]
            if op == "ccaa":
                assert numpy.abs(numpy.array(nested) - want).max() <= 1e-13 * numpy.abs(want).max()
            else:
                assert numpy.array_equal(numpy.array(nested), want), (op, bra, ket)


def test_legacy_density_symbols_keep_the_reference_abi(dev):
    """the eight density_tensors.c entry points, called with build_density_tensors.py:70-157's argument lists"""
    from qodeapplications_b200 import lib as xr
    lib = xr.load()
    z_lists, _ = inputs()
    charges = list(z_lists)
    idx = {chg: i for i, chg in enumerate(charges)}
    n_orbs, n_core = CASE["n_orbs"], CASE["n_core"]
    i64, dbl = ctypes.c_int64, ctypes.c_double
    n_elec = numpy.array([z_lists[c].configs.shape[1] for c in charges], dtype=numpy.int64)
    n_configs = numpy.array([z_lists[c].configs.shape[0] for c in charges], dtype=numpy.int64)
    n_states = numpy.array([z_lists[c].coeffs.shape[0] for c in charges], dtype=numpy.int64)
    configs = [numpy.ascontiguousarray(z_lists[c].configs, dtype=numpy.int64) for c in charges]
    coeffs = [numpy.ascontiguousarray(z_lists[c].coeffs, dtype=numpy.float64) for c in charges]
    combos = [do.comb_table(int(e) - 2 * n_core, 2 * (n_orbs - n_core)) for e in n_elec]
    ptrs = lambda arrays, ct: (ctypes.POINTER(ct) * len(arrays))(*[a.ctypes.data_as(ctypes.POINTER(ct)) for a in arrays])
    for key in GOLDEN.files:
        if key == "total_size" or key.startswith("ccaa"):
            continue
        op, bra, ket = key.split("_")
        bra, ket = int(bra), int(ket)
        storage = numpy.zeros(GOLDEN[key].size)
        getattr(lib, op + "_tensor")(storage.ctypes.data, idx[bra], idx[ket], n_elec.ctypes.data, n_states.ctypes.data,
                                     ctypes.cast(ptrs(coeffs, dbl), ctypes.c_void_p), n_configs.ctypes.data,
                                     ctypes.cast(ptrs(configs, i64), ctypes.c_void_p), n_orbs, n_core,
                                     ctypes.cast(ptrs(combos, i64), ctypes.c_void_p), 1)
        assert numpy.array_equal(storage.reshape(GOLDEN[key].shape), GOLDEN[key]), key
    # the reference ACCUMULATES into storage: a second call doubles it exactly
    storage = GOLDEN["ca_0_0"].reshape(-1).copy()
    lib.ca_tensor(storage.ctypes.data, idx[0], idx[0], n_elec.ctypes.data, n_states.ctypes.data,
                  ctypes.cast(ptrs(coeffs, dbl), ctypes.c_void_p), n_configs.ctypes.data, ctypes.cast(ptrs(configs, i64), ctypes.c_void_p),
                  n_orbs, n_core, ctypes.cast(ptrs(combos, i64), ctypes.c_void_p), 1)
    assert numpy.allclose(storage, 2 * GOLDEN["ca_0_0"].reshape(-1), rtol=1e-15, atol=1e-16)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
def test_be_631g_shapes_against_reference_c(dev):
    """Be/6-31G fragment shapes (9 spatial orbitals, 1 frozen core, 11/4/8 states of 4/3/5 electrons; 120/16/560
    configurations): every tensor equals the reference's compiled C bit for bit."""
    n_orbs, n_core = 9, 1
    z_lists = do.make_states(n_orbs, n_core, 4, {0: 11, +1: 4, -1: 8}, seed=21)
    ref = do.reference_c()
    masks = {c: dev.upload(numpy.bitwise_or.reduce(numpy.left_shift(numpy.uint64(1), z_lists[c].configs.astype(numpy.uint64)), axis=1)
                           .view(numpy.int64), numpy.int64) for c in z_lists}
    z = {c: dev.upload(z_lists[c].coeffs) for c in z_lists}
    for op in do.OPS:
        for bra in z_lists:
            ket = bra - do.op_dchg(op)
            if ket not in z_lists or (op == "ccaa" and bra != +1):      # ccaa (0,0) is 100 MB: one small sector is enough here
                continue
            want = ref.tensor(op, z_lists, bra, ket, n_orbs, n_core)
            rho = dev.zeros(want.shape)
            dev.ctx.density_tensor(op, rho, want.shape[0], want.shape[1], z[bra], z[bra].shape[1], z[ket], z[ket].shape[1], masks[ket],
                                   z_lists[bra].configs.shape[1], z_lists[ket].configs.shape[1], n_orbs, n_core)
            assert numpy.array_equal(dev.download(rho), want), (op, bra, ket)


def test_ci_vectors_to_hamiltonian_blocks_without_leaving_the_gpu(dev):
    """CI vectors -> densities (device_result=True) -> build_matrix_elements: H1 and H2 equal the element oracle fed with
    the REFERENCE's densities (tests/golden/density_tensors.npz), and nothing density-sized crosses PCIe."""
    from qodeapplications_b200 import synth
    from qodeapplications_b200.general.build_density_tensors import build_density_tensors
    from qodeapplications_b200.general.build_H import build_matrix_elements
    z_lists, _ = inputs()
    n_orbs, n_core, n = CASE["n_orbs"], CASE["n_core"], 2 * CASE["n_orbs"]
    symm, nuc = synth.make_integrals(2, n, numpy.random.default_rng(9))
    n_states = {chg: z_lists[chg].coeffs.shape[0] for chg in z_lists}
    order = synth.general_state_indices(n_states)
    device_frags, host_frags = [], []
    for m in range(2):
        rho_dev, _ = build_density_tensors(z_lists, n_orbs, symm.V[m, m, m, m], n_core, device=dev, device_result=True)
        device_frags.append(synth.fragment(rho_dev, CASE["n_elec_ref"], order))
        rho_ref = {}
        for key in GOLDEN.files:
            if key != "total_size" and not key.startswith("ccaa"):
                op, bra, ket = key.split("_")
                rho_ref.setdefault(op, {})[int(bra), int(ket)] = GOLDEN[key]
        ccaa = {}
        for chg in z_lists:     # the golden ccaa scalars belong to the fixture's own V: recompute them for this fragment's V
            t = do.tensor("ccaa", z_lists, chg, chg, n_orbs, n_core)
            ccaa[chg, chg] = numpy.einsum("pqrs,ijpqsr->ij", symm.V[m, m, m, m], t)
        rho_ref["ccaa"] = ccaa
        host_frags.append(synth.fragment(rho_ref, CASE["n_elec_ref"], order))
    h2d = dev.h2d_bytes
    eng = build_matrix_elements(device_frags, symm, nuc, device=dev)
    H1, H2 = eng.H1(0), eng.H2(0, 1)
    assert dev.h2d_bytes - h2d < 4 * n ** 4 * 8 * 4       # integrals and offset tables only
    ref1 = go.block_monomer(host_frags, symm, nuc, 0)
    ref2 = go.block_dimer(host_frags, symm, nuc, 0, 1)
    assert numpy.abs(H1 - ref1).max() <= 1e-10 * numpy.abs(ref1).max()
    assert numpy.abs(H2 - ref2).max() <= 1e-10 * numpy.abs(ref2).max()


def test_contracted_density_equals_tensor_times_weights(dev):
    """xr_density_contracted (the ccaa tensor reduced with V without ever being stored) against the stored tensor"""
    z_lists, V = inputs()
    n_orbs, n_core, n = CASE["n_orbs"], CASE["n_core"], 2 * CASE["n_orbs"]
    masks = {c: dev.upload(numpy.bitwise_or.reduce(numpy.left_shift(numpy.uint64(1), z_lists[c].configs.astype(numpy.uint64)), axis=1)
                           .view(numpy.int64), numpy.int64) for c in z_lists}
    z = {c: dev.upload(z_lists[c].coeffs) for c in z_lists}
    for op, bra, ket in (("ccaa", 0, 0), ("ccaa", -1, -1), ("ca", 1, 1), ("caa", 0, -1)):
        T = n ** len(op)
        w = numpy.random.default_rng(len(op) + bra).standard_normal(T)
        Nb, Nk = z[bra].shape[0], z[ket].shape[0]
        args = (Nb, Nk, z[bra], z[bra].shape[1], z[ket], z[ket].shape[1], masks[ket], z_lists[bra].configs.shape[1],
                z_lists[ket].configs.shape[1], n_orbs, n_core)
        rho = dev.empty((Nb * Nk, T))
        dev.ctx.density_tensor(op, rho, *args)
        want = dev.download(rho) @ w
        out = dev.upload(numpy.full(Nb * Nk, 0.5))
        dev.ctx.density_contracted(op, out, dev.upload(w), *args, accumulate=True)
        got = dev.download(out)
        assert numpy.abs(got - 0.5 - want).max() <= 1e-13 * max(1.0, numpy.abs(want).max())
        again = dev.empty((Nb * Nk,))
        dev.ctx.density_contracted(op, again, dev.upload(w), *args)
        twice = dev.empty((Nb * Nk,))
        dev.ctx.density_contracted(op, twice, dev.upload(w), *args)
        assert numpy.array_equal(dev.download(again), dev.download(twice)), "fixed-order reduction must be reproducible"


def test_argument_errors_are_loud(dev):
    """bad arguments come back as XRError with a message (no silent fallback, no partial result)"""
    from qodeapplications_b200.lib import XRError
    z_lists, _ = inputs()
    n_orbs, n_core = CASE["n_orbs"], CASE["n_core"]
    z0 = dev.upload(z_lists[0].coeffs)
    m0 = dev.upload(numpy.zeros(z_lists[0].configs.shape[0], dtype=numpy.int64), numpy.int64)
    rho = dev.zeros((9 * 64,))
    good = (3, 3, z0, 15, z0, 15, m0, 4, 4, n_orbs, n_core)
    with pytest.raises(XRError, match="must consist of c and a"):
        dev.ctx.density_tensor("cx", rho, *good)
    with pytest.raises(XRError, match="does not connect"):
        dev.ctx.density_tensor("a", rho, *good)
    with pytest.raises(XRError, match="is not C"):
        dev.ctx.density_tensor("ca", rho, 3, 3, z0, 14, z0, 15, m0, 4, 4, n_orbs, n_core)
    with pytest.raises(XRError, match="2\\*n_orbs <= 64"):
        dev.ctx.density_tensor("ca", rho, 3, 3, z0, 15, z0, 15, m0, 4, 4, 40, n_core)
    with pytest.raises(XRError, match="sign must be"):
        dev.ctx.gemm_dd(2, 2, 2, rho, 2, rho, 2, None, 0, 0.5, dev.zeros((4,)), 2)
    with pytest.raises(XRError, match="spectator offset table"):
        t = dev.upload(numpy.zeros(4, dtype=numpy.int64), numpy.int64)
        dev.ctx.embed_add(rho, rho, 2, 2, 2, 3, t, t, None)
    with pytest.raises(XRError, match="do not multiply"):
        dev.ctx.embed_add(rho, rho, 4, 4, 4, 1, t, t, None, dims_sub=[3, 2], min_transitions=2)
