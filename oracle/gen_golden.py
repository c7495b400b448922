"""Generate tests/golden/*.npz by running the REFERENCE ITSELF in this container.

TEST INFRASTRUCTURE.  Runs only where /root/reference exists (this container, not the GPU
box).  With oracle/qode_shim on sys.path the reference's own files import unmodified:

* general-XRCC/build_H.py (+ its H_contractions.c, compiled where it lies by the shim's
  import_C) evaluated element by element in the loop order of general-XRCC/test_H.py:65-142;
* hermitian-XRCC/get_xr_result.get_xr_H (XR_term, diagrammatic_expansion, precontract,
  diagrams/*) at xr_order 0, 1, 2.

Inputs come from qodeapplications_b200.synth (seeded); each fixture stores the config, the
seed, a checksum of the inputs and the reference's outputs.

    python oracle/gen_golden.py            # writes tests/golden/
"""
import hashlib
import itertools
import io
import contextlib
import os
import sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REFERENCE = os.environ.get("XR_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from qodeapplications_b200 import synth      # noqa: E402


def _reference_paths(kind):
    shim = os.path.join(HERE, "qode_shim")
    if kind == "general":
        return [shim, os.path.join(REFERENCE, "general-XRCC")]
    return [shim, os.path.join(REFERENCE, "hermitian-XRCC")]


def _purge(names):
    for name in list(sys.modules):
        if name in names or name.split(".")[0] in names:
            del sys.modules[name]


def import_reference_general():
    _purge({"build_H"})
    paths = _reference_paths("general")
    sys.path[:0] = paths
    try:
        import build_H
    finally:
        for p in paths:
            sys.path.remove(p)
    return build_H


def import_reference_hermitian():
    _purge({"get_xr_result", "XR_term", "diagrammatic_expansion", "precontract", "diagrams", "diagram_lists",
            "XR_tensor", "excitonic"})
    paths = _reference_paths("hermitian")
    sys.path[:0] = paths
    try:
        import get_xr_result
        import XR_tensor
    finally:
        for p in paths:
            sys.path.remove(p)
    return get_xr_result, XR_tensor


class _obj(object):
    pass


def input_checksum(system):
    h = hashlib.sha256()
    for m, rho in enumerate(system["densities"]):
        for op in sorted(k for k in rho if k not in ("n_elec", "n_states", "n_states_bra", "KetCoeffs")):
            for key in sorted(rho[op]):
                h.update(numpy.ascontiguousarray(rho[op][key]).tobytes())
    F = system["n_frag"]
    for m1 in range(F):
        for m2 in range(F):
            h.update(system["symm"].T[m1, m2].tobytes())
            h.update(system["symm"].V[m1, m2, m1, m2].tobytes())
    return h.hexdigest()


# ----------------------------------------------------------------------------- general-XRCC

def reference_general(system, trimers=True):
    """H1[m], H2[(m1,m2)], H3[(m1,m2,m3)] exactly as general-XRCC/test_H.py would fill them
    (dense, state order = fragment.state_indices, product basis row-major), computed by the
    reference's own build_matrix_elements, one element per call."""
    build_H = import_reference_general()
    frags = system["fragments"]
    integrals = _obj()
    integrals.T, integrals.U, integrals.V = system["symm"].T, system["symm"].U, system["symm"].V
    compute = build_H.build_matrix_elements(frags, integrals, system["nuc"])
    F = len(frags)
    H1, H2, H3 = {}, {}, {}
    for M in range(F):
        st = frags[M].state_indices
        H1[M] = numpy.array([[compute.monomer(M, I, J) for J in st] for I in st], dtype=float)
    for M in range(F):
        for N in range(M + 1, F):
            basis = [(iM, iN) for iM in frags[M].state_indices for iN in frags[N].state_indices]
            mat = numpy.zeros((len(basis), len(basis)))
            for i, I in enumerate(basis):
                for j, J in enumerate(basis):
                    val = compute.dimer((M, N), I, J)
                    mat[i, j] = 0. if val is None else val
            H2[(M, N)] = mat
    if trimers:
        for M in range(F):
            for N in range(M + 1, F):
                for O in range(N + 1, F):
                    basis = [(iM, iN, iO) for iM in frags[M].state_indices for iN in frags[N].state_indices
                             for iO in frags[O].state_indices]
                    mat = numpy.zeros((len(basis), len(basis)))
                    for i, I in enumerate(basis):
                        for j, J in enumerate(basis):
                            val = compute.trimer((M, N, O), I, J)
                            mat[i, j] = 0. if val is None else val
                    H3[(M, N, O)] = mat
    return H1, H2, H3


# --------------------------------------------------------------------------- hermitian-XRCC

def _wrap_blocks(blocked, init):
    class wrapped(object):
        def __init__(self):
            self._cache = {}
        def __getitem__(self, key):
            if key not in self._cache:
                self._cache[key] = init(blocked[key])
            return self._cache[key]
    return wrapped()


def reference_hermitian(system, xr_order, frags=(0, 1)):
    get_xr_result, XR_tensor = import_reference_hermitian()
    from qode.util import struct
    init = XR_tensor.init
    dens = []
    for m in frags:
        rho = system["densities"][m]
        wrapped = {}
        for key, val in rho.items():
            if key in ("n_elec", "n_states", "n_states_bra"):
                wrapped[key] = val
            else:
                wrapped[key] = {chgs: init(t) for chgs, t in val.items()}
        dens.append(wrapped)
    def wrap_set(s):
        out = struct(S=_wrap_blocks(s.S, init), T=_wrap_blocks(s.T, init), U=_wrap_blocks(s.U, init),
                     V=_wrap_blocks(s.V, init))
        if hasattr(s, "V_half"):
            out.V_half = _wrap_blocks(s.V_half, init)
            out.V_diff = _wrap_blocks(s.V_diff, init)
        return out
    symm, bior = wrap_set(system["symm"]), wrap_set(system["bior"])
    charges = system["charges"]
    with contextlib.redirect_stdout(io.StringIO()):
        H1, H2 = get_xr_result.get_xr_H((symm, bior, system["nuc"]), dens, xr_order, [charges, charges])
    return [numpy.asarray(h) for h in H1], numpy.asarray(H2)


def reference_hermitian_blocks(system, labels_by_family):
    """Every 2-fragment diagram block blocks[(0,1)][charges][label] of the reference, for every charge
    combination, plus the 1-fragment blocks -- straight from the reference's diagrammatic_expansion."""
    get_xr_result, XR_tensor = import_reference_hermitian()
    import diagrammatic_expansion
    from diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams
    from precontract import precontract
    from qode.util import struct, timer
    init = XR_tensor.init
    dens = []
    for rho in system["densities"][:2]:
        wrapped = {}
        for key, val in rho.items():
            wrapped[key] = val if key in ("n_elec", "n_states", "n_states_bra") else {c: init(t) for c, t in val.items()}
        dens.append(wrapped)
    symm = system["symm"]
    S = _wrap_blocks(symm.S, init)
    cache = precontract(dens, S, timer())
    families = {
        "S": (S, S_diagrams),
        "ST": (struct(S=S, T=_wrap_blocks(symm.T, init)), ST_diagrams),
        "SU": (struct(S=S, U=_wrap_blocks(symm.U, init)), SU_diagrams),
        "SV": (struct(S=S, V=_wrap_blocks(symm.V, init)), SV_diagrams),
    }
    charges = system["charges"]
    out = {}
    for fam, labels in labels_by_family.items():
        ints, diagrams = families[fam]
        blk = diagrammatic_expansion.blocks(densities=dens, integrals=ints, diagrams=diagrams, contract_cache=cache,
                                            timings=timer(), precon_timings=timer())
        for label in labels:
            if label in diagrams.catalog.get(1, {}):
                for m in (0, 1):
                    for c in charges:
                        out["%s|%d|%d,%d" % (label, m, c, c)] = numpy.asarray(blk[(m,)][((c, c),)][label])
                continue
            for ci0, ci1, cj0, cj1 in itertools.product(charges, repeat=4):
                if ci0 + ci1 != cj0 + cj1:
                    continue
                val = blk[(0, 1)][((ci0, cj0), (ci1, cj1))][label]
                if val is not None:
                    out["%s|%d,%d,%d,%d" % (label, ci0, ci1, cj0, cj1)] = numpy.asarray(val)
    return out


# ------------------------------------------------------------------------------------- main

def main():
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # general-XRCC: 2-fragment toy (all dimer branches) and 3-fragment toy (all 12 trimer branches)
    for name in ("toy", "toy3", "toy5", "toyh", "toyh3"):
        system = synth.make_system(name)
        H1, H2, H3 = reference_general(system, trimers=True)
        payload = {"config": name, "seed": synth.SEEDS[name], "input_sha256": input_checksum(system)}
        for m, h in H1.items():
            payload["H1_%d" % m] = h
        for (m1, m2), h in H2.items():
            payload["H2_%d%d" % (m1, m2)] = h
        for (m1, m2, m3), h in H3.items():
            # keep the fixture small: non-zeros only (index + value); shape recorded
            nz = numpy.nonzero(h)
            payload["H3_%d%d%d_shape" % (m1, m2, m3)] = numpy.array(h.shape)
            payload["H3_%d%d%d_rows" % (m1, m2, m3)] = nz[0].astype(numpy.int32)
            payload["H3_%d%d%d_cols" % (m1, m2, m3)] = nz[1].astype(numpy.int32)
            payload["H3_%d%d%d_vals" % (m1, m2, m3)] = h[nz]
        path = os.path.join(out_dir, "general_%s.npz" % name)
        numpy.savez_compressed(path, **payload)
        print("wrote", path, {k: getattr(v, "shape", v) for k, v in payload.items() if k.startswith("H2")})

    # hermitian-XRCC: toy dimer at xr_order 0, 1, 2 (orders 1, 2 need the rank-4/5/6 densities)
    for order, ops in ((0, synth.OPS_ORDER0), (1, synth.OPS_ORDER1)):
        system = synth.make_system("toy", ops=ops, with_bior=True)
        H1, H2 = reference_hermitian(system, order)
        path = os.path.join(out_dir, "hermitian_toy_order%d.npz" % order)
        numpy.savez_compressed(path, config="toy", seed=synth.SEEDS["toy"], xr_order=order, ops=",".join(ops),
                               input_sha256=input_checksum(system), H1_0=H1[0], H1_1=H1[1], H2=H2)
        print("wrote", path, H2.shape, float(numpy.abs(H2).max()))


def main_blocks():
    out_dir = os.path.join(REPO, "tests", "golden")
    system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    labels = {"S": ["s01"], "ST": ["t00", "t01", "s01t10", "s01t00", "s01t11", "s01t01"],
              "SU": ["u000", "u100", "u001", "u101", "s01u010", "s01u000", "s01u011", "s01u001", "s01u110", "s01u100",
                     "s01u111", "s01u101"],
              "SV": ["v0000", "v0101", "v0001", "v0100", "v0011", "s01v0100", "s01v1101", "s01v0000", "s01v0101", "s01v1100",
                     "s01v1111", "s01v0001", "s01v0111"]}
    blocks = reference_hermitian_blocks(system, labels)
    path = os.path.join(out_dir, "hermitian_toy_blocks.npz")
    numpy.savez_compressed(path, input_sha256=input_checksum(system), **blocks)
    print("wrote", path, len(blocks), "blocks")


if __name__ == "__main__":
    if "--blocks" in sys.argv:
        main_blocks()
        sys.exit(0)
    main()
