"""Overlap-matrix golden vectors for get_xr_S (hermitian-XRCC/get_xr_result.py:357-422), from the REFERENCE itself
(run through oracle/qode_shim, see gen_golden.py): tests/golden/hermitian_toy_S.npz.   python oracle/gen_golden_S.py

The reference's get_xr_S returns an all-ZERO matrix: its last loop assigns `D.S2[i,j] = H2blocked[i_,j_]` into the diagram
dictionary of diagram_lists.py (:421) instead of into the result (recorded below as `reference_returns_zeros`).  Everything
before that loop is sound, so the pinned quantity is what it computes up to there --
`XR_term.dimer_matrix(S_blocks, {0: D.S0[0], 2: []})` at order 0 and `{0: [], 2: D.S2[1]}` at order 1, charge-blocked --
evaluated by the unmodified reference modules; the final (global i0, global i1) ordering it meant to apply is get_xr_H's
(:300-353).
"""
import contextlib
import io
import os
import sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import gen_golden as gg                       # noqa: E402
from qodeapplications_b200 import synth      # noqa: E402

if __name__ == "__main__":
    system = synth.make_system("toy", ops=synth.OPS_ORDER1, with_bior=True)
    get_xr_result, XR_tensor = gg.import_reference_hermitian()
    import diagrammatic_expansion, XR_term
    import diagram_lists as D
    from diagrams import S_diagrams
    from precontract import precontract
    from qode.util import struct, timer
    init = XR_tensor.init
    dens = []
    for rho in system["densities"][:2]:
        dens.append({k: (v if k in ("n_elec", "n_states", "n_states_bra") else {c: init(t) for c, t in v.items()}) for k, v in rho.items()})
    S = gg._wrap_blocks(system["symm"].S, init)
    charges = system["charges"]
    payload = {"input_sha256": gg.input_checksum(system)}
    with contextlib.redirect_stdout(io.StringIO()):
        S_ref = get_xr_result.get_xr_S((struct(S=S), None, None), dens, 0, [charges, charges])
    for key in [k for k in D.S2 if isinstance(k, tuple)]:      # undo the reference's stray writes into the diagram dictionary
        del D.S2[key]
    payload["reference_returns_zeros"] = bool(numpy.all(numpy.asarray(S_ref) == 0))
    print("reference get_xr_S returns zeros:", payload["reference_returns_zeros"], numpy.asarray(S_ref).shape)
    cache = precontract(dens, S, timer())
    blk = diagrammatic_expansion.blocks(densities=dens, integrals=S, diagrams=S_diagrams, contract_cache=cache, timings=timer(),
                                        precon_timings=timer())
    all_dimer_charges = [(c0, c1) for c0 in charges for c1 in charges]
    for order, active in ((0, {0: D.S0[0], 2: []}), (1, {0: [], 2: D.S2[1]})):
        S2 = numpy.asarray(XR_term.dimer_matrix(blk, active, (0, 1), all_dimer_charges, timer()))
        payload["S2_blocked_order%d" % order] = S2
        print(order, S2.shape, float(numpy.abs(S2).max()))
    numpy.savez_compressed(os.path.join(gg.REPO, "tests", "golden", "hermitian_toy_S.npz"), **payload)
