"""What does the REFERENCE do with bra_det / ket_det at xr_order 1 and 2?  

hermitian-XRCC/get_xr_result.py:133-296 threads both flags into every XR_term.dimer_matrix call of orders 1 and 2, but the
overlap blocks it inverts are built WITHOUT them (S_blocks, :65), so XR_term._evaluate_block (:62-80) adds four-index
blocks into two-index arrays and numpy raises before any matrix exists; at order 2 the traced vector reaches
numpy.linalg.inv.  This script runs the unmodified reference (through oracle/qode_shim, like every gen_golden*.py) on the
determinant toy systems of gen_golden_det.py and prints, per (variant, order), what it raised -- the pin behind
qodeapplications_b200.hermitian.get_xr_H refusing the same combinations.
    python oracle/check_reference_det_orders.py        -> one JSON line
"""
import contextlib, io, json, os, sys, traceback
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import gen_golden as gg                       # noqa: E402
from qodeapplications_b200 import synth      # noqa: E402


def det_system(which, ops, n_orb=6, seed=23):
    n_states, n_dets = {0: 3, +1: 2, -1: 2}, {0: 4, +1: 3, -1: 3}
    rng = numpy.random.default_rng(seed)
    symm, bior, nuc = synth.make_integrals(2, n_orb, rng, with_bior=True)
    rho0 = synth.make_det_densities(n_orb, n_dets, n_states, rng, ops=ops) if which == "bra" else \
        synth.make_det_densities(n_orb, n_states, n_dets, rng, ops=ops)
    rho1 = synth.make_densities(n_orb, n_states, rng, ops=ops)
    return dict(charges=list(n_states), densities=[rho0, rho1], symm=symm, bior=bior, nuc=nuc)


def main():
    get_xr_result, XR_tensor = gg.import_reference_hermitian()
    from qode.util import struct
    init = XR_tensor.init

    def wrap_set(s):
        out = struct(S=gg._wrap_blocks(s.S, init), T=gg._wrap_blocks(s.T, init), U=gg._wrap_blocks(s.U, init), V=gg._wrap_blocks(s.V, init))
        if hasattr(s, "V_diff"):
            out.V_half, out.V_diff = gg._wrap_blocks(s.V_half, init), gg._wrap_blocks(s.V_diff, init)
        return out

    report = {}
    for which in ("bra", "ket"):
        for order in (0, 1, 2):
            system = det_system(which, {0: synth.OPS_ORDER0, 1: synth.OPS_ORDER1, 2: synth.OPS_ORDER2}[order])
            dens = [{k: (v if k in ("n_elec", "n_states", "n_states_bra") else {c: init(t) for c, t in v.items()}) for k, v in rho.items()}
                    for rho in system["densities"]]
            ch = system["charges"]
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    H1, H2 = get_xr_result.get_xr_H((wrap_set(system["symm"]), wrap_set(system["bior"]), system["nuc"]), dens, order,
                                                    [ch, ch], bra_det=(which == "bra"), ket_det=(which == "ket"))
                report["%s_det order %d" % (which, order)] = {"ok": True, "H2_shape": list(numpy.asarray(H2).shape)}
            except Exception as exc:
                last = traceback.extract_tb(sys.exc_info()[2])[-1]
                report["%s_det order %d" % (which, order)] = {"ok": False, "raises": type(exc).__name__, "message": str(exc)[:160],
                                                               "at": "%s:%d" % (os.path.basename(last.filename), last.lineno)}
    print(json.dumps(report))


if __name__ == "__main__":
    main()
