"""bra_det / ket_det golden vectors from the REFERENCE itself (get_xr_H(..., bra_det=True) / ket_det=True at
xr_order 0, run through oracle/qode_shim): tests/golden/hermitian_det_{bra,ket}.npz.
    python oracle/gen_golden_det.py"""
import contextlib, io, os, sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import gen_golden as gg                       # noqa: E402
from qodeapplications_b200 import synth      # noqa: E402


def reference_det(system, which):
    get_xr_result, XR_tensor = gg.import_reference_hermitian()
    from qode.util import struct
    init = XR_tensor.init
    dens = []
    for rho in system["densities"]:
        wrapped = {}
        for key, val in rho.items():
            wrapped[key] = val if key in ("n_elec", "n_states", "n_states_bra") else {c: init(t) for c, t in val.items()}
        dens.append(wrapped)
    def wrap_set(s):
        out = struct(S=gg._wrap_blocks(s.S, init), T=gg._wrap_blocks(s.T, init), U=gg._wrap_blocks(s.U, init),
                     V=gg._wrap_blocks(s.V, init))
        if hasattr(s, "V_diff"):
            out.V_half, out.V_diff = gg._wrap_blocks(s.V_half, init), gg._wrap_blocks(s.V_diff, init)
        return out
    charges = system["charges"]
    with contextlib.redirect_stdout(io.StringIO()):
        H1, H2 = get_xr_result.get_xr_H((wrap_set(system["symm"]), wrap_set(system["bior"]), system["nuc"]), dens, 0,
                                        [charges, charges], bra_det=(which == "bra"), ket_det=(which == "ket"))
    return [numpy.asarray(h) for h in H1], numpy.asarray(H2)


if __name__ == "__main__":
    out_dir = os.path.join(gg.REPO, "tests", "golden")
    for which in ("bra", "ket"):
        system = synth.make_det_system(which)
        H1, H2 = reference_det(system, which)
        numpy.savez_compressed(os.path.join(out_dir, "hermitian_det_%s.npz" % which), which=which, H1_0=H1[0], H1_1=H1[1], H2=H2)
        print(which, [h.shape for h in H1], H2.shape, float(numpy.abs(H2).max()), int(numpy.count_nonzero(H2)))
