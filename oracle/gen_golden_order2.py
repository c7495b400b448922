"""Second-order (S^2) golden vectors from the REFERENCE itself (see gen_golden.py for how it is run):
tests/golden/hermitian_toy_order2.npz (get_xr_H at xr_order 2) and hermitian_toy_blocks2.npz (every
charge-resolved block of the 28 active second-order diagrams).   python oracle/gen_golden_order2.py"""
import os
import sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import gen_golden as gg                       # noqa: E402
from qodeapplications_b200 import synth      # noqa: E402

LABELS2 = {"S": ["s01s10", "s01s01"],
           "ST": ["s01s10t00", "s01s01t10", "s01s10t01", "s01s01t00", "s01s01t11"],
           "SU": ["s01s10u000", "s01s01u010", "s01s10u001", "s01s01u000", "s01s01u011",
                  "s01s10u100", "s01s01u110", "s01s10u101", "s01s01u100", "s01s01u111"],
           "SV": ["s01s01v1100", "s01s10v0000", "s01s10v0101", "s01s01v0100", "s01s01v1101", "s01s10v0001", "s01s10v0100",
                  "s01s01v0000", "s01s01v0101", "s01s01v1111", "s01s10v0011"]}

if __name__ == "__main__":
    out_dir = os.path.join(gg.REPO, "tests", "golden")
    system = synth.make_system("toy", ops=synth.OPS_ORDER2, with_bior=True)
    H1, H2 = gg.reference_hermitian(system, 2)
    numpy.savez_compressed(os.path.join(out_dir, "hermitian_toy_order2.npz"), config="toy", seed=synth.SEEDS["toy"], xr_order=2,
                           ops=",".join(synth.OPS_ORDER2), input_sha256=gg.input_checksum(system), H1_0=H1[0], H1_1=H1[1], H2=H2)
    print("order 2 H2", H2.shape, float(numpy.abs(H2).max()))
    blocks = gg.reference_hermitian_blocks(system, LABELS2)
    numpy.savez_compressed(os.path.join(out_dir, "hermitian_toy_blocks2.npz"), input_sha256=gg.input_checksum(system), **blocks)
    print("blocks", len(blocks))
