"""Third/fourth-order (S^3, S^4) golden vectors from the REFERENCE itself (see gen_golden.py): every charge-resolved
block of the 56 diagrams diagram_lists.py activates at orders 3 and 4 -> tests/golden/hermitian_toy4_blocks34.npz
(system "toy4": n = 5 spin orbitals, densities up to 8 operators).   python oracle/gen_golden_order34.py"""
import os
import sys
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import gen_golden as gg                       # noqa: E402
from qodeapplications_b200 import synth      # noqa: E402
from qodeapplications_b200.hermitian import diagram_lists as D     # label lists only (data)

LABELS34 = {"S": D.S2[3] + D.S2[4], "ST": D.ST2[3] + D.ST2[4], "SU": D.SU2[3] + D.SU2[4], "SV": D.SV2[3] + D.SV2[4]}

if __name__ == "__main__":
    system = synth.make_system("toy4", ops=synth.OPS_ORDER4, with_bior=True)
    blocks = gg.reference_hermitian_blocks(system, LABELS34)
    path = os.path.join(gg.REPO, "tests", "golden", "hermitian_toy4_blocks34.npz")
    numpy.savez_compressed(path, input_sha256=gg.input_checksum(system), **{k: v.astype(numpy.float64) for k, v in blocks.items()})
    print("blocks", len(blocks), "labels", len(set(k.split("|")[0] for k in blocks)), os.path.getsize(path))
