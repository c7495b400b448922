"""Which diagrams enter at which order in S (hermitian-XRCC/diagram_lists.py:10-71).  Index = order.
All five orders of the reference (0-4) are listed; asking for more raises."""

class _orders(dict):
    def __missing__(self, order):
        raise NotImplementedError("diagrams of S-order %r are not built in qodeapplications_b200 yet" % (order,))

S0, S2 = _orders(), _orders()
ST1, ST2 = _orders(), _orders()
SU1, SU2 = _orders(), _orders()
SV1, SV2 = _orders(), _orders()

S0[0] = ["identity"]
S2[1] = ["s01"]
S2[2] = ["s01s10", "s01s01"]
S2[3] = ["s01s01s10"]
S2[4] = ["s01s01s10s10", "s01s01s01s10"]

ST1[0] = ["t00"]
ST2[0] = ["t01"]
ST2[1] = ["s01t10", "s01t00", "s01t11", "s01t01"]
ST2[2] = ["s01s10t00", "s01s01t10", "s01s10t01", "s01s01t00", "s01s01t11"]
ST2[3] = ["s01s01s10t10", "s01s01s10t00", "s01s01s10t11", "s01s01s01t10", "s01s01s10t01"]
ST2[4] = ["s01s01s10s10t00", "s01s01s01s10t10", "s01s01s10s10t01", "s01s01s01s10t00", "s01s01s01s10t11"]

SU1[0] = ["u000"]
SU2[0] = ["u100", "u001", "u101"]
SU2[1] = ["s01u010", "s01u000", "s01u011", "s01u001",
          "s01u110", "s01u100", "s01u111", "s01u101"]
SU2[2] = ["s01s10u000", "s01s01u010", "s01s10u001", "s01s01u000", "s01s01u011",
          "s01s10u100", "s01s01u110", "s01s10u101", "s01s01u100", "s01s01u111"]
SU2[3] = ["s01s01s10u010", "s01s01s10u000", "s01s01s10u011", "s01s01s01u010", "s01s01s10u001",
          "s01s01s10u110", "s01s01s10u100", "s01s01s10u111", "s01s01s01u110", "s01s01s10u101"]
SU2[4] = ["s01s01s10s10u000", "s01s01s01s10u010", "s01s01s10s10u001", "s01s01s01s10u000", "s01s01s01s10u011",
          "s01s01s10s10u100", "s01s01s01s10u110", "s01s01s10s10u101", "s01s01s01s10u100", "s01s01s01s10u111"]

SV1[0] = ["v0000"]
SV2[0] = ["v0101", "v0001", "v0100", "v0011"]
SV2[1] = ["s01v0100", "s01v1101", "s01v0000", "s01v0101", "s01v1100", "s01v1111", "s01v0001", "s01v0111"]
SV2[2] = ["s01s01v1100", "s01s10v0000", "s01s10v0101", "s01s01v0100", "s01s01v1101", "s01s10v0001", "s01s10v0100",
          "s01s01v0000", "s01s01v0101", "s01s01v1111", "s01s10v0011"]
SV2[3] = ["s01s01s10v0100", "s01s01s10v1101", "s01s01s01v1100", "s01s01s10v0000", "s01s01s10v0101", "s01s01s10v1100",
          "s01s01s10v1111", "s01s01s01v0100", "s01s01s01v1101", "s01s01s10v0001", "s01s01s10v0111"]
SV2[4] = ["s01s01s01s10v1100", "s01s01s10s10v0000", "s01s01s10s10v0101", "s01s01s01s10v0100", "s01s01s01s10v1101",
          "s01s01s10s10v0001", "s01s01s10s10v0100", "s01s01s01s01v1100", "s01s01s01s10v0000", "s01s01s01s10v0101",
          "s01s01s01s10v1111", "s01s01s10s10v0011"]
