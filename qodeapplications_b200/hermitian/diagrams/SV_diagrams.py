"""catalog of S*V diagrams (hermitian-XRCC/diagrams/SV_diagrams.py:27-90; every diagram diagram_lists.py activates, orders 0-4)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment

v0000 = make_one_fragment("v0000")
_p = [(+1, (0, 1))]
_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]

_rules = {
    "v0101": ((0, 0), _p), "v0001": ((-1, +1), _pm), "v0100": ((+1, -1), _pm), "v0011": ((-2, +2), _pp),
    "s01v0100": ((0, 0), _pp), "s01v1101": ((0, 0), _pp), "s01v0000": ((-1, +1), _pm), "s01v0101": ((-1, +1), _pm),
    "s01v1100": ((+1, -1), _pm), "s01v1111": ((-1, +1), _pm), "s01v0001": ((-2, +2), _pp), "s01v0111": ((-2, +2), _pp),
    "s01v0011": ((-3, +3), _pm),
    "s01s01v1100": ((0, 0), _pp), "s01s10v0000": ((0, 0), _pp), "s01s10v0101": ((0, 0), _p),
    "s01s01v0100": ((-1, +1), _pm), "s01s01v1101": ((-1, +1), _pm), "s01s10v0001": ((-1, +1), _pm),
    "s01s10v0100": ((+1, -1), _pm), "s01s01v0000": ((-2, +2), _pp), "s01s01v0101": ((-2, +2), _pp),
    "s01s01v1111": ((-2, +2), _pp), "s01s10v0011": ((-2, +2), _pp),
    "s01s01s10v0100": ((0, 0), _pp), "s01s01s10v1101": ((0, 0), _pp), "s01s01s01v1100": ((-1, +1), _pm),
    "s01s01s10v0000": ((-1, +1), _pm), "s01s01s10v0101": ((-1, +1), _pm), "s01s01s10v1100": ((+1, -1), _pm),
    "s01s01s10v1111": ((-1, +1), _pm), "s01s01s01v0100": ((-2, +2), _pp), "s01s01s01v1101": ((-2, +2), _pp),
    "s01s01s10v0001": ((-2, +2), _pp), "s01s01s10v0111": ((-2, +2), _pp),
    "s01s01s01s10v1100": ((0, 0), _pp), "s01s01s10s10v0000": ((0, 0), _pp), "s01s01s10s10v0101": ((0, 0), _p),
    "s01s01s01s10v0100": ((-1, +1), _pm), "s01s01s01s10v1101": ((-1, +1), _pm), "s01s01s10s10v0001": ((-1, +1), _pm),
    "s01s01s10s10v0100": ((+1, -1), _pm), "s01s01s01s01v1100": ((-2, +2), _pp), "s01s01s01s10v0000": ((-2, +2), _pp),
    "s01s01s01s10v0101": ((-2, +2), _pp), "s01s01s01s10v1111": ((-2, +2), _pp), "s01s01s10s10v0011": ((-2, +2), _pp),
}

catalog = {}
catalog[1] = {
    "v0000": build_diagram(v0000, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {}
for _label, (_Dchgs, _perms) in _rules.items():
    _fn = make_two_fragment(_label)
    globals()[_label] = _fn
    catalog[2][_label] = build_diagram(_fn, Dchgs=_Dchgs, permutations=_perms)
