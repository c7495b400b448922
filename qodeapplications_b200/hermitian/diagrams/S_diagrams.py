"""catalog of S diagrams (hermitian-XRCC/diagrams/S_diagrams.py:25-40; every diagram diagram_lists.py activates, orders 0-4)."""
from .build_diagram import build_diagram
from .specs import identity, make_two_fragment

_p = [(+1, (0, 1))]
_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]
_rules = {"s01": ((-1, +1), _pm), "s01s10": ((0, 0), _p), "s01s01": ((-2, +2), _pp),
          "s01s01s10": ((-1, +1), _pm), "s01s01s10s10": ((0, 0), _p), "s01s01s01s10": ((-2, +2), _pp)}

catalog = {}
catalog[0] = {
    "identity": build_diagram(identity, Dchgs=None, permutations=None),
}
catalog[2] = {}
for _label, (_Dchgs, _perms) in _rules.items():
    _fn = make_two_fragment(_label)
    globals()[_label] = _fn
    catalog[2][_label] = build_diagram(_fn, Dchgs=_Dchgs, permutations=_perms)
