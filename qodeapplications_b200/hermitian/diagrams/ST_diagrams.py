"""catalog of S*T diagrams (hermitian-XRCC/diagrams/ST_diagrams.py:27-60; every diagram diagram_lists.py activates, orders 0-4)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment

t00 = make_one_fragment("t00")
_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]
RULES = {
    "t01": ((-1, +1), _pm),
    "s01t10": ((0, 0), _pp), "s01t00": ((-1, +1), _pm), "s01t11": ((-1, +1), _pm), "s01t01": ((-2, +2), _pp),
    "s01s10t00": ((0, 0), _pp), "s01s01t10": ((-1, +1), _pm), "s01s10t01": ((-1, +1), _pm), "s01s01t00": ((-2, +2), _pp),
    "s01s01t11": ((-2, +2), _pp),
    "s01s01s10t10": ((0, 0), _pp), "s01s01s10t00": ((-1, +1), _pm), "s01s01s10t11": ((-1, +1), _pm),
    "s01s01s01t10": ((-2, +2), _pp), "s01s01s10t01": ((-2, +2), _pp),
    "s01s01s10s10t00": ((0, 0), _pp), "s01s01s01s10t10": ((-1, +1), _pm), "s01s01s10s10t01": ((-1, +1), _pm),
    "s01s01s01s10t00": ((-2, +2), _pp), "s01s01s01s10t11": ((-2, +2), _pp),
}

catalog = {}
catalog[1] = {
    "t00": build_diagram(t00, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {}
for _label, (_Dchgs, _perms) in RULES.items():
    _fn = make_two_fragment(_label)
    globals()[_label] = _fn
    catalog[2][_label] = build_diagram(_fn, Dchgs=_Dchgs, permutations=_perms)
