"""catalog of S*U diagrams (hermitian-XRCC/diagrams/SU_diagrams.py:27-62; S-orders 0-2 built so far)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment, u100

u000 = make_one_fragment("u000")
_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]

# label suffix (after "u<nucleus>") -> (Dchgs, permutations), identical for both nucleus choices
_rules = {"s01u%s10": ((0, 0), _pp), "s01u%s00": ((-1, +1), _pm), "s01u%s11": ((-1, +1), _pm), "s01u%s01": ((-2, +2), _pp),
          "s01s10u%s00": ((0, 0), _pp), "s01s01u%s10": ((-1, +1), _pm), "s01s10u%s01": ((-1, +1), _pm),
          "s01s01u%s00": ((-2, +2), _pp), "s01s01u%s11": ((-2, +2), _pp)}

catalog = {}
catalog[1] = {
    "u000": build_diagram(u000, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "u100": build_diagram(u100, Dchgs=(0, 0), permutations=_pp),
    "u001": build_diagram(make_two_fragment("u001"), Dchgs=(-1, +1), permutations=_pm),
    "u101": build_diagram(make_two_fragment("u101"), Dchgs=(-1, +1), permutations=_pm),
}
for _pattern, (_Dchgs, _perms) in _rules.items():
    for _n in "01":
        _label = _pattern % _n
        _fn = make_two_fragment(_label)
        globals()[_label] = _fn
        catalog[2][_label] = build_diagram(_fn, Dchgs=_Dchgs, permutations=_perms)
