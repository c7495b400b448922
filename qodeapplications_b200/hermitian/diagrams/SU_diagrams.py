"""catalog of S*U diagrams (hermitian-XRCC/diagrams/SU_diagrams.py:27-90; every diagram diagram_lists.py activates, orders 0-4).
Apart from u100 they are the S*T diagrams with T replaced by U<nucleus fragment> (nucleus 0 or 1), same charge rules."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment, u100
from .ST_diagrams import RULES as _T_RULES

u000 = make_one_fragment("u000")
_pp = [(+1, (0, 1)), (+1, (1, 0))]

catalog = {}
catalog[1] = {
    "u000": build_diagram(u000, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "u100": build_diagram(u100, Dchgs=(0, 0), permutations=_pp),
}
for _t_label, (_Dchgs, _perms) in _T_RULES.items():
    _k = _t_label.index("t")
    for _n in "01":
        _label = _t_label[:_k] + "u" + _n + _t_label[_k + 1:]
        if _label == "u100":
            continue          # the nucleus-on-the-other-fragment one-electron term is the special delta diagram above
        _fn = make_two_fragment(_label)
        globals()[_label] = _fn
        catalog[2][_label] = build_diagram(_fn, Dchgs=_Dchgs, permutations=_perms)
