"""catalog of S*U diagrams (hermitian-XRCC/diagrams/SU_diagrams.py:27-50; orders 0-1 built so far)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment, u100

u000 = make_one_fragment("u000")
u001, u101 = make_two_fragment("u001"), make_two_fragment("u101")
_order1 = {l: make_two_fragment(l) for l in ("s01u010", "s01u000", "s01u011", "s01u001", "s01u110", "s01u100", "s01u111", "s01u101")}
globals().update(_order1)

_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]

catalog = {}
catalog[1] = {
    "u000": build_diagram(u000, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "u100": build_diagram(u100, Dchgs=(0, 0),   permutations=_pp),
    "u001": build_diagram(u001, Dchgs=(-1, +1), permutations=_pm),
    "u101": build_diagram(u101, Dchgs=(-1, +1), permutations=_pm),
    "s01u010": build_diagram(_order1["s01u010"], Dchgs=(0, 0),   permutations=_pp),
    "s01u000": build_diagram(_order1["s01u000"], Dchgs=(-1, +1), permutations=_pm),
    "s01u011": build_diagram(_order1["s01u011"], Dchgs=(-1, +1), permutations=_pm),
    "s01u001": build_diagram(_order1["s01u001"], Dchgs=(-2, +2), permutations=_pp),
    "s01u110": build_diagram(_order1["s01u110"], Dchgs=(0, 0),   permutations=_pp),
    "s01u100": build_diagram(_order1["s01u100"], Dchgs=(-1, +1), permutations=_pm),
    "s01u111": build_diagram(_order1["s01u111"], Dchgs=(-1, +1), permutations=_pm),
    "s01u101": build_diagram(_order1["s01u101"], Dchgs=(-2, +2), permutations=_pp),
}
