"""catalog of S*T diagrams (hermitian-XRCC/diagrams/ST_diagrams.py:27-46; S-orders 0-2 built so far)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment

t00 = make_one_fragment("t00")
_two = {l: make_two_fragment(l) for l in ("t01", "s01t10", "s01t00", "s01t11", "s01t01",
                                          "s01s10t00", "s01s01t10", "s01s10t01", "s01s01t00", "s01s01t11")}
globals().update(_two)

_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]

catalog = {}
catalog[1] = {
    "t00": build_diagram(t00, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "t01":       build_diagram(_two["t01"],       Dchgs=(-1, +1), permutations=_pm),
    "s01t10":    build_diagram(_two["s01t10"],    Dchgs=(0, 0),   permutations=_pp),
    "s01t00":    build_diagram(_two["s01t00"],    Dchgs=(-1, +1), permutations=_pm),
    "s01t11":    build_diagram(_two["s01t11"],    Dchgs=(-1, +1), permutations=_pm),
    "s01t01":    build_diagram(_two["s01t01"],    Dchgs=(-2, +2), permutations=_pp),
    "s01s10t00": build_diagram(_two["s01s10t00"], Dchgs=(0, 0),   permutations=_pp),
    "s01s01t10": build_diagram(_two["s01s01t10"], Dchgs=(-1, +1), permutations=_pm),
    "s01s10t01": build_diagram(_two["s01s10t01"], Dchgs=(-1, +1), permutations=_pm),
    "s01s01t00": build_diagram(_two["s01s01t00"], Dchgs=(-2, +2), permutations=_pp),
    "s01s01t11": build_diagram(_two["s01s01t11"], Dchgs=(-2, +2), permutations=_pp),
}
