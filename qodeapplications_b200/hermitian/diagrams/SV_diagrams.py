"""catalog of S*V diagrams (hermitian-XRCC/diagrams/SV_diagrams.py:27-50; orders 0-1 built so far)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment

v0000 = make_one_fragment("v0000")
_two = {l: make_two_fragment(l) for l in ("v0101", "v0001", "v0100", "v0011", "s01v0100", "s01v1101", "s01v0000", "s01v0101",
                                          "s01v1100", "s01v1111", "s01v0001", "s01v0111", "s01v0011")}
globals().update(_two)

_p = [(+1, (0, 1))]
_pm = [(+1, (0, 1)), (-1, (1, 0))]
_pp = [(+1, (0, 1)), (+1, (1, 0))]

catalog = {}
catalog[1] = {
    "v0000": build_diagram(v0000, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "v0101":    build_diagram(_two["v0101"],    Dchgs=(0, 0),   permutations=_p),
    "v0001":    build_diagram(_two["v0001"],    Dchgs=(-1, +1), permutations=_pm),
    "v0100":    build_diagram(_two["v0100"],    Dchgs=(+1, -1), permutations=_pm),
    "v0011":    build_diagram(_two["v0011"],    Dchgs=(-2, +2), permutations=_pp),
    "s01v0100": build_diagram(_two["s01v0100"], Dchgs=(0, 0),   permutations=_pp),
    "s01v1101": build_diagram(_two["s01v1101"], Dchgs=(0, 0),   permutations=_pp),
    "s01v0000": build_diagram(_two["s01v0000"], Dchgs=(-1, +1), permutations=_pm),
    "s01v0101": build_diagram(_two["s01v0101"], Dchgs=(-1, +1), permutations=_pm),
    "s01v1100": build_diagram(_two["s01v1100"], Dchgs=(+1, -1), permutations=_pm),
    "s01v1111": build_diagram(_two["s01v1111"], Dchgs=(-1, +1), permutations=_pm),
    "s01v0001": build_diagram(_two["s01v0001"], Dchgs=(-2, +2), permutations=_pp),
    "s01v0111": build_diagram(_two["s01v0111"], Dchgs=(-2, +2), permutations=_pp),
    "s01v0011": build_diagram(_two["s01v0011"], Dchgs=(-3, +3), permutations=_pm),
}
