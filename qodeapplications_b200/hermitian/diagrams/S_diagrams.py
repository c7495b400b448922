"""catalog of S diagrams (hermitian-XRCC/diagrams/S_diagrams.py:25-40; S-orders 0-2 built so far)."""
from .build_diagram import build_diagram
from .specs import identity, make_two_fragment

s01, s01s10, s01s01 = (make_two_fragment(l) for l in ("s01", "s01s10", "s01s01"))

catalog = {}
catalog[0] = {
    "identity": build_diagram(identity, Dchgs=None, permutations=None),
}
catalog[2] = {
    "s01":    build_diagram(s01,    Dchgs=(-1, +1), permutations=[(+1, (0, 1)), (-1, (1, 0))]),
    "s01s10": build_diagram(s01s10, Dchgs=(0, 0),   permutations=[(+1, (0, 1))]),
    "s01s01": build_diagram(s01s01, Dchgs=(-2, +2), permutations=[(+1, (0, 1)), (+1, (1, 0))]),
}
