"""catalog of S*T diagrams (hermitian-XRCC/diagrams/ST_diagrams.py:27-40; orders 0-1 built so far)."""
from .build_diagram import build_diagram
from .specs import make_one_fragment, make_two_fragment

t00 = make_one_fragment("t00")
t01, s01t10, s01t00, s01t11, s01t01 = (make_two_fragment(l) for l in ("t01", "s01t10", "s01t00", "s01t11", "s01t01"))

catalog = {}
catalog[1] = {
    "t00": build_diagram(t00, Dchgs=(0,), permutations=[(+1, (0,))]),
}
catalog[2] = {
    "t01":    build_diagram(t01,    Dchgs=(-1, +1), permutations=[(+1, (0, 1)), (-1, (1, 0))]),
    "s01t10": build_diagram(s01t10, Dchgs=(0, 0),   permutations=[(+1, (0, 1)), (+1, (1, 0))]),
    "s01t00": build_diagram(s01t00, Dchgs=(-1, +1), permutations=[(+1, (0, 1)), (-1, (1, 0))]),
    "s01t11": build_diagram(s01t11, Dchgs=(-1, +1), permutations=[(+1, (0, 1)), (-1, (1, 0))]),
    "s01t01": build_diagram(s01t01, Dchgs=(-2, +2), permutations=[(+1, (0, 1)), (+1, (1, 0))]),
}
