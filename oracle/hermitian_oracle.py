"""CPU oracle for the hermitian-XRCC Hamiltonian build (diagram blocks of S-orders 0-4, get_xr_H at xr_order 0-2).  TEST INFRASTRUCTURE: only
tests/, smoke() and bench.py's CPU legs may import this.

Third-party dependency of the reference on this path: ``qode`` (github adutoi/Qode, version
unpinned -- no requirements or lock file, README.rst:28-33) plus tensorly/opt_einsum, none of
them under /root/reference or installed.  The arithmetic below that seam is plain tensor contraction,
so this file restates it with numpy.einsum from the UN-precontracted definition that every reference
diagram function carries as a comment (e.g. hermitian-XRCC/diagrams/SV_2mer_0.py:29-31), plus

  * build_diagram.py:87-97   -- n_j0 parity of the unpermuted first fragment's ket electrons, permutations
  * diagrammatic_expansion.py:27-59 -- transposition back by [perm] + [n + perm], sum over permutations
  * XR_term.py:22-165        -- packing into charge-blocked matrices, Kronecker deltas over spectators
  * get_xr_result.py:86-213, 300-353 -- which families enter at xr_order 0 / 1, S2inv, final reorder.

Parity status: PINNED against the reference itself: tests/golden/hermitian_toy_order{0,1}.npz (H1, H2 of
the reference's own get_xr_H, run through oracle/qode_shim by oracle/gen_golden.py; also order 2) and
tests/golden/hermitian_toy_blocks{,2}.npz (every diagram block of orders 0-2 for every charge combination).
"""
import itertools
import numpy

# label: (coefficient, parity shift or None, [(tensor name, index letters), ...]);  i,j = bra,ket of diagram
# fragment 0; k,l = bra,ket of diagram fragment 1.  Tensor names: <op><frag> density, s/t<ff>, u<n>_<ff>, v<ffff>.
TWO_FRAGMENT = {
    "t01":      (1, 0,     [("c0", "ijp"), ("a1", "klq"), ("t01", "pq")]),                       # ST_2mer_0.py:29-31
    "u001":     (1, 0,     [("c0", "ijp"), ("a1", "klq"), ("u0_01", "pq")]),                     # SU_2mer_0.py:102-104
    "u101":     (1, 0,     [("c0", "ijp"), ("a1", "klq"), ("u1_01", "pq")]),                     # SU_2mer_0.py:113-115
    "v0101":    (4, None,  [("ca0", "ijpr"), ("ca1", "klqs"), ("v0101", "pqrs")]),               # SV_2mer_0.py:29-31
    "v0001":    (2, 1,     [("cca0", "ijpqr"), ("a1", "kls"), ("v0001", "pqrs")]),               # SV_2mer_0.py:40-42
    "v0100":    (2, 0,     [("caa0", "ijpsr"), ("c1", "klq"), ("v0100", "pqrs")]),               # SV_2mer_0.py:51-53
    "v0011":    (1, None,  [("cc0", "ijpq"), ("aa1", "klsr"), ("v0011", "pqrs")]),               # SV_2mer_0.py:62-64
    "s01":      (1, 0,     [("c0", "ijp"), ("a1", "klq"), ("s01", "pq")]),                       # S_2mer_1.py:29-31
    "s01t10":   (-1, None, [("ca0", "ijtq"), ("ca1", "klpu"), ("s01", "tu"), ("t10", "pq")]),    # ST_2mer_1.py:29-32
    "s01t00":   (1, 1,     [("cca0", "ijptq"), ("a1", "klu"), ("s01", "tu"), ("t00", "pq")]),    # ST_2mer_1.py:42-45
    "s01t11":   (1, 1,     [("c0", "ijt"), ("caa1", "klpuq"), ("s01", "tu"), ("t11", "pq")]),    # ST_2mer_1.py:55-58
    "s01t01":   (1, None,  [("cc0", "ijpt"), ("aa1", "kluq"), ("s01", "tu"), ("t01", "pq")]),    # ST_2mer_1.py:68-71
    "s01v0100": (-2, None, [("ccaa0", "ijptsr"), ("ca1", "klqu"), ("s01", "tu"), ("v0100", "pqrs")]),     # SV_2mer_1.py:29-32
    "s01v1101": (2, None,  [("ca0", "ijtr"), ("ccaa1", "klpqus"), ("s01", "tu"), ("v1101", "pqrs")]),     # SV_2mer_1.py:42-45
    "s01v0000": (1, 0,     [("cccaa0", "ijpqtsr"), ("a1", "klu"), ("s01", "tu"), ("v0000", "pqrs")]),     # SV_2mer_1.py:55-58
    "s01v0101": (4, 0,     [("cca0", "ijptr"), ("caa1", "klqus"), ("s01", "tu"), ("v0101", "pqrs")]),     # SV_2mer_1.py:68-71
    "s01v1100": (1, 0,     [("caa0", "ijtsr"), ("cca1", "klpqu"), ("s01", "tu"), ("v1100", "pqrs")]),     # SV_2mer_1.py:81-84
    "s01v1111": (1, 0,     [("c0", "ijt"), ("ccaaa1", "klpqusr"), ("s01", "tu"), ("v1111", "pqrs")]),     # SV_2mer_1.py:94-97
    "s01v0001": (2, None,  [("ccca0", "ijpqtr"), ("aa1", "klus"), ("s01", "tu"), ("v0001", "pqrs")]),     # SV_2mer_1.py:107-110
    "s01v0111": (-2, None, [("cc0", "ijpt"), ("caaa1", "klqusr"), ("s01", "tu"), ("v0111", "pqrs")]),     # SV_2mer_1.py:120-123
}
for _n in "01":      # SU_2mer_1.py: the ST forms with t## -> u<n>_##
    TWO_FRAGMENT["s01u%s10" % _n] = (-1, None, [("ca0", "ijtq"), ("ca1", "klpu"), ("s01", "tu"), ("u%s_10" % _n, "pq")])
    TWO_FRAGMENT["s01u%s00" % _n] = (1, 1, [("cca0", "ijptq"), ("a1", "klu"), ("s01", "tu"), ("u%s_00" % _n, "pq")])
    TWO_FRAGMENT["s01u%s11" % _n] = (1, 1, [("c0", "ijt"), ("caa1", "klpuq"), ("s01", "tu"), ("u%s_11" % _n, "pq")])
    TWO_FRAGMENT["s01u%s01" % _n] = (1, None, [("cc0", "ijpt"), ("aa1", "kluq"), ("s01", "tu"), ("u%s_01" % _n, "pq")])

# ---- S-order 2 (S_2mer_2.py, ST_2mer_2.py, SU_2mer_2.py, SV_2mer_2.py: the commented un-precontracted forms) ----
_H = 0.5
TWO_FRAGMENT.update({
    "s01s10":    (-1, None, [("ca0", "ijps"), ("ca1", "klrq"), ("s01", "pq"), ("s10", "rs")]),
    "s01s01":    (_H, None, [("cc0", "ijpr"), ("aa1", "klsq"), ("s01", "pq"), ("s01", "rs")]),
    "s01s10t00": (-1, None, [("ccaa0", "ijptwq"), ("ca1", "klvu"), ("s01", "tu"), ("s10", "vw"), ("t00", "pq")]),
    "s01s01t10": (_H, 1,    [("cca0", "ijtvq"), ("caa1", "klpwu"), ("s01", "tu"), ("s01", "vw"), ("t10", "pq")]),
    "s01s10t01": (1, 1,     [("cca0", "ijptw"), ("caa1", "klvuq"), ("s01", "tu"), ("s10", "vw"), ("t01", "pq")]),
    "s01s01t00": (_H, None, [("ccca0", "ijptvq"), ("aa1", "klwu"), ("s01", "tu"), ("s01", "vw"), ("t00", "pq")]),
    "s01s01t11": (_H, None, [("cc0", "ijtv"), ("caaa1", "klpwuq"), ("s01", "tu"), ("s01", "vw"), ("t11", "pq")]),
    "s01s01v1100": (_H, None, [("ccaa0", "ijtvsr"), ("ccaa1", "klpqwu"), ("s01", "tu"), ("s01", "vw"), ("v1100", "pqrs")]),
    "s01s10v0000": (-1, None, [("cccaaa0", "ijpqtwsr"), ("ca1", "klvu"), ("s01", "tu"), ("s10", "vw"), ("v0000", "pqrs")]),
    "s01s10v0101": (-4, None, [("ccaa0", "ijptwr"), ("ccaa1", "klqvus"), ("s01", "tu"), ("s10", "vw"), ("v0101", "pqrs")]),
    "s01s01v0100": (1, 0,     [("cccaa0", "ijptvsr"), ("caa1", "klqwu"), ("s01", "tu"), ("s01", "vw"), ("v0100", "pqrs")]),
    "s01s01v1101": (1, 1,     [("cca0", "ijtvr"), ("ccaaa1", "klpqwus"), ("s01", "tu"), ("s01", "vw"), ("v1101", "pqrs")]),
    "s01s10v0001": (2, 0,     [("cccaa0", "ijpqtwr"), ("caa1", "klvus"), ("s01", "tu"), ("s10", "vw"), ("v0001", "pqrs")]),
    "s01s10v0100": (2, 1,     [("ccaaa0", "ijptwsr"), ("cca1", "klqvu"), ("s01", "tu"), ("s10", "vw"), ("v0100", "pqrs")]),
    "s01s01v0000": (_H, None, [("ccccaa0", "ijpqtvsr"), ("aa1", "klwu"), ("s01", "tu"), ("s01", "vw"), ("v0000", "pqrs")]),
    "s01s01v0101": (2, None,  [("ccca0", "ijptvr"), ("caaa1", "klqwus"), ("s01", "tu"), ("s01", "vw"), ("v0101", "pqrs")]),
    "s01s01v1111": (_H, None, [("cc0", "ijtv"), ("ccaaaa1", "klpqwusr"), ("s01", "tu"), ("s01", "vw"), ("v1111", "pqrs")]),
    "s01s10v0011": (-1, None, [("ccca0", "ijpqtw"), ("caaa1", "klvusr"), ("s01", "tu"), ("s10", "vw"), ("v0011", "pqrs")]),
})
# ---- S-orders 3 and 4 (S_2mer_{3,4}.py, ST_2mer_{3,4}.py, SV_2mer_{3,4}.py), the diagrams diagram_lists.py activates ----
TWO_FRAGMENT.update({
    "s01s01s10":         (1 / 2, 1, [("cca0", "ijpru"), ("caa1", "kltsq"), ("s01", "pq"), ("s01", "rs"), ("s10", "tu")]),
    "s01s01s10s10":      (1 / 4, None, [("ccaa0", "ijprwu"), ("ccaa1", "kltvsq"), ("s01", "pq"), ("s01", "rs"), ("s10", "tu"), ("s10", "vw")]),
    "s01s01s01s10":      (-1 / 6, None, [("ccca0", "ijprtw"), ("caaa1", "klvusq"), ("s01", "pq"), ("s01", "rs"), ("s01", "tu"), ("s10", "vw")]),
    "s01s01s10t10":      (1 / 2, None, [("ccaa0", "ijtvyq"), ("ccaa1", "klpxwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("t10", "pq")]),
    "s01s01s10t00":      (1 / 2, 0, [("cccaa0", "ijptvyq"), ("caa1", "klxwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("t00", "pq")]),
    "s01s01s10t11":      (1 / 2, 0, [("cca0", "ijtvy"), ("ccaaa1", "klpxwuq"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("t11", "pq")]),
    "s01s01s01t10":      (-1 / 6, None, [("ccca0", "ijtvxq"), ("caaa1", "klpywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("t10", "pq")]),
    "s01s01s10t01":      (-1 / 2, None, [("ccca0", "ijptvy"), ("caaa1", "klxwuq"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("t01", "pq")]),
    "s01s01s10s10t00":   (1 / 4, None, [("cccaaa0", "ijptvayq"), ("ccaa1", "klxzwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("t00", "pq")]),
    "s01s01s01s10t10":   (1 / 6, 0, [("cccaa0", "ijtvxaq"), ("ccaaa1", "klpzywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("t10", "pq")]),
    "s01s01s10s10t01":   (1 / 4, 0, [("cccaa0", "ijptvay"), ("ccaaa1", "klxzwuq"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("t01", "pq")]),
    "s01s01s01s10t00":   (-1 / 6, None, [("ccccaa0", "ijptvxaq"), ("caaa1", "klzywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("t00", "pq")]),
    "s01s01s01s10t11":   (-1 / 6, None, [("ccca0", "ijtvxa"), ("ccaaaa1", "klpzywuq"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("t11", "pq")]),
    "s01s01s10v0100":    (1, None, [("cccaaa0", "ijptvysr"), ("ccaa1", "klqxwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v0100", "pqrs")]),
    "s01s01s10v1101":    (-1, None, [("ccaa0", "ijtvyr"), ("cccaaa1", "klpqxwus"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v1101", "pqrs")]),
    "s01s01s01v1100":    (1 / 6, 0, [("cccaa0", "ijtvxsr"), ("ccaaa1", "klpqywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("v1100", "pqrs")]),
    "s01s01s10v0000":    (1 / 2, 1, [("ccccaaa0", "ijpqtvysr"), ("caa1", "klxwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v0000", "pqrs")]),
    "s01s01s10v0101":    (2, 1, [("cccaa0", "ijptvyr"), ("ccaaa1", "klqxwus"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v0101", "pqrs")]),
    "s01s01s10v1100":    (1 / 2, 1, [("ccaaa0", "ijtvysr"), ("cccaa1", "klpqxwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v1100", "pqrs")]),
    "s01s01s10v1111":    (1 / 2, 1, [("cca0", "ijtvy"), ("cccaaaa1", "klpqxwusr"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v1111", "pqrs")]),
    "s01s01s01v0100":    (-1 / 3, None, [("ccccaa0", "ijptvxsr"), ("caaa1", "klqywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("v0100", "pqrs")]),
    "s01s01s01v1101":    (1 / 3, None, [("ccca0", "ijtvxr"), ("ccaaaa1", "klpqywus"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("v1101", "pqrs")]),
    "s01s01s10v0001":    (-1, None, [("ccccaa0", "ijpqtvyr"), ("caaa1", "klxwus"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v0001", "pqrs")]),
    "s01s01s10v0111":    (1, None, [("ccca0", "ijptvy"), ("ccaaaa1", "klqxwusr"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("v0111", "pqrs")]),
    "s01s01s01s10v1100": (-1 / 6, None, [("cccaaa0", "ijtvxasr"), ("cccaaa1", "klpqzywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v1100", "pqrs")]),
    "s01s01s10s10v0000": (1 / 4, None, [("ccccaaaa0", "ijpqtvaysr"), ("ccaa1", "klxzwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("v0000", "pqrs")]),
    "s01s01s10s10v0101": (1, None, [("cccaaa0", "ijptvayr"), ("cccaaa1", "klqxzwus"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("v0101", "pqrs")]),
    "s01s01s01s10v0100": (1 / 3, 1, [("ccccaaa0", "ijptvxasr"), ("ccaaa1", "klqzywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v0100", "pqrs")]),
    "s01s01s01s10v1101": (1 / 3, 0, [("cccaa0", "ijtvxar"), ("cccaaaa1", "klpqzywus"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v1101", "pqrs")]),
    "s01s01s10s10v0001": (1 / 2, 1, [("ccccaaa0", "ijpqtvayr"), ("ccaaa1", "klxzwus"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("v0001", "pqrs")]),
    "s01s01s10s10v0100": (1 / 2, 0, [("cccaaaa0", "ijptvaysr"), ("cccaa1", "klqxzwu"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("v0100", "pqrs")]),
    "s01s01s01s01v1100": (1 / 24, None, [("ccccaa0", "ijtvxzsr"), ("ccaaaa1", "klpqaywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s01", "za"), ("v1100", "pqrs")]),
    "s01s01s01s10v0000": (-1 / 6, None, [("cccccaaa0", "ijpqtvxasr"), ("caaa1", "klzywu"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v0000", "pqrs")]),
    "s01s01s01s10v0101": (-2 / 3, None, [("ccccaa0", "ijptvxar"), ("ccaaaa1", "klqzywus"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v0101", "pqrs")]),
    "s01s01s01s10v1111": (-1 / 6, None, [("ccca0", "ijtvxa"), ("cccaaaaa1", "klpqzywusr"), ("s01", "tu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za"), ("v1111", "pqrs")]),
    "s01s01s10s10v0011": (1 / 4, None, [("ccccaa0", "ijpqtvay"), ("ccaaaa1", "klxzwusr"), ("s01", "tu"), ("s01", "vw"), ("s10", "xy"), ("s10", "za"), ("v0011", "pqrs")]),
})
for _n in "01":      # SU_2mer_{2,3,4}.py: the ST forms with t## -> u<n>_##
    for _t, _u in (("s01s10t00", "s01s10u%s00"), ("s01s01t10", "s01s01u%s10"), ("s01s10t01", "s01s10u%s01"),
                   ("s01s01t00", "s01s01u%s00"), ("s01s01t11", "s01s01u%s11"),
                   ("s01s01s10t10", "s01s01s10u%s10"), ("s01s01s10t00", "s01s01s10u%s00"), ("s01s01s10t11", "s01s01s10u%s11"),
                   ("s01s01s01t10", "s01s01s01u%s10"), ("s01s01s10t01", "s01s01s10u%s01"),
                   ("s01s01s10s10t00", "s01s01s10s10u%s00"), ("s01s01s01s10t10", "s01s01s01s10u%s10"),
                   ("s01s01s10s10t01", "s01s01s10s10u%s01"), ("s01s01s01s10t00", "s01s01s01s10u%s00"),
                   ("s01s01s01s10t11", "s01s01s01s10u%s11")):
        _c, _s, _terms = TWO_FRAGMENT[_t]
        TWO_FRAGMENT[_u % _n] = (_c, _s, [(("u%s_%s" % (_n, _nm[1:])) if _nm[0] == "t" else _nm, _ix) for _nm, _ix in _terms])

ONE_FRAGMENT = {
    "t00":   [("ca0", "ijpq"), ("t00", "pq")],              # ST_1mer_0.py:24-30
    "u000":  [("ca0", "ijpq"), ("u0_00", "pq")],            # SU_1mer_0.py:24-30
    "v0000": [("ccaa0", "ijpqsr"), ("v0000", "pqrs")],      # SV_1mer_0.py:24-31
}

_PM = [(+1, (0, 1)), (-1, (1, 0))]
_PP = [(+1, (0, 1)), (+1, (1, 0))]
# label: (Dchgs, permutations)   -- S_diagrams.py:33-40, ST_diagrams.py:35-40, SU_diagrams.py:35-46, SV_diagrams.py:35-48
CATALOG2 = {
    "t01": ((-1, 1), _PM), "u001": ((-1, 1), _PM), "u101": ((-1, 1), _PM), "u100": ((0, 0), _PP),
    "v0101": ((0, 0), [(+1, (0, 1))]), "v0001": ((-1, 1), _PM), "v0100": ((1, -1), _PM), "v0011": ((-2, 2), _PP),
    "s01": ((-1, 1), _PM),
    "s01t10": ((0, 0), _PP), "s01t00": ((-1, 1), _PM), "s01t11": ((-1, 1), _PM), "s01t01": ((-2, 2), _PP),
    "s01v0100": ((0, 0), _PP), "s01v1101": ((0, 0), _PP), "s01v0000": ((-1, 1), _PM), "s01v0101": ((-1, 1), _PM),
    "s01v1100": ((1, -1), _PM), "s01v1111": ((-1, 1), _PM), "s01v0001": ((-2, 2), _PP), "s01v0111": ((-2, 2), _PP),
}
CATALOG2.update({
    "s01s10": ((0, 0), [(+1, (0, 1))]), "s01s01": ((-2, 2), _PP),
    "s01s10t00": ((0, 0), _PP), "s01s01t10": ((-1, 1), _PM), "s01s10t01": ((-1, 1), _PM), "s01s01t00": ((-2, 2), _PP),
    "s01s01t11": ((-2, 2), _PP),
    "s01s01v1100": ((0, 0), _PP), "s01s10v0000": ((0, 0), _PP), "s01s10v0101": ((0, 0), [(+1, (0, 1))]),
    "s01s01v0100": ((-1, 1), _PM), "s01s01v1101": ((-1, 1), _PM), "s01s10v0001": ((-1, 1), _PM), "s01s10v0100": ((1, -1), _PM),
    "s01s01v0000": ((-2, 2), _PP), "s01s01v0101": ((-2, 2), _PP), "s01s01v1111": ((-2, 2), _PP), "s01s10v0011": ((-2, 2), _PP),
})
_P1 = [(+1, (0, 1))]
CATALOG2.update({
    "s01s01s10": ((-1, 1), _PM), "s01s01s10s10": ((0, 0), _P1), "s01s01s01s10": ((-2, 2), _PP),
    "s01s01s10t10": ((0, 0), _PP), "s01s01s10t00": ((-1, 1), _PM), "s01s01s10t11": ((-1, 1), _PM), "s01s01s01t10": ((-2, 2), _PP),
    "s01s01s10t01": ((-2, 2), _PP), "s01s01s10s10t00": ((0, 0), _PP), "s01s01s01s10t10": ((-1, 1), _PM),
    "s01s01s10s10t01": ((-1, 1), _PM), "s01s01s01s10t00": ((-2, 2), _PP), "s01s01s01s10t11": ((-2, 2), _PP),
    "s01s01s10v0100": ((0, 0), _PP), "s01s01s10v1101": ((0, 0), _PP), "s01s01s01v1100": ((-1, 1), _PM), "s01s01s10v0000": ((-1, 1), _PM),
    "s01s01s10v0101": ((-1, 1), _PM), "s01s01s10v1100": ((1, -1), _PM), "s01s01s10v1111": ((-1, 1), _PM), "s01s01s01v0100": ((-2, 2), _PP),
    "s01s01s01v1101": ((-2, 2), _PP), "s01s01s10v0001": ((-2, 2), _PP), "s01s01s10v0111": ((-2, 2), _PP),
    "s01s01s01s10v1100": ((0, 0), _PP), "s01s01s10s10v0000": ((0, 0), _PP), "s01s01s10s10v0101": ((0, 0), _P1),
    "s01s01s01s10v0100": ((-1, 1), _PM), "s01s01s01s10v1101": ((-1, 1), _PM), "s01s01s10s10v0001": ((-1, 1), _PM),
    "s01s01s10s10v0100": ((1, -1), _PM), "s01s01s01s01v1100": ((-2, 2), _PP), "s01s01s01s10v0000": ((-2, 2), _PP),
    "s01s01s01s10v0101": ((-2, 2), _PP), "s01s01s01s10v1111": ((-2, 2), _PP), "s01s01s10s10v0011": ((-2, 2), _PP),
})
for _n in "01":
    for _t in ("s01s01s10t10", "s01s01s10t00", "s01s01s10t11", "s01s01s01t10", "s01s01s10t01",
               "s01s01s10s10t00", "s01s01s01s10t10", "s01s01s10s10t01", "s01s01s01s10t00", "s01s01s01s10t11"):
        _k = _t.index("t")
        CATALOG2[_t[:_k] + "u" + _n + _t[_k + 1:]] = CATALOG2[_t]
for _n in "01":
    CATALOG2["s01s10u%s00" % _n] = ((0, 0), _PP)
    CATALOG2["s01s01u%s10" % _n] = ((-1, 1), _PM)
    CATALOG2["s01s10u%s01" % _n] = ((-1, 1), _PM)
    CATALOG2["s01s01u%s00" % _n] = ((-2, 2), _PP)
    CATALOG2["s01s01u%s11" % _n] = ((-2, 2), _PP)
for _n in "01":
    CATALOG2["s01u%s10" % _n] = ((0, 0), _PP)
    CATALOG2["s01u%s00" % _n] = ((-1, 1), _PM)
    CATALOG2["s01u%s11" % _n] = ((-1, 1), _PM)
    CATALOG2["s01u%s01" % _n] = ((-2, 2), _PP)

LISTS = {   # diagram_lists.py:10-71, orders 0 and 1
    "S0": {0: ["identity"]}, "S2": {1: ["s01"], 2: ["s01s10", "s01s01"]},
    "ST1": {0: ["t00"]}, "ST2": {0: ["t01"], 1: ["s01t10", "s01t00", "s01t11", "s01t01"],
                                  2: ["s01s10t00", "s01s01t10", "s01s10t01", "s01s01t00", "s01s01t11"]},
    "SU1": {0: ["u000"]}, "SU2": {0: ["u100", "u001", "u101"],
                                  1: ["s01u010", "s01u000", "s01u011", "s01u001", "s01u110", "s01u100", "s01u111", "s01u101"],
                                  2: ["s01s10u000", "s01s01u010", "s01s10u001", "s01s01u000", "s01s01u011",
                                      "s01s10u100", "s01s01u110", "s01s10u101", "s01s01u100", "s01s01u111"]},
    "SV1": {0: ["v0000"]}, "SV2": {0: ["v0101", "v0001", "v0100", "v0011"],
                                   1: ["s01v0100", "s01v1101", "s01v0000", "s01v0101", "s01v1100", "s01v1111", "s01v0001", "s01v0111"],
                                   2: ["s01s01v1100", "s01s10v0000", "s01s10v0101", "s01s01v0100", "s01s01v1101", "s01s10v0001",
                                       "s01s10v0100", "s01s01v0000", "s01s01v0101", "s01s01v1111", "s01s10v0011"]},
}


class integrals(object):
    """S, T, U, V (+ V_diff) blocked containers with ndarray blocks"""
    def __init__(self, S, T=None, U=None, V=None):
        self.S, self.T, self.U, self.V = S, T, U, V


def _tensor(name, dens, ints, frags, chgs):
    """resolve a tensor name for diagram fragments -> absolute fragments `frags`, charges `chgs`"""
    digits = [int(c) for c in name if c.isdigit()]
    head = name.rstrip("0123456789_")
    if name[0] in "stuv" and all(c.isdigit() or c == "_" for c in name[1:]):
        block = tuple(frags[d] for d in digits)
        return numpy.asarray({"s": ints.S, "t": ints.T, "u": ints.U, "v": ints.V}[name[0]][block])
    op, d = name[:-1], int(name[-1])
    return numpy.asarray(dens[frags[d]][op][chgs[d]])


def _einsum(terms, dens, ints, frags, chgs, out):
    operands = [_tensor(name, dens, ints, frags, chgs) for name, _ in terms]
    return numpy.einsum(",".join(idx for _, idx in terms) + "->" + out, *operands, optimize=True)


def dimer_block(label, dens, ints, subsystem, charges):
    """block [N_i0, N_i1, N_j0, N_j1] (subsystem order) of a 2-fragment diagram, or None.
    subsystem = (m0, m1) ascending; charges = ((chg_i0, chg_j0), (chg_i1, chg_j1))."""
    Dchgs, permutations = CATALOG2[label]
    n_j0 = dens[subsystem[0]]["n_elec"][charges[0][1]] % 2           # unpermuted first fragment (build_diagram.py:87-92)
    total = None
    for phase, perm in permutations:
        frags = [subsystem[m] for m in perm]
        chgs = [charges[m] for m in perm]
        if any(chgs[m][0] - chgs[m][1] != Dchgs[m] for m in range(2)):
            continue
        if label == "u100":                                          # SU_2mer_0.py:37-48
            blk = _einsum([("ca0", "ijpq"), ("u1_00", "pq")], dens, ints, frags, chgs, "ij")
            n1 = dens[frags[1]]["n_states"][chgs[1][1]]
            val = numpy.einsum("ij,kl->ikjl", blk, numpy.eye(n1))
        else:
            coef, shift, terms = TWO_FRAGMENT[label]
            val = coef * _einsum(terms, dens, ints, frags, chgs, "ikjl")
            if shift is not None:
                val = val * (-1) ** (n_j0 + shift)
        val = phase * val.transpose(list(perm) + [2 + m for m in perm])      # diagrammatic_expansion.py:32,58
        total = val if total is None else total + val
    return total


def monomer_block(label, dens, ints, m, chg_i, chg_j):
    if chg_i != chg_j:
        return None
    return _einsum(ONE_FRAGMENT[label], dens, ints, [m], [(chg_i, chg_j)], "ij")


def monomer_matrix(dens, ints, labels, m, charge_blocks):
    """XR_term.py:96-115"""
    n = dens[m]["n_states"]
    off = numpy.concatenate([[0], numpy.cumsum([n[c] for c in charge_blocks])])
    M = numpy.zeros((off[-1], off[-1]))
    for a, ci in enumerate(charge_blocks):
        for label in labels:
            M[off[a]:off[a + 1], off[a]:off[a + 1]] += monomer_block(label, dens, ints, m, ci, ci)
    return M


def dimer_matrix(dens, ints, active, charge_blocks, subsystem=(0, 1)):
    """XR_term.py:117-165 + _evaluate_block :22-94, charge-blocked ordering; active = {frag_order: [labels]}"""
    m0, m1 = subsystem
    n0, n1 = dens[m0]["n_states"], dens[m1]["n_states"]
    sizes = [n0[c0] * n1[c1] for c0, c1 in charge_blocks]
    off = numpy.concatenate([[0], numpy.cumsum(sizes)])
    M = numpy.zeros((off[-1], off[-1]))
    for a, (ci0, ci1) in enumerate(charge_blocks):
        for b, (cj0, cj1) in enumerate(charge_blocks):
            shape = (n0[ci0], n1[ci1], n0[cj0], n1[cj1])
            block = numpy.zeros(shape)
            for frag_order, labels in active.items():
                for label in labels:
                    if frag_order == 0:
                        if ci0 == cj0 and ci1 == cj1:
                            block += numpy.einsum("ij,kl->ikjl", numpy.eye(shape[0]), numpy.eye(shape[1]))
                    elif frag_order == 1:
                        if ci0 == cj0 and ci1 == cj1:
                            block += numpy.einsum("ij,kl->ikjl", monomer_block(label, dens, ints, m0, ci0, cj0), numpy.eye(shape[1]))
                            block += numpy.einsum("ij,kl->ikjl", numpy.eye(shape[0]), monomer_block(label, dens, ints, m1, ci1, cj1))
                    elif ci0 + ci1 == cj0 + cj1:
                        val = dimer_block(label, dens, ints, subsystem, ((ci0, cj0), (ci1, cj1)))
                        if val is not None:
                            block += val
            M[off[a]:off[a + 1], off[b]:off[b + 1]] += block.reshape(shape[0] * shape[1], shape[2] * shape[3])
    return M


def reorder(Hblocked, dens, monomer_charges):
    """get_xr_result.py:300-353: charge-blocked -> (global i0, global i1) row-major"""
    dims0 = [dens[0]["n_states"][c] for c in monomer_charges[0]]
    dims1 = [dens[1]["n_states"][c] for c in monomer_charges[1]]
    mapping = numpy.zeros((sum(dims0), sum(dims1)), dtype=int)
    idx, beg0 = 0, 0
    for d0 in dims0:
        beg1 = 0
        for d1 in dims1:
            for a in range(d0):
                for b in range(d1):
                    mapping[beg0 + a, beg1 + b] = idx
                    idx += 1
            beg1 += d1
        beg0 += d0
    order = mapping.reshape(-1)
    return Hblocked[numpy.ix_(order, order)]


def get_xr_H(symm, bior, dens, xr_order, monomer_charges):
    """get_xr_result.py:45-355 for xr_order 0 and 1.  symm/bior: objects with S,T,U,V (bior also V_diff)."""
    charges = [(c0, c1) for c0 in monomer_charges[0] for c1 in monomer_charges[1]]
    L = LISTS
    mk = lambda T=None, U=None, V=None: integrals(symm.S, T, U, V)
    if xr_order == 0:
        H1 = [monomer_matrix(dens, mk(T=bior.T), L["ST1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(U=bior.U), L["SU1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(V=bior.V), L["SV1"][0], m, monomer_charges[m]) for m in (0, 1)]
        H2 = (dimer_matrix(dens, mk(T=bior.T), {2: L["ST2"][0]}, charges) + dimer_matrix(dens, mk(U=bior.U), {2: L["SU2"][0]}, charges)
              + dimer_matrix(dens, mk(V=bior.V), {2: L["SV2"][0]}, charges))
    elif xr_order == 1:
        H1 = [monomer_matrix(dens, mk(T=symm.T), L["ST1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(U=symm.U), L["SU1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(V=symm.V), L["SV1"][0], m, monomer_charges[m]) for m in (0, 1)]
        S2 = dimer_matrix(dens, mk(), {0: L["S0"][0], 2: L["S2"][1]}, charges)
        S2inv = numpy.linalg.inv(S2)
        S2H2 = dimer_matrix(dens, mk(T=symm.T), {1: L["ST1"][0], 2: L["ST2"][0]}, charges)
        S2H2 += dimer_matrix(dens, mk(U=symm.U), {1: L["SU1"][0], 2: L["SU2"][0]}, charges)
        S2H2 += dimer_matrix(dens, mk(T=bior.T), {2: L["ST2"][1]}, charges)
        S2H2 += dimer_matrix(dens, mk(U=bior.U), {2: L["SU2"][1]}, charges)
        S2H2 += dimer_matrix(dens, mk(V=bior.V_diff), {1: L["SV1"][0], 2: L["SV2"][0]}, charges)
        S2H2 += dimer_matrix(dens, mk(V=bior.V), {2: L["SV2"][1]}, charges)
        H2 = S2inv @ S2H2
        H2 -= dimer_matrix(dens, mk(T=symm.T), {1: L["ST1"][0]}, charges)
        H2 -= dimer_matrix(dens, mk(U=symm.U), {1: L["SU1"][0]}, charges)
        H2 -= dimer_matrix(dens, mk(V=symm.V), {1: L["SV1"][0]}, charges)
    elif xr_order == 2:                                  # get_xr_result.py:214-296
        H1 = [monomer_matrix(dens, mk(T=symm.T), L["ST1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(U=symm.U), L["SU1"][0], m, monomer_charges[m])
              + monomer_matrix(dens, mk(V=symm.V), L["SV1"][0], m, monomer_charges[m]) for m in (0, 1)]
        S2 = dimer_matrix(dens, mk(), {0: L["S0"][0], 2: L["S2"][1] + L["S2"][2]}, charges)
        S2inv = numpy.linalg.inv(S2)
        S2H2 = dimer_matrix(dens, mk(T=symm.T), {1: L["ST1"][0], 2: L["ST2"][0] + L["ST2"][1]}, charges)
        S2H2 += dimer_matrix(dens, mk(U=symm.U), {1: L["SU1"][0], 2: L["SU2"][0] + L["SU2"][1]}, charges)
        S2H2 += dimer_matrix(dens, mk(T=bior.T), {2: L["ST2"][2]}, charges)
        S2H2 += dimer_matrix(dens, mk(U=bior.U), {2: L["SU2"][2]}, charges)
        S2H2 += dimer_matrix(dens, mk(V=symm.V), {1: L["SV1"][0], 2: L["SV2"][0]}, charges)
        S2H2 += dimer_matrix(dens, mk(V=bior.V_diff), {2: L["SV2"][1]}, charges)
        S2H2 += dimer_matrix(dens, mk(V=bior.V), {2: L["SV2"][2]}, charges)
        H2 = S2inv @ S2H2
        H2 -= dimer_matrix(dens, mk(T=symm.T), {1: L["ST1"][0]}, charges)
        H2 -= dimer_matrix(dens, mk(U=symm.U), {1: L["SU1"][0]}, charges)
        H2 -= dimer_matrix(dens, mk(V=symm.V), {1: L["SV1"][0]}, charges)
    else:
        raise NotImplementedError(xr_order)
    return H1, reorder(H2, dens, monomer_charges)
