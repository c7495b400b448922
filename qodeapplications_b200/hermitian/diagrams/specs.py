"""Table-driven diagram definitions and the one routine that evaluates them on the GPU.

Every two-fragment diagram of the reference (hermitian-XRCC/diagrams/S*_2mer_*.py) has the form

    coefficient * (-1)**(X.n_j0 + shift) * raw( A(i0,j0, ...) @ B(i1,j1, ...) [@ s01(v,w) | s10(v,w)] )

with A a density or rho x integral precontraction of diagram fragment 0, B the same for fragment 1,
from S-order 2 on one or more bare overlap blocks as further factors, and equal letters contracted.  A row of the
table is ``label: (coefficient, parity shift or None, [(operand name, free letters), ...])``; the
operand names are exactly the attribute names the reference reads from ``X`` (build_diagram.py:116-134),
so ``frag_resolve`` resolves them the same way.  A row becomes a function with the reference's signature
``fn(X, contract_last=False) -> ndarray[N_i0, N_i1, N_j0, N_j1]``; the product is evaluated by
``Contractor.multi_contract`` (pairwise xr_gemm_scatter calls, the last one writing -- or accumulating in
place -- the [i0,i1,j0,j1] block).  One-fragment diagrams are a single precontraction; ``u100`` carries a
Kronecker delta over the spectator fragment (SU_2mer_0.py:26-48); ``identity`` is the 0-mer
(S_0mer_0.py:21-22).  File:line of the reference definition is given per row or group.
"""
import re
import numpy

from ..tensor import DeviceTensor

H = 0.5
# label: (coef, parity shift (None = no n_j0 phase), [(operand, free letters), ...])
TWO_FRAGMENT = {
    # ---- order 0 --------------------------------------------------------------------------------
    "t01":      (1, 0,     [("c0p_Tp1", "q"), ("a1", "q")]),                                  # ST_2mer_0.py:25-34
    "u001":     (1, 0,     [("c0p_U0p1", "q"), ("a1", "q")]),                                 # SU_2mer_0.py:98-107
    "u101":     (1, 0,     [("c0p_U1p1", "q"), ("a1", "q")]),                                 # SU_2mer_0.py:109-118
    "v0101":    (4, None,  [("ca0pr_Vp1r1", "qs"), ("ca1", "qs")]),                           # SV_2mer_0.py:25-34
    "v0001":    (2, 1,     [("cca0pqr_Vpqr1", "s"), ("a1", "s")]),                            # SV_2mer_0.py:36-45
    "v0100":    (2, 0,     [("caa0psr_Vp1rs", "q"), ("c1", "q")]),                            # SV_2mer_0.py:47-56
    "v0011":    (1, None,  [("cc0pq_Vpq11", "rs"), ("aa1", "sr")]),                           # SV_2mer_0.py:58-67
    # ---- order 1 --------------------------------------------------------------------------------
    "s01":      (1, 0,     [("c0p_Sp1", "q"), ("a1", "q")]),                                  # S_2mer_1.py:25-34
    "s01t10":   (-1, None, [("ca0Xq_T1q", "tp"), ("ca1Xu_S0u", "pt")]),                       # ST_2mer_1.py:25-36
    "s01t00":   (1, 1,     [("cca0pXq_Tpq", "t"), ("a1u_S0u", "t")]),                         # ST_2mer_1.py:38-49
    "s01t11":   (1, 1,     [("c0t_St1", "u"), ("caa1pXq_Tpq", "u")]),                         # ST_2mer_1.py:51-62
    "s01t01":   (1, None,  [("cc0pX_Tp1", "tq"), ("aa1uX_S0u", "qt")]),                       # ST_2mer_1.py:64-75
    "s01v0100": (-2, None, [("ccaa0pXsr_Vp1rs", "tq"), ("ca1Xu_S0u", "qt")]),                 # SV_2mer_1.py:25-36
    "s01v1101": (2, None,  [("ca0tX_St1", "ru"), ("ccaa1pqXs_Vpq0s", "ur")]),                 # SV_2mer_1.py:38-49
    "s01v0000": (1, 0,     [("cccaa0pqXsr_Vpqrs", "t"), ("a1u_S0u", "t")]),                   # SV_2mer_1.py:51-62
    "s01v0101": (4, 0,     [("cca0pXr_Vp1r1", "tqs"), ("caa1XuX_S0u", "qst")]),               # SV_2mer_1.py:64-75
    "s01v1100": (1, 0,     [("caa0Xsr_V11rs", "tpq"), ("cca1XXu_S0u", "pqt")]),               # SV_2mer_1.py:77-88
    "s01v1111": (1, 0,     [("c0t_St1", "u"), ("ccaaa1pqXsr_Vpqrs", "u")]),                   # SV_2mer_1.py:90-101
    "s01v0001": (2, None,  [("ccca0pqXr_Vpqr1", "ts"), ("aa1uX_S0u", "st")]),                 # SV_2mer_1.py:103-114
    "s01v0111": (-2, None, [("cc0Xt_St1", "pu"), ("caaa1qXsr_V0qrs", "up")]),                 # SV_2mer_1.py:116-127
    "s01v0011": (1, 0,     [("ccc0pqX_Vpq11", "trs"), ("aaa1uXX_S0u", "srt")]),               # SV_2mer_1.py:129-140
    # ---- order 2 --------------------------------------------------------------------------------
    "s01s10":   (-1, None, [("ca0pX_Sp1", "sq"), ("ca1rX_Sr0", "qs")]),                       # S_2mer_2.py:25-36
    "s01s01":   (H, None,  [("cc0pX_Sp1", "rq"), ("aa1sX_S0s", "qr")]),                       # S_2mer_2.py:38-49
    "s01s10t00": (-1, None, [("ccaa0pXXq_Tpq", "tw"), ("ca1Xu_S0u", "vt"), ("s10", "vw")]),   # ST_2mer_2.py:25-38
    "s01s01t10": (H, 1,    [("cca0XXq_T1q", "tvp"), ("caa1XXu_S0u", "pwt"), ("s01", "vw")]),  # ST_2mer_2.py:40-53
    "s01s10t01": (1, 1,    [("cca0pXX_Tp1", "twq"), ("caa1XuX_S0u", "vqt"), ("s10", "vw")]),  # ST_2mer_2.py:55-68
    "s01s01t00": (H, None, [("ccca0pXXq_Tpq", "tv"), ("aa1Xu_S0u", "wt"), ("s01", "vw")]),    # ST_2mer_2.py:70-83
    "s01s01t11": (H, None, [("cc0tX_St1", "vu"), ("caaa1pXXq_Tpq", "wu"), ("s01", "vw")]),    # ST_2mer_2.py:85-98
    "s01s01v1100": (H, None, [("ccaa0XXsr_V11rs", "tvpq"), ("ccaa1XXXu_S0u", "pqwt"), ("s01", "vw")]),   # SV_2mer_2.py
    "s01s10v0000": (-1, None, [("cccaaa0pqXXsr_Vpqrs", "tw"), ("ca1Xu_S0u", "vt"), ("s10", "vw")]),
    "s01s10v0101": (-4, None, [("ccaa0pXXr_Vp1r1", "twqs"), ("ccaa1XXuX_S0u", "qvst"), ("s10", "vw")]),
    "s01s01v0100": (1, 0,  [("cccaa0pXXsr_Vp1rs", "tvq"), ("caa1XXu_S0u", "qwt"), ("s01", "vw")]),
    "s01s01v1101": (1, 1,  [("cca0tXX_St1", "vru"), ("ccaaa1pqXXs_Vpq0s", "wur"), ("s01", "vw")]),
    "s01s10v0001": (2, 0,  [("cccaa0pqXXr_Vpqr1", "tws"), ("caa1XuX_S0u", "vst"), ("s10", "vw")]),
    "s01s10v0100": (2, 1,  [("ccaaa0pXXsr_Vp1rs", "twq"), ("cca1XXu_S0u", "qvt"), ("s10", "vw")]),
    "s01s01v0000": (H, None, [("ccccaa0pqXXsr_Vpqrs", "tv"), ("aa1Xu_S0u", "wt"), ("s01", "vw")]),
    "s01s01v0101": (2, None, [("ccca0pXXr_Vp1r1", "tvqs"), ("caaa1XXuX_S0u", "qwst"), ("s01", "vw")]),
    "s01s01v1111": (H, None, [("cc0tX_St1", "vu"), ("ccaaaa1pqXXsr_Vpqrs", "wu"), ("s01", "vw")]),
    "s01s10v0011": (-1, None, [("ccca0pqXX_Vpq11", "twrs"), ("caaa1XuXX_S0u", "vsrt"), ("s10", "vw")]),
    # ---- orders 3 and 4 (the diagrams diagram_lists.py:10-71 activates) ------------------------------------------
    "s01s01s10":         (1 / 2, 1, [("cca0pXX_Sp1", "ruq"), ("caa1XsX_S0s", "tqr"), ("s10", "tu")]),   # S_2mer_3.py
    "s01s01s10s10":      (1 / 4, None, [("ccaa0pXXX_Sp1", "rwuq"), ("ccaa1XXsX_S0s", "tvqr"), ("s10", "tu"), ("s10", "vw")]),   # S_2mer_4.py
    "s01s01s01s10":      (-1 / 6, None, [("ccca0pXXX_Sp1", "rtwq"), ("caaa1XXsX_S0s", "vuqr"), ("s01", "tu"), ("s10", "vw")]),   # S_2mer_4.py
    "s01s01s10t10":      (1 / 2, None, [("ccaa0XXXq_T1q", "tvyp"), ("ccaa1XXXu_S0u", "pxwt"), ("s01", "vw"), ("s10", "xy")]),   # ST_2mer_3.py
    "s01s01s10t00":      (1 / 2, 0, [("cccaa0pXXXq_Tpq", "tvy"), ("caa1XXu_S0u", "xwt"), ("s01", "vw"), ("s10", "xy")]),   # ST_2mer_3.py
    "s01s01s10t11":      (1 / 2, 0, [("cca0tXX_St1", "vyu"), ("ccaaa1pXXXq_Tpq", "xwu"), ("s01", "vw"), ("s10", "xy")]),   # ST_2mer_3.py
    "s01s01s01t10":      (-1 / 6, None, [("ccca0XXXq_T1q", "tvxp"), ("caaa1XXXu_S0u", "pywt"), ("s01", "vw"), ("s01", "xy")]),   # ST_2mer_3.py
    "s01s01s10t01":      (-1 / 2, None, [("ccca0pXXX_Tp1", "tvyq"), ("caaa1XXuX_S0u", "xwqt"), ("s01", "vw"), ("s10", "xy")]),   # ST_2mer_3.py
    "s01s01s10s10t00":   (1 / 4, None, [("cccaaa0pXXXXq_Tpq", "tvay"), ("ccaa1XXXu_S0u", "xzwt"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # ST_2mer_4.py
    "s01s01s01s10t10":   (1 / 6, 0, [("cccaa0XXXXq_T1q", "tvxap"), ("ccaaa1XXXXu_S0u", "pzywt"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # ST_2mer_4.py
    "s01s01s10s10t01":   (1 / 4, 0, [("cccaa0pXXXX_Tp1", "tvayq"), ("ccaaa1XXXuX_S0u", "xzwqt"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # ST_2mer_4.py
    "s01s01s01s10t00":   (-1 / 6, None, [("ccccaa0pXXXXq_Tpq", "tvxa"), ("caaa1XXXu_S0u", "zywt"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # ST_2mer_4.py
    "s01s01s01s10t11":   (-1 / 6, None, [("ccca0tXXX_St1", "vxau"), ("ccaaaa1pXXXXq_Tpq", "zywu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # ST_2mer_4.py
    "s01s01s10v0100":    (1, None, [("cccaaa0pXXXsr_Vp1rs", "tvyq"), ("ccaa1XXXu_S0u", "qxwt"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s10v1101":    (-1, None, [("ccaa0tXXX_St1", "vyru"), ("cccaaa1pqXXXs_Vpq0s", "xwur"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s01v1100":    (1 / 6, 0, [("cccaa0XXXsr_V11rs", "tvxpq"), ("ccaaa1XXXXu_S0u", "pqywt"), ("s01", "vw"), ("s01", "xy")]),   # SV_2mer_3.py
    "s01s01s10v0000":    (1 / 2, 1, [("ccccaaa0pqXXXsr_Vpqrs", "tvy"), ("caa1XXu_S0u", "xwt"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s10v0101":    (2, 1, [("cccaa0pXXXr_Vp1r1", "tvyqs"), ("ccaaa1XXXuX_S0u", "qxwst"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s10v1100":    (1 / 2, 1, [("ccaaa0XXXsr_V11rs", "tvypq"), ("cccaa1XXXXu_S0u", "pqxwt"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s10v1111":    (1 / 2, 1, [("cca0tXX_St1", "vyu"), ("cccaaaa1pqXXXsr_Vpqrs", "xwu"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s01v0100":    (-1 / 3, None, [("ccccaa0pXXXsr_Vp1rs", "tvxq"), ("caaa1XXXu_S0u", "qywt"), ("s01", "vw"), ("s01", "xy")]),   # SV_2mer_3.py
    "s01s01s01v1101":    (1 / 3, None, [("ccca0tXXX_St1", "vxru"), ("ccaaaa1pqXXXs_Vpq0s", "ywur"), ("s01", "vw"), ("s01", "xy")]),   # SV_2mer_3.py
    "s01s01s10v0001":    (-1, None, [("ccccaa0pqXXXr_Vpqr1", "tvys"), ("caaa1XXuX_S0u", "xwst"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s10v0111":    (1, None, [("ccca0XtXX_St1", "pvyu"), ("ccaaaa1qXXXsr_V0qrs", "xwup"), ("s01", "vw"), ("s10", "xy")]),   # SV_2mer_3.py
    "s01s01s01s10v1100": (-1 / 6, None, [("cccaaa0XXXXsr_V11rs", "tvxapq"), ("cccaaa1XXXXXu_S0u", "pqzywt"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s10s10v0000": (1 / 4, None, [("ccccaaaa0pqXXXXsr_Vpqrs", "tvay"), ("ccaa1XXXu_S0u", "xzwt"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s10s10v0101": (1, None, [("cccaaa0pXXXXr_Vp1r1", "tvayqs"), ("cccaaa1XXXXuX_S0u", "qxzwst"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s01s10v0100": (1 / 3, 1, [("ccccaaa0pXXXXsr_Vp1rs", "tvxaq"), ("ccaaa1XXXXu_S0u", "qzywt"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s01s10v1101": (1 / 3, 0, [("cccaa0tXXXX_St1", "vxaru"), ("cccaaaa1pqXXXXs_Vpq0s", "zywur"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s10s10v0001": (1 / 2, 1, [("ccccaaa0pqXXXXr_Vpqr1", "tvays"), ("ccaaa1XXXuX_S0u", "xzwst"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s10s10v0100": (1 / 2, 0, [("cccaaaa0pXXXXsr_Vp1rs", "tvayq"), ("cccaa1XXXXu_S0u", "qxzwt"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s01s01v1100": (1 / 24, None, [("ccccaa0XXXXsr_V11rs", "tvxzpq"), ("ccaaaa1XXXXXu_S0u", "pqaywt"), ("s01", "vw"), ("s01", "xy"), ("s01", "za")]),   # SV_2mer_4.py
    "s01s01s01s10v0000": (-1 / 6, None, [("cccccaaa0pqXXXXsr_Vpqrs", "tvxa"), ("caaa1XXXu_S0u", "zywt"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s01s10v0101": (-2 / 3, None, [("ccccaa0pXXXXr_Vp1r1", "tvxaqs"), ("ccaaaa1XXXXuX_S0u", "qzywst"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s01s10v1111": (-1 / 6, None, [("ccca0tXXX_St1", "vxau"), ("cccaaaaa1pqXXXXsr_Vpqrs", "zywu"), ("s01", "vw"), ("s01", "xy"), ("s10", "za")]),   # SV_2mer_4.py
    "s01s01s10s10v0011": (1 / 4, None, [("ccccaa0pqXXXX_Vpq11", "tvayrs"), ("ccaaaa1XXXuXX_S0u", "xzwsrt"), ("s01", "vw"), ("s10", "xy"), ("s10", "za")]),   # SV_2mer_4.py
}
for _n in "01":     # SU_2mer_1.py / SU_2mer_2.py: the ST rows with T -> U<nucleus fragment>
    for _t, _u in (("s01t10", "s01u%s10"), ("s01t00", "s01u%s00"), ("s01t11", "s01u%s11"), ("s01t01", "s01u%s01"),
                   ("s01s10t00", "s01s10u%s00"), ("s01s01t10", "s01s01u%s10"), ("s01s10t01", "s01s10u%s01"),
                   ("s01s01t00", "s01s01u%s00"), ("s01s01t11", "s01s01u%s11"),
                   ("s01s01s10t10", "s01s01s10u%s10"), ("s01s01s10t00", "s01s01s10u%s00"), ("s01s01s10t11", "s01s01s10u%s11"),
                   ("s01s01s01t10", "s01s01s01u%s10"), ("s01s01s10t01", "s01s01s10u%s01"),
                   ("s01s01s10s10t00", "s01s01s10s10u%s00"), ("s01s01s01s10t10", "s01s01s01s10u%s10"),
                   ("s01s01s10s10t01", "s01s01s10s10u%s01"), ("s01s01s01s10t00", "s01s01s01s10u%s00"),
                   ("s01s01s01s10t11", "s01s01s01s10u%s11")):
        _c, _s, _ops = TWO_FRAGMENT[_t]
        TWO_FRAGMENT[_u % _n] = (_c, _s, [(_name.replace("_T", "_U" + _n), _idx) for _name, _idx in _ops])

ONE_FRAGMENT = {
    "t00":   "ca0pq_Tpq",          # ST_1mer_0.py:24
    "u000":  "ca0pq_U0pq",         # SU_1mer_0.py:24
    "v0000": "ccaa0pqsr_Vpqrs",    # SV_1mer_0.py:24-31
}


def _state_axes(contract_last):
    """diagram_hack.state_indices (:26-33): the free state axes of the result and the renaming that makes the two
    ket (or bra) state indices one contracted label when the last pair is traced"""
    if contract_last == "ket":
        return {"j0": "_tr", "j1": "_tr"}, ["i0", "i1"]
    if contract_last == "bra":
        return {"i0": "_tr", "i1": "_tr"}, ["j0", "j1"]
    if contract_last:
        raise ValueError("contract_last must be False, 'ket' or 'bra'")
    return {}, ["i0", "i1", "j0", "j1"]


def no_result(X, contract):
    """diagram_hack.no_result (:35-41): a trace over two state indices of different length does not exist"""
    (i0s, j0s), (i1s, j1s) = X.n_states[0], X.n_states[1]
    if contract == "ket":
        return j0s != j1s
    if contract == "bra":
        return i0s != i1s
    return False


def _contractor(X):
    return X._info.contract_cache.general.contractor


def _state_labels(name):
    """a density / precontraction of diagram fragment d carries (bra, ket) state axes i<d>, j<d>; a bare integral none"""
    m = re.match(r"^[ca]+(\d)", name)
    return ["i" + m.group(1), "j" + m.group(1)] if m else []


def make_two_fragment(label):
    coef, shift, operands = TWO_FRAGMENT[label]

    def factors(X, contract_last):
        rename, free_axes = _state_axes(contract_last)
        out = []
        for name, free in operands:
            T = getattr(X, name)
            if T is None:
                raise RuntimeError("diagram %s: operand %s is not available for these charges" % (label, name))
            out.append((T, [rename.get(l, l) for l in _state_labels(name)] + list(free)))
        factor = float(coef) if shift is None else float(coef) * (-1.0) ** (X.n_j0 + shift)
        return out, factor, free_axes

    def contraction(X, contract_last=False):
        if no_result(X, contract_last):
            return []
        ops, factor, free_axes = factors(X, contract_last)
        return _contractor(X).multi_contract(ops, free_axes, alpha=factor).host()

    def accumulate(X, phase, out, offset, strides, contract_last=False):
        """out[offset + sum over free state axes a of a*strides[a]] += phase * diagram; free axes are i0,i1,j0,j1, or
        only i0,i1 / j0,j1 when the ket / bra pair is traced (contract_last).  False if that trace does not exist."""
        if no_result(X, contract_last):
            return False
        ops, factor, free_axes = factors(X, contract_last)
        _contractor(X).multi_contract(ops, free_axes, alpha=phase * factor, out=out, out_offset=offset,
                                      out_strides={a: strides[a] for a in free_axes}, accumulate=True)
        return True

    contraction.__name__ = label
    contraction.accumulate = accumulate
    contraction.spec = TWO_FRAGMENT[label]
    return contraction


def make_one_fragment(label):
    name = ONE_FRAGMENT[label]

    def contraction(X):
        return getattr(X, name).host()

    def accumulate(X, phase, out, offset, strides):
        """out[offset + i0*strides['i0'] + j0*strides['j0'] + d*strides['delta']] += phase * block[i0,j0]
        for d < strides['n_delta'] (d runs over the spectator states the block is diagonal in; XR_term.py:69-80)"""
        block = getattr(X, name)
        C = _contractor(X)
        n_delta = int(strides.get("n_delta", 1))
        ones = C.dev.upload(numpy.ones((n_delta, 1)))
        C.contract(block, ["i0", "j0"], DeviceTensor(ones, C.dev), ["d", "one"], ["i0", "j0", "d", "one"], alpha=phase,
                   out=out, out_offset=offset,
                   out_strides={"i0": strides["i0"], "j0": strides["j0"], "d": strides.get("delta", 0), "one": 0},
                   accumulate=True)

    contraction.__name__ = label
    contraction.accumulate = accumulate
    return contraction


def _u100_pieces(X, special_processing):
    """SU_2mer_0.py:37-96.  Returns (res, K, free axes, transpose?) for the traced variants: with
    res[i0,j0] = sum_pq ca0[i0,j0,p,q] U[frag1,frag0,frag0][p,q],
      bra_det (special_processing 0/1): out[i0,i1] = res[i0,i1]                 if N_i1 == N_j1
                                                   = sum_z res[i0,z] K[z,i1]    otherwise   (K = KetCoeffs1[j1,i1])
      ket_det (special_processing 2/3): out[j0,j1] = res[j1,j0]                 if N_i1 == N_j1
                                                   = sum_z res[z,j0] K[z,j1]    otherwise   (K = KetCoeffs1[i1,j1])
    and the odd values (1, 3) return the transposed array (the reference transposes inside the diagram)."""
    (i0s, j0s), (i1s, j1s) = X.n_states[0], X.n_states[1]
    which = "ket" if special_processing <= 1 else "bra"
    if no_result(X, which):
        return None
    res = X.ca0pq_U1pq
    K = None if i1s == j1s else X.KetCoeffs1
    if K is None and i1s != j1s:
        raise RuntimeError("u100: KetCoeffs are needed when bra and ket state counts differ")
    return res, K, which


def _delta_range(X, m):
    """(first ket state, count) that the Kronecker delta on diagram fragment m pairs with its bra states 0, 1, ...: the
    whole sector, or -- when the fragment's densities are a bra slab [lo, hi) -- ket states lo, lo + 1, ..."""
    n_bra, n_ket = X.n_states[m]
    lo = X.bra_offset[m]
    return (0, min(n_bra, n_ket)) if lo is None else (lo, n_bra)


def u100(X, special_processing=None):
    """delta(i1,j1) * sum_pq ca0[i0,j0,p,q] U[frag1, frag0, frag0][p,q]   (SU_2mer_0.py:26-96)"""
    (i0s, j0s), (i1s, j1s) = X.n_states[0], X.n_states[1]
    if special_processing is None:
        block = X.ca0pq_U1pq.host()
        result = numpy.zeros((i0s, i1s, j0s, j1s))
        lo, n_delta = _delta_range(X, 1)
        for i1 in range(n_delta):
            result[:, i1, :, lo + i1] = block
        return result
    if special_processing not in (0, 1, 2, 3):
        raise ValueError("special processing %r can not be handled" % (special_processing,))
    pieces = _u100_pieces(X, special_processing)
    if pieces is None:
        return []
    res, K, which = pieces
    if K is not None:       # the product with the ket coefficients runs on the device like every other contraction
        C = _contractor(X)
        res = C.contract(res, ["a", "z"] if which == "ket" else ["z", "a"], K, ["z", "b"], ["a", "b"])
        out = res.host()
    else:
        out = res.host() if which == "ket" else res.host().T
    return out.T if special_processing in (1, 3) else out


def _u100_accumulate(X, phase, out, offset, strides, special_processing=None):
    C = _contractor(X)
    if special_processing is None:
        lo, n_delta = _delta_range(X, 1)
        ones = C.dev.upload(numpy.ones((n_delta, 1)))
        C.contract(X.ca0pq_U1pq, ["i0", "j0"], DeviceTensor(ones, C.dev), ["d", "one"], ["i0", "j0", "d", "one"], alpha=phase,
                   out=out, out_offset=offset + lo * strides["j1"],
                   out_strides={"i0": strides["i0"], "j0": strides["j0"], "d": strides["i1"] + strides["j1"], "one": 0},
                   accumulate=True)
        return True
    pieces = _u100_pieces(X, special_processing)
    if pieces is None:
        return False
    res, K, which = pieces
    # (the odd special_processing values only say that the diagram's fragments are swapped with respect to the
    #  subsystem; here that is already expressed by which subsystem slot strides['i0'] / strides['i1'] point at)
    if which == "ket":
        a, b = "i0", "i1"
        if K is None:
            C.multi_contract([(res, [a, b])], [a, b], alpha=phase, out=out, out_offset=offset,
                             out_strides={a: strides[a], b: strides[b]}, accumulate=True)
        else:
            C.contract(res, [a, "z"], K, ["z", b], [a, b], alpha=phase, out=out, out_offset=offset,
                       out_strides={a: strides[a], b: strides[b]}, accumulate=True)
    else:
        a, b = "j0", "j1"
        if K is None:
            C.multi_contract([(res, [b, a])], [a, b], alpha=phase, out=out, out_offset=offset,
                             out_strides={a: strides[a], b: strides[b]}, accumulate=True)
        else:
            C.contract(res, ["z", a], K, ["z", b], [a, b], alpha=phase, out=out, out_offset=offset,
                       out_strides={a: strides[a], b: strides[b]}, accumulate=True)
    return True


u100.accumulate = _u100_accumulate


def identity(X):
    return 1
