"""Table-driven diagram definitions and the one routine that evaluates them on the GPU.

Every two-fragment diagram of the reference (hermitian-XRCC/diagrams/S*_2mer_*.py) has the form

    coefficient * (-1)**(X.n_j0 + shift) * raw( A(i0,j0, a-free...) @ B(i1,j1, b-free...) )

with A a density or rho x integral precontraction of diagram fragment 0, B the same for fragment 1,
and equal letters in the two free-index lists contracted.  A row of the table below is
``label: (coefficient, parity shift or None, A name, A free letters, B name, B free letters)``;
it is turned into a function with the reference's signature ``fn(X, contract_last=False) ->
ndarray[N_i0, N_i1, N_j0, N_j1]``, evaluated as ONE xr_gemm_scatter over K = the shared letters.
One-fragment diagrams are a single precontraction; ``u100`` carries a Kronecker delta over the
spectator fragment (SU_2mer_0.py:26-48); ``identity`` is the 0-mer (S_0mer_0.py:21-22).
File:line of the reference definition is given per row.
"""
import numpy

from ..tensor import DeviceTensor

# label: (coef, parity shift (None = no n_j0 phase), A, A free, B, B free)
TWO_FRAGMENT = {
    # ---- order 0 --------------------------------------------------------------------------------
    "t01":      (1, 0,    "c0p_Tp1", "q",            "a1", "q"),                    # ST_2mer_0.py:25-34
    "u001":     (1, 0,    "c0p_U0p1", "q",           "a1", "q"),                    # SU_2mer_0.py:98-107
    "u101":     (1, 0,    "c0p_U1p1", "q",           "a1", "q"),                    # SU_2mer_0.py:109-118
    "v0101":    (4, None, "ca0pr_Vp1r1", "qs",       "ca1", "qs"),                  # SV_2mer_0.py:25-34
    "v0001":    (2, 1,    "cca0pqr_Vpqr1", "s",      "a1", "s"),                    # SV_2mer_0.py:36-45
    "v0100":    (2, 0,    "caa0psr_Vp1rs", "q",      "c1", "q"),                    # SV_2mer_0.py:47-56
    "v0011":    (1, None, "cc0pq_Vpq11", "rs",       "aa1", "sr"),                  # SV_2mer_0.py:58-67
    # ---- order 1 --------------------------------------------------------------------------------
    "s01":      (1, 0,    "c0p_Sp1", "q",            "a1", "q"),                    # S_2mer_1.py:25-34
    "s01t10":   (-1, None, "ca0Xq_T1q", "tp",        "ca1Xu_S0u", "pt"),            # ST_2mer_1.py:25-36
    "s01t00":   (1, 1,    "cca0pXq_Tpq", "t",        "a1u_S0u", "t"),               # ST_2mer_1.py:38-49
    "s01t11":   (1, 1,    "c0t_St1", "u",            "caa1pXq_Tpq", "u"),           # ST_2mer_1.py:51-62
    "s01t01":   (1, None, "cc0pX_Tp1", "tq",         "aa1uX_S0u", "qt"),            # ST_2mer_1.py:64-75
    "s01u010":  (-1, None, "ca0Xq_U01q", "tp",       "ca1Xu_S0u", "pt"),            # SU_2mer_1.py
    "s01u000":  (1, 1,    "cca0pXq_U0pq", "t",       "a1u_S0u", "t"),
    "s01u011":  (1, 1,    "c0t_St1", "u",            "caa1pXq_U0pq", "u"),
    "s01u001":  (1, None, "cc0pX_U0p1", "tq",        "aa1uX_S0u", "qt"),
    "s01u110":  (-1, None, "ca0Xq_U11q", "tp",       "ca1Xu_S0u", "pt"),
    "s01u100":  (1, 1,    "cca0pXq_U1pq", "t",       "a1u_S0u", "t"),
    "s01u111":  (1, 1,    "c0t_St1", "u",            "caa1pXq_U1pq", "u"),
    "s01u101":  (1, None, "cc0pX_U1p1", "tq",        "aa1uX_S0u", "qt"),
    "s01v0100": (-2, None, "ccaa0pXsr_Vp1rs", "tq",  "ca1Xu_S0u", "qt"),            # SV_2mer_1.py:25-36
    "s01v1101": (2, None, "ca0tX_St1", "ru",         "ccaa1pqXs_Vpq0s", "ur"),      # SV_2mer_1.py:38-49
    "s01v0000": (1, 0,    "cccaa0pqXsr_Vpqrs", "t",  "a1u_S0u", "t"),               # SV_2mer_1.py:51-62
    "s01v0101": (4, 0,    "cca0pXr_Vp1r1", "tqs",    "caa1XuX_S0u", "qst"),         # SV_2mer_1.py:64-75
    "s01v1100": (1, 0,    "caa0Xsr_V11rs", "tpq",    "cca1XXu_S0u", "pqt"),         # SV_2mer_1.py:77-88
    "s01v1111": (1, 0,    "c0t_St1", "u",            "ccaaa1pqXsr_Vpqrs", "u"),     # SV_2mer_1.py:90-101
    "s01v0001": (2, None, "ccca0pqXr_Vpqr1", "ts",   "aa1uX_S0u", "st"),            # SV_2mer_1.py:103-114
    "s01v0111": (-2, None, "cc0Xt_St1", "pu",        "caaa1qXsr_V0qrs", "up"),      # SV_2mer_1.py:116-127
    "s01v0011": (1, 0,    "ccc0pqX_Vpq11", "trs",    "aaa1uXX_S0u", "srt"),         # SV_2mer_1.py:129-140
}

ONE_FRAGMENT = {
    "t00":   "ca0pq_Tpq",          # ST_1mer_0.py:24
    "u000":  "ca0pq_U0pq",         # SU_1mer_0.py:24
    "v0000": "ccaa0pqsr_Vpqrs",    # SV_1mer_0.py:24-31
}


def _require_plain(contract_last):
    if contract_last:
        raise NotImplementedError("contract_last (bra_det/ket_det gradient variants, diagram_hack.py:26-41) is not "
                                  "built yet; see DESIGN.md 'next'")


def _contractor(X):
    return X._info.contract_cache.general.contractor


def make_two_fragment(label):
    coef, shift, nameA, freeA, nameB, freeB = TWO_FRAGMENT[label]

    def operands(X):
        A, B = getattr(X, nameA), getattr(X, nameB)
        if A is None or B is None:
            raise RuntimeError("diagram %s: operand %s is not available for these charges" % (label, nameA if A is None else nameB))
        factor = float(coef) if shift is None else float(coef) * (-1.0) ** (X.n_j0 + shift)
        return A, ["i0", "j0"] + list(freeA), B, ["i1", "j1"] + list(freeB), factor

    def contraction(X, contract_last=False):
        _require_plain(contract_last)
        A, idxA, B, idxB, factor = operands(X)
        out = _contractor(X).contract(A, idxA, B, idxB, ["i0", "i1", "j0", "j1"], alpha=factor)
        return out.host()

    def accumulate(X, phase, out, offset, strides):
        """out[offset + i0*strides['i0'] + i1*strides['i1'] + j0*strides['j0'] + j1*strides['j1']] += phase * diagram"""
        A, idxA, B, idxB, factor = operands(X)
        _contractor(X).contract(A, idxA, B, idxB, ["i0", "i1", "j0", "j1"], alpha=phase * factor, out=out,
                                out_offset=offset, out_strides=strides, accumulate=True)

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


def u100(X, special_processing=None):
    """delta(i1,j1) * sum_pq ca0[i0,j0,p,q] U[frag1, frag0, frag0][p,q]   (SU_2mer_0.py:26-48)"""
    if special_processing is not None:
        raise NotImplementedError("u100 special_processing (bra_det/ket_det) is not built yet")
    (i0s, j0s), (i1s, j1s) = X.n_states[0], X.n_states[1]
    block = X.ca0pq_U1pq.host()
    result = numpy.zeros((i0s, i1s, j0s, j1s))
    for i1 in range(min(i1s, j1s)):
        result[:, i1, :, i1] = block
    return result


def _u100_accumulate(X, phase, out, offset, strides):
    (i0s, j0s), (i1s, j1s) = X.n_states[0], X.n_states[1]
    C = _contractor(X)
    n_delta = min(i1s, j1s)
    ones = C.dev.upload(numpy.ones((n_delta, 1)))
    C.contract(X.ca0pq_U1pq, ["i0", "j0"], DeviceTensor(ones, C.dev), ["d", "one"], ["i0", "j0", "d", "one"], alpha=phase,
               out=out, out_offset=offset,
               out_strides={"i0": strides["i0"], "j0": strides["j0"], "d": strides["i1"] + strides["j1"], "one": 0},
               accumulate=True)


u100.accumulate = _u100_accumulate


def identity(X):
    return 1
