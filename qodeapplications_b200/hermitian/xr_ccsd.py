"""The Hamiltonian build of hermitian-XRCC/mains/xr_ccsd.py:69-213 (everything between ``get_ints`` and ``excitonic.ccsd``)
for ANY S-order the diagram lists cover -- get_xr_H hard-codes orders 0-2 (get_xr_result.py:86-298), this caller does not:

    build_H(ints=(symm_ints, bior_ints, nuc_rep), dens=[rho0, rho1], order, monomer_charges=[0, +1, -1]) -> (H1, H2)

``order`` as on xr_ccsd.py's command line: "proper" (zeroth-order diagrams with biorthogonal integrals, no explicit S,
:107-115), "M=1" (monomer Hamiltonians only, :116-119) or an integer 0..4 (:120-148: S through that order, diagrams of
exactly that order with biorthogonal integrals, of order-1 with the V_diff integrals, everything lower with the symmetric
ones, ``S2inv @ S2H2`` minus the monomer terms).  H2 comes back in the (global state of fragment 0, global state of
fragment 1) ordering the script's reorder loop produces (:190-213).

Everything is accumulated in HBM by the GEMM epilogues, as in get_xr_result.get_xr_H (whose helpers are used here).
"""
from . import diagrammatic_expansion, XR_term
from . import diagram_lists as D
from .diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams
from .get_xr_result import checked_inverse
from .precontract import precontract
from .tensor import Contractor, DeviceStore, default_device
from .util import struct, timer


def build_H(ints, dens, order, monomer_charges=(0, +1, -1), device=None):
    symm_ints, bior_ints, nuc_rep = ints
    dev = device or default_device()
    store, contractor = DeviceStore(dev), Contractor(dev)
    diag_timer, precon_timer, matrix_timer = timer(), timer(), timer()
    contract_cache = precontract(dens, symm_ints.S, precon_timer, store=store, contractor=contractor)
    S = symm_ints.S
    cache = {}

    def blocks(kind):
        """lazily: S, ST/SU/SV with symm or bior integrals, SV with V_diff (xr_ccsd.py:72-81)"""
        if kind not in cache:
            family, which = kind.split("_") if "_" in kind else (kind, "symm")
            source = {"symm": symm_ints, "bior": bior_ints, "diff": bior_ints}[which]
            if family == "S":
                integrals, diagrams = S, S_diagrams
            elif family == "ST":
                integrals, diagrams = struct(S=S, T=source.T), ST_diagrams
            elif family == "SU":
                integrals, diagrams = struct(S=S, U=source.U), SU_diagrams
            else:
                integrals, diagrams = struct(S=S, V=source.V_diff if which == "diff" else source.V), SV_diagrams
            cache[kind] = diagrammatic_expansion.blocks(densities=dens, integrals=integrals, diagrams=diagrams, contract_cache=contract_cache,
                                                        timings=diag_timer, precon_timings=precon_timer)
        return cache[kind]

    monomer_charges = list(monomer_charges)
    all_dimer_charges = [(c0, c1) for c0 in monomer_charges for c1 in monomer_charges]
    lower = lambda lists, stop: sum((lists[o] for o in range(stop)), [])
    mono = {"ST": {1: D.ST1[0]}, "SU": {1: D.SU1[0]}, "SV": {1: D.SV1[0]}}         # the same at every order (:104)
    terms = []                       # (blocks, {fragment order: labels}) of S2H2
    if order == "proper":
        explicit_S, one_mer = False, "bior"
        terms += [(blocks("ST_bior"), {2: D.ST2[0]}), (blocks("SU_bior"), {2: D.SU2[0]}), (blocks("SV_bior"), {2: D.SV2[0]})]
    elif order == "M=1":
        explicit_S, one_mer = False, "symm"
    else:
        order = int(order)
        if order < 0:
            raise NotImplementedError("S-order %r" % (order,))
        explicit_S, one_mer = True, "symm"
        top_1 = lambda lists: lists[0] if order == 0 else []                        # higher-order monomer terms are zero (:124-126)
        terms += [(blocks("ST_bior"), {1: top_1(D.ST1), 2: D.ST2[order]}), (blocks("SU_bior"), {1: top_1(D.SU1), 2: D.SU2[order]}),
                  (blocks("SV_bior"), {1: top_1(D.SV1), 2: D.SV2[order]})]
        if order > 0:
            terms += [(blocks("ST_symm"), {1: D.ST1[0], 2: lower(D.ST2, order)}), (blocks("SU_symm"), {1: D.SU1[0], 2: lower(D.SU2, order)}),
                      (blocks("SV_diff"), {1: D.SV1[0] if order == 1 else [], 2: D.SV2[order - 1]})]
        if order > 1:
            terms += [(blocks("SV_symm"), {1: D.SV1[0], 2: lower(D.SV2, order - 1)})]

    H1 = []
    for m in (0, 1):
        M = None
        for family in ("ST", "SU", "SV"):
            M = XR_term.monomer_matrix(blocks(family + "_" + one_mer), mono[family], m, monomer_charges, matrix_timer, device_result=True, into=M)
        H1.append(M.host())

    def dimer_sum(selection, scale=1.0):
        out = None
        for op_blocks, active in selection:
            out = XR_term.dimer_matrix(op_blocks, active, (0, 1), all_dimer_charges, matrix_timer, ordering="final", device_result=True,
                                       into=out, scale=scale)
        return out

    if not terms:                    # "M=1": no dimer information at all (:116-119)
        dim = sum(dens[0]["n_states"][c] for c in monomer_charges) * sum(dens[1]["n_states"][c] for c in monomer_charges)
        return H1, dev.download(dev.zeros((dim, dim)))
    S2H2 = dimer_sum(terms)
    if not explicit_S:
        return H1, S2H2.host()
    S_active = {0: D.S0[0], 2: sum((D.S2[n] for n in range(1, order + 1)), [])}
    S2 = XR_term.dimer_matrix(blocks("S"), S_active, (0, 1), all_dimer_charges, matrix_timer, ordering="final", device_result=True)
    out = dimer_sum([(blocks(family + "_symm"), mono[family]) for family in ("ST", "SU", "SV")], scale=-1.0)
    contractor.contract(checked_inverse(S2, dev, contractor), ["a", "k"], S2H2, ["k", "b"], ["a", "b"], out=out, accumulate=True)
    return H1, out.host()
