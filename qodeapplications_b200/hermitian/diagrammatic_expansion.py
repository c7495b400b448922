"""Lazy nested ``blocks[subsystem][charges][label]`` -- drop-in for
hermitian-XRCC/diagrammatic_expansion.py (same class name and constructor signature, :111-124).

``blocks[...][...][label]`` returns what the reference returns: the ndarray block summed over the
catalog's permutations, each transposed back by ``[perm] + [n + perm]`` (:27-59, :76-93), or None.
The extra method ``_charges.accumulate(label, out, offset, slot_strides)`` adds the same block in
place into a device matrix: the transposes become offset tables of the GEMM epilogue and nothing is
materialised (this is the path XR_term uses).
"""
import numpy

from .precontract import precontract
from .util import struct


def _build_block(diagram_term, permutation, bra_det, ket_det, label):
    """diagrammatic_expansion.py:27-59 of the reference"""
    frag_order = len(permutation)
    if frag_order == 0:
        return diagram_term()
    reorder = [m for m in permutation] + [frag_order + m for m in permutation]
    if bra_det and not ket_det:
        if label == "u100":
            return diagram_term(special_processing=reorder[0])       # transposition is handled inside the diagram
        if frag_order == 2:
            result = diagram_term(contract_last="ket")
            return numpy.asarray(result).transpose(reorder[:2]) if len(result) else []
    elif ket_det and not bra_det:
        if label == "u100":
            return diagram_term(special_processing=reorder[2])
        if frag_order == 2:
            result = diagram_term(contract_last="bra")
            return numpy.asarray(result).transpose([i - 2 for i in reorder[:2]]) if len(result) else []
    return numpy.asarray(diagram_term()).transpose(reorder)


class _charges(object):
    def __init__(self, supersys_info, subsystem, charges, diagrams, bra_det, ket_det):
        self._supersys_info, self._subsystem, self._charges = supersys_info, subsystem, charges
        self._diagrams, self._bra_det, self._ket_det = diagrams, bra_det, ket_det
        self._results = {}
        self._terms = {}

    def _get_terms(self, label):
        if label not in self._terms:
            frag_order = len(self._subsystem)
            try:
                entry = self._diagrams.catalog[frag_order][label]
            except KeyError:
                raise NotImplementedError("diagram '{}' not implemented for {} bodies".format(label, frag_order))
            self._terms[label] = entry(self._supersys_info, tuple(zip(self._subsystem, self._charges)))
        return self._terms[label]

    def __getitem__(self, label):
        if label not in self._results:
            self._results[label] = None
            for term_permutation in self._get_terms(label):
                if term_permutation is not None:
                    term, permutation = term_permutation
                    result = _build_block(term, permutation, self._bra_det, self._ket_det, label)
                    if (self._bra_det or self._ket_det) and numpy.ndim(result) > 0 and len(result) == 0:
                        continue
                    if self._results[label] is None:
                        self._results[label] = result
                    else:
                        self._results[label] = self._results[label] + result
        return self._results[label]

    def accumulate(self, label, out, offset, slot_strides, scale=1.0):
        """out += this block, where slot_strides[("i"|"j", subsystem slot)] is the element stride of that
        slot's bra/ket state index inside `out` (plus optional "delta"/"n_delta" for spectator deltas).  With
        bra_det (ket_det) only the bra (ket) strides are used: the other pair is traced (diagrammatic_expansion.py:33-56).
        Returns False when no permutation is charge-allowed (or no trace exists)."""
        det = "ket" if (self._bra_det and not self._ket_det) else ("bra" if (self._ket_det and not self._bra_det) else False)
        added = False
        for term_permutation in self._get_terms(label):
            if term_permutation is None:
                continue
            term, permutation = term_permutation
            if term.accumulate_into is None:
                raise NotImplementedError("diagram '%s' has no in-place form" % label)
            strides = {}
            for d, slot in enumerate(permutation):
                for side in ("i", "j"):
                    if (side, slot) in slot_strides:
                        strides["%s%d" % (side, d)] = slot_strides[(side, slot)]
            for key in ("delta", "n_delta"):
                if key in slot_strides:
                    strides[key] = slot_strides[key]
            kw = {}
            if det and len(permutation) == 2:
                if label == "u100":
                    kw["special_processing"] = permutation[0] if det == "ket" else 2 + permutation[0]
                else:
                    kw["contract_last"] = det
            if term.accumulate_into(out, offset, strides, scale, **kw) is not False:
                added = True
        return added


class _subsystem(object):
    def __init__(self, supersys_info, subsystem, diagrams, bra_det, ket_det):
        self._supersys_info, self._subsystem, self._diagrams = supersys_info, subsystem, diagrams
        self._bra_det, self._ket_det = bra_det, ket_det
        self._items = {}
    def __getitem__(self, charges):
        if charges is None:
            charges = tuple()
        charges = tuple(charges)
        if charges not in self._items:
            self._items[charges] = _charges(self._supersys_info, self._subsystem, charges, self._diagrams,
                                            self._bra_det, self._ket_det)
        return self._items[charges]


class blocks(object):
    def __init__(self, densities, integrals, diagrams, contract_cache, timings, precon_timings, bra_det=False, ket_det=False):
        general = precontract(densities, integrals, precon_timings, store=contract_cache.store,
                              contractor=contract_cache.contractor)
        contract_cache = struct(rho_S=contract_cache, general=general)
        self._supersys_info = struct(densities=densities, integrals=integrals, contract_cache=contract_cache, timings=timings)
        self._diagrams = diagrams
        self._bra_det, self._ket_det = bra_det, ket_det
        self._items = {}
        self.densities = densities
        self.dev = general.store.dev
    def __getitem__(self, subsystem):
        if subsystem is None:
            subsystem = tuple()
        subsystem = tuple(subsystem)
        if subsystem not in self._items:
            self._items[subsystem] = _subsystem(self._supersys_info, subsystem, self._diagrams, self._bra_det, self._ket_det)
        return self._items[subsystem]
