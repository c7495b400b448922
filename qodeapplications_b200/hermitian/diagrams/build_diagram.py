"""Book-keeping between anonymised diagram fragments (0, 1, ...) and the actual fragments/charges of
a subsystem -- drop-in for hermitian-XRCC/diagrams/build_diagram.py (same two public names, same
call signatures, same attribute-name grammar on ``X``).

``build_diagram(contraction, Dchgs, permutations)`` -> ``get_permuted_diagrams(supersys_info,
subsys_chgs)`` -> ``[(do_contraction, permutation) | None, ...]``  (build_diagram.py:39-65).
``do_contraction(**kw)`` returns ``phase * contraction(X, **kw)`` as in the reference; in addition it
carries ``accumulate_into(out, offset, strides)``, which lets XR_term add the diagram straight into
the packed matrix on the device instead of materialising and transposing a block.
"""
import re

from ..tensor import DeviceTensor, as_host


def build_diagram(contraction, Dchgs, permutations):
    def get_permuted_diagrams(supersys_info, subsys_chgs):
        label = contraction.__name__
        def permuted_diagram(X, phase):
            def do_contraction(**args):
                supersys_info.timings.start()
                result = phase * contraction(X, **args)
                supersys_info.timings.record(label)
                return result
            def accumulate_into(out, offset, strides, scale=1.0, **args):
                supersys_info.timings.start()
                added = contraction.accumulate(X, phase * scale, out, offset, strides, **args)
                supersys_info.timings.record(label)
                return added
            do_contraction.accumulate_into = accumulate_into if hasattr(contraction, "accumulate") else None
            do_contraction.phase = phase
            return do_contraction
        if permutations is None:        # 0-mer / identity (S_diagrams.py:29-31)
            return [(permuted_diagram(None, +1), tuple())]
        permuted_diagrams = []
        for phase, permutation in permutations:
            X = frag_resolve(supersys_info, subsys_chgs, permutation)
            if all(X.Dchg[m] == Dchg for m, Dchg in enumerate(Dchgs)):
                permuted_diagrams.append((permuted_diagram(X, phase), permutation))
            else:
                permuted_diagrams.append(None)
        return permuted_diagrams
    return get_permuted_diagrams


class _indexable(object):
    def __init__(self, rule):
        self._rule = rule
    def __getitem__(self, indices):
        if not isinstance(indices, tuple):
            indices = (indices,)
        return self._rule(*indices)


class frag_resolve(object):
    """build_diagram.py:81-134.  subsys_chgs = ((absolute fragment, (bra charge, ket charge)), ...) in
    ascending fragment order; permutation = order in which the diagram uses them."""
    def __init__(self, supersys_info, subsys_chgs, permutation):
        self._info = supersys_info
        self._n_frag = len(subsys_chgs)
        self._storage = {}
        n_j = 0
        digits = "".join(str(i) for i in range(self._n_frag))
        for m, (frag_idx, (_, chg_j)) in enumerate(subsys_chgs):       # before permutation (build_diagram.py:87-92)
            n_j += supersys_info.densities[frag_idx]["n_elec"][chg_j]
            self._storage["n_j" + digits[:m + 1]] = n_j % 2
        self._subsys_chgs = [subsys_chgs[m] for m in permutation]
        self.permutation = tuple(permutation)
        self._frags = [f for f, _ in self._subsys_chgs]
        self._chgs = [c for _, c in self._subsys_chgs]
        dens = supersys_info.densities
        self._storage["Dchg"] = _indexable(lambda m: self._chgs[m][0] - self._chgs[m][1])
        self._storage["n_states"] = _indexable(lambda m: (dens[self._frags[m]]["n_states_bra"][self._chgs[m][0]],
                                                          dens[self._frags[m]]["n_states"][self._chgs[m][1]]))
        # first bra state held of diagram fragment m when its densities are a bra slab (hermitian/distributed.py), else None
        self._storage["bra_offset"] = _indexable(lambda m: dens[self._frags[m]]["bra_offset"].get(self._chgs[m][0], 0)
                                                 if "bra_offset" in dens[self._frags[m]] else None)

    # absolute fragment / charges of diagram fragment m
    def fragment(self, m):
        return self._frags[m]
    def charges(self, m):
        return self._chgs[m]

    def _integrals(self, kind):
        ints = self._info.integrals
        if kind == "s":
            return ints.S if hasattr(ints, "S") else ints
        return getattr(ints, kind.upper())

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        if attr[:3] == "n_j" or attr in ("Dchg", "n_states", "bra_offset"):
            return self._storage[attr]
        if attr in self._storage:
            return self._storage[attr]
        frag_indices = tuple(int(c) for c in attr if c.isdigit())
        template = re.sub(r"\d", "#", attr).replace("U#_", "U#")
        absolute = tuple(self._frags[m] for m in frag_indices)
        if "_" in template and template[0] not in "stuv":               # a precontraction
            cache = self._info.contract_cache.rho_S if "_S" in template else self._info.contract_cache.general
            entry = cache[template][absolute if len(absolute) > 1 else absolute[0]]
            try:
                value = entry[self._chgs[frag_indices[0]]]
            except KeyError:
                value = None
        elif template[0] in "stuv" and set(template[1:]) <= set("#_"):    # an integral block: s01, t10, u0_01, v0101
            block = self._integrals(template[0])[absolute if len(absolute) > 1 else absolute[0]]
            value = self._info.contract_cache.general.store.get(as_host(block))
        else:                                                            # a single-fragment density: ca0, a1, KetCoeffs1
            name = template[:-1]
            try:
                rho = self._info.densities[absolute[0]][name][self._chgs[frag_indices[0]]]
                value = self._info.contract_cache.general.store.get(rho)
            except KeyError:
                value = None
        self._storage[attr] = value
        return value
