"""rho x integral precontractions, cached -- drop-in for hermitian-XRCC/precontract.py.

Same label grammar (precontract.py:32-80):
    <rho type>#<rho index letters>_<S|T|U|V><integral index letters>
rho letters: a lower-case letter is contracted with the same letter of the integral, ``X`` is kept
free; integral letters: a letter is contracted (it lives on the rho's fragment), ``#`` is a free index
on another fragment of the diagram; for ``U`` the first ``#`` names the nucleus fragment.  Result index
order: (bra state, ket state, rho-free..., integral-free...)  (precontract.py:45-59,75).

``precontract(densities, integrals, timings)[label][frag indices][(chg_i, chg_j)]`` returns the cached
device tensor (None when the charges do not match the operator string, precontract.py:66-79).
Each one is a single xr_gemm_scatter call: rows = (state pair, rho-free indices), K = the contracted
letters in the rho's own order (the tiny integral block is permuted to match on the host; a density is
only permuted, by xr_permute_copy, when an ``X`` sits between contracted letters).
"""
import numpy

from .tensor import Contractor, DeviceStore, FactoredTensor, as_host, default_device


def parse_label(label):
    """-> (rho_type, rho_idx letters, int_type, int_idx letters, has_nucleus_slot)"""
    rho_label, int_label = label.split("_")
    rho_type, rho_idx = rho_label.split("#")
    int_type, int_idx = int_label[0], int_label[1:]
    nucleus = False
    if int_type == "U":
        nucleus = True
        int_idx = int_idx[1:]
    return rho_type, rho_idx, int_type, int_idx, nucleus


class _by_charges(object):
    def __init__(self, rule, charges):
        self._rule, self._charges, self._cache = rule, set(charges), {}
    def __getitem__(self, key):
        chg_i, chg_j = key
        if chg_i not in self._charges or chg_j not in self._charges:
            raise KeyError(key)
        if key not in self._cache:
            self._cache[key] = self._rule(chg_i, chg_j)
        return self._cache[key]


class _by_fragments(object):
    def __init__(self, rule, n_frag, n_indices):
        self._rule, self._n_frag, self._n_indices, self._cache = rule, n_frag, n_indices, {}
    def __getitem__(self, indices):
        if not isinstance(indices, tuple):
            indices = (indices,)
        if len(indices) != self._n_indices or any(not (0 <= m < self._n_frag) for m in indices):
            raise KeyError(indices)
        if indices not in self._cache:
            self._cache[indices] = self._rule(*indices)
        return self._cache[indices]


class precontract(object):
    def __init__(self, densities, integrals, timings=None, store=None, contractor=None):
        self._densities = densities
        self._integrals = integrals
        self._timings = timings
        dev = (store.dev if store is not None else None) or default_device()
        self.store = store or DeviceStore(dev)
        self.contractor = contractor or Contractor(dev)
        self._labels = {}
        self._int_cache = {}

    def __getitem__(self, label):
        if label not in self._labels:
            self._labels[label] = self._mother_rule(label)
        return self._labels[label]

    def _integral_block(self, int_type, blocks):
        ints = self._integrals
        if int_type == "S":
            source = ints[blocks] if not hasattr(ints, "S") else ints.S[blocks]
        else:
            source = getattr(ints, int_type)[blocks]
        return as_host(source)

    def _mother_rule(self, label):
        rho_type, rho_idx, int_type, int_idx, nucleus = parse_label(label)
        n_indices = label.count("#")
        n_frag = len(self._densities)
        Dchg = rho_type.count("a") - rho_type.count("c")
        contracted = [c for c in rho_idx if c != "X"]
        if sorted(contracted) != sorted(c for c in int_idx if c != "#"):
            raise ValueError("precontraction label %r: contracted letters do not pair up" % label)

        def contract_rho_int(*indices):
            blocks, slot = [], 0
            if nucleus:
                slot += 1
                blocks.append(indices[slot])
            for c in int_idx:
                if c == "#":
                    slot += 1
                    blocks.append(indices[slot])
                else:
                    blocks.append(indices[0])
            blocks = tuple(blocks)
            rho_m = self._densities[indices[0]]
            # integral as [free '#' axes in order..., contracted letters in the rho's order]
            key = (int_type, blocks, int_idx, tuple(contracted))
            if key not in self._int_cache:
                block = self._integral_block(int_type, blocks)
                axes_free = [a for a, c in enumerate(int_idx) if c == "#"]
                axes_k = [int_idx.index(c) for c in contracted]
                self._int_cache[key] = self.store.get(numpy.ascontiguousarray(block.transpose(axes_free + axes_k)))
            Wt = self._int_cache[key]
            n_free_int = sum(1 for c in int_idx if c == "#")

            def contract_rho_int_m(chg_i, chg_j):
                if chg_i - chg_j != Dchg:
                    return None
                try:
                    rho = rho_m[rho_type][chg_i, chg_j]
                except (KeyError, TypeError):
                    raise RuntimeError("missing density %s for charges %r" % (rho_type, (chg_i, chg_j)))
                if self._timings is not None:
                    self._timings.start()
                R = self.store.get(rho)
                idxR = ["i", "j"] + ["x%d" % k if c == "X" else c for k, c in enumerate(rho_idx)]
                idxW = ["f%d" % k for k in range(n_free_int)] + contracted
                free = [l for l in idxR if l not in contracted] + idxW[:n_free_int]
                if isinstance(R, FactoredTensor):
                    # a lazy density stays lazy (as under tensornet): the integral joins the factors of every term and is
                    # contracted inside the diagram that uses this operand
                    result = R.appended(idxR, (Wt, idxW), free)
                else:
                    result = self.contractor.contract(R, idxR, Wt, idxW, free)
                if self._timings is not None:
                    self._timings.record(label)
                return result
            return _by_charges(contract_rho_int_m, rho_m["n_states"].keys())

        if label.count("_") != 1:
            raise NotImplementedError(label)
        return _by_fragments(contract_rho_int, n_frag, n_indices)
