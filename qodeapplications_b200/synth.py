"""Seeded synthetic inputs of the shapes BASELINE.md section 3 names.

Everything is FP64, C-contiguous, standard normal from ``numpy.random.default_rng``:

* integrals: one global h (F*n)^2 and V (F*n)^4 are drawn, symmetrised
  (h_pq = h_qp; V_pqrs = V_qpsr = V_rspq) and sliced into fragment blocks, so that the
  cross-block relation V[m1,m2,m3,m4][p,q,r,s] == V[m2,m1,m4,m3][q,p,s,r] holds
  (SURVEY.md App. F item 3); ``U[m]`` is a separate symmetric draw per nucleus fragment;
  ``S`` = identity + 0.05 * symmetric noise; ``nuc`` is a symmetric F x F matrix.
* densities: per fragment / operator string / (bra charge, ket charge) a dense
  ``[N_bra, N_ket, n, ...]`` array scaled by n^(-k/2), antisymmetric inside each run of
  like operators (frag-states/densities_frags.py:44-64) and conjugate-consistent
  (Be-states/densities.py:195-206: rho[rev_op][(cj,ci)][j,i,reversed] = rho[op][(ci,cj)][i,j,...]).

The same arrays serve both reference paths: hermitian-XRCC reads ``rho[op][(ci,cj)]`` as one
tensor, general-XRCC reads ``rho[op][(ci,cj)][i][j]`` (an ndarray row view gives exactly that);
general-XRCC's ``ccaa`` entries are the scalars sum_pqrs V[m,m,m,m][p,q,r,s]*ccaa[p,q,s,r]
(general-XRCC/build_density_tensors.py:129-132).
"""
import itertools
import numpy

# name -> (n_frag, n spin orbitals, {charge: n_states}); BASELINE.md section 3
CONFIGS = {
    "toy":  dict(n_frag=2, n_orb=6,  n_states={0: 3,   +1: 2,   -1: 2}),
    "toy3": dict(n_frag=3, n_orb=5,  n_states={0: 2,   +1: 2,   -1: 2}),
    "toy5": dict(n_frag=3, n_orb=4,  n_states={0: 2, +1: 2, -1: 2, +2: 1, -2: 1}),   # five charge states (general-XRCC/Be631g.py:73)
    "toy4": dict(n_frag=2, n_orb=5,  n_states={0: 3,   +1: 2,   -1: 2}),               # S-orders 3-4 (8-operator densities)
    "toyh": dict(n_frag=2, n_orb=[6, 4], n_states={0: 3, +1: 2, -1: 2}),             # fragments with different orbital counts
    "toyh3": dict(n_frag=3, n_orb=[5, 4, 3], n_states={0: 2, +1: 2, -1: 2}),         # ... and in one trimer
    "mid":  dict(n_frag=2, n_orb=8,  n_states={0: 5,   +1: 3,   -1: 4}),
    "herm49":  dict(n_frag=2, n_orb=18, n_states={0: 24, +1: 8,  -1: 17}),           # hermitian path at larger state counts
    "herm100": dict(n_frag=2, n_orb=18, n_states={0: 48, +1: 17, -1: 35}),           # (cfg4's 96:34:70 halved; order-0 densities 7 GB)
    "cfg1": dict(n_frag=2, n_orb=18, n_states={0: 11,  +1: 4,   -1: 8}),
    "cfg2": dict(n_frag=2, n_orb=18, n_states={0: 11,  +1: 4,   -1: 8}),
    "cfg3": dict(n_frag=3, n_orb=18, n_states={0: 11,  +1: 4,   -1: 8}),
    "cfg4": dict(n_frag=4, n_orb=18, n_states={0: 96,  +1: 34,  -1: 70}),
    "cfg5": dict(n_frag=2, n_orb=48, n_states={0: 478, +1: 174, -1: 348}),
}
SEEDS = {"toy": 11, "toy3": 13, "mid": 17, "toy5": 19, "toyh": 29, "toy4": 31, "cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5, "herm49": 41, "herm100": 43, "toyh3": 37}
N_ELEC_REF = 4

OPS_ORDER0 = ("a", "c", "aa", "cc", "ca", "caa", "cca", "ccaa")
OPS_ORDER1 = OPS_ORDER0 + ("caaa", "ccca", "ccaaa", "cccaa")
OPS_ORDER2 = OPS_ORDER1 + ("ccaaaa", "cccaaa", "ccccaa")
OPS_ORDER4 = OPS_ORDER2 + ("cccaaaa", "ccccaaa", "cccaaaaa", "ccccaaaa", "cccccaaa")
OPS_GENERAL = ("a", "c", "aa", "cc", "ca", "caa", "cca")      # + precontracted scalar "ccaa"


def op_dchg(op):
    """bra charge - ket charge selected by an operator string (precontract.py:66)."""
    return op.count("a") - op.count("c")


def conjugate_op(op):
    """Be-states/densities.py:198: reverse the string and swap c <-> a."""
    return op[::-1].translate(str.maketrans("ca", "ac"))


def _like_runs(op):
    runs, start = [], 0
    for pos in range(1, len(op) + 1):
        if pos == len(op) or op[pos] != op[start]:
            if pos - start > 1:
                runs.append(tuple(range(start, pos)))
            start = pos
    return runs


def _perm_parity(perm):
    perm = list(perm)
    parity = 1
    for i in range(len(perm)):
        while perm[i] != i:
            j = perm[i]
            perm[i], perm[j] = perm[j], perm[i]
            parity = -parity
    return parity


def antisymmetrize(t, op):
    """Antisymmetrise the orbital axes (axes 2..) of t inside every run of like operators."""
    for run in _like_runs(op):
        axes = [2 + a for a in run]
        acc = numpy.zeros_like(t)
        count = 0
        for perm in itertools.permutations(range(len(axes))):
            order = list(range(t.ndim))
            for slot, src in zip(axes, perm):
                order[slot] = axes[src]
            acc += _perm_parity(perm) * t.transpose(order)
            count += 1
        t = acc / count
    return numpy.ascontiguousarray(t)


def _reverse_conjugate(t):
    """[i, j, x1..xk] -> [j, i, xk..x1]"""
    k = t.ndim - 2
    return numpy.ascontiguousarray(t.transpose([1, 0] + [1 + k - a for a in range(k)]))


class blocked_integrals(object):
    """``X[m1, m2, ...]`` -> contiguous fragment block of one global tensor (cached)."""
    def __init__(self, full, n_orb, lead_blocks=None):
        self._full = full          # ndarray, or list of ndarrays when lead_blocks (U: one per nucleus fragment)
        self._n = n_orb            # int (homogeneous fragments) or list of per-fragment orbital counts
        self._lead = lead_blocks
        self._cache = {}
    def _slice(self, m):
        if isinstance(self._n, int):
            return slice(m * self._n, (m + 1) * self._n)
        lo = sum(self._n[:m])
        return slice(lo, lo + self._n[m])
    def __getitem__(self, frags):
        if not isinstance(frags, tuple):
            frags = (frags,)
        if frags not in self._cache:
            if self._lead:
                full, rest = self._full[frags[0]], frags[1:]
            else:
                full, rest = self._full, frags
            index = tuple(self._slice(m) for m in rest)
            self._cache[frags] = numpy.ascontiguousarray(full[index])
        return self._cache[frags]


class integral_set(object):
    def __init__(self, S, T, U, V, V_half=None, V_diff=None):
        self.S, self.T, self.U, self.V = S, T, U, V
        if V_half is not None:
            self.V_half = V_half
        if V_diff is not None:
            self.V_diff = V_diff


def make_integrals(n_frag, n_orb, rng, with_bior=False):
    if not isinstance(n_orb, int):
        return _make_integrals_hetero(list(n_orb), rng)
    dim = n_frag * n_orb
    def sym2():
        h = rng.standard_normal((dim, dim))
        return (h + h.T) / 2
    def sym4():
        V = rng.standard_normal((dim,) * 4)
        V = (V + V.transpose(1, 0, 3, 2)) / 2
        V = (V + V.transpose(2, 3, 0, 1)) / 2
        return V / n_orb
    noise = sym2()
    S = numpy.eye(dim)
    for m1 in range(n_frag):
        for m2 in range(n_frag):
            if m1 != m2:
                S[m1*n_orb:(m1+1)*n_orb, m2*n_orb:(m2+1)*n_orb] = 0.05 * noise[m1*n_orb:(m1+1)*n_orb, m2*n_orb:(m2+1)*n_orb]
    T = sym2()
    U = [sym2() for _ in range(n_frag)]
    V = sym4()
    symm = integral_set(S=blocked_integrals(S, n_orb), T=blocked_integrals(T, n_orb),
                        U=blocked_integrals(U, n_orb, lead_blocks=True), V=blocked_integrals(V, n_orb))
    nuc = rng.standard_normal((n_frag, n_frag))
    nuc = (nuc + nuc.T) / 2
    if not with_bior:
        return symm, nuc
    # "bior" set: an independent (non-symmetric) draw, plus V_half / V_diff = V_half - V
    # (hermitian-XRCC/get_ints.py:136-151 names; values are synthetic)
    Tb = T + 0.1 * rng.standard_normal((dim, dim))
    Ub = [u + 0.1 * rng.standard_normal((dim, dim)) for u in U]
    Vb = V + 0.1 * rng.standard_normal((dim,) * 4) / n_orb
    Vh = V + 0.05 * rng.standard_normal((dim,) * 4) / n_orb
    bior = integral_set(S=blocked_integrals(S, n_orb), T=blocked_integrals(Tb, n_orb),
                        U=blocked_integrals(Ub, n_orb, lead_blocks=True), V=blocked_integrals(Vb, n_orb),
                        V_half=blocked_integrals(Vh, n_orb), V_diff=blocked_integrals(Vh - V, n_orb))
    return symm, bior, nuc


def _make_integrals_hetero(n_orbs, rng):
    """fragments with different orbital counts (general-XRCC's C kernels take n_orb1, n_orb2 separately)"""
    dim = sum(n_orbs)
    sym2 = lambda: (lambda h: (h + h.T) / 2)(rng.standard_normal((dim, dim)))
    V = rng.standard_normal((dim,) * 4)
    V = (V + V.transpose(1, 0, 3, 2)) / 2
    V = (V + V.transpose(2, 3, 0, 1)) / 2 / max(n_orbs)
    S = numpy.eye(dim) + 0.05 * sym2() * (1 - numpy.eye(dim))
    symm = integral_set(S=blocked_integrals(S, n_orbs), T=blocked_integrals(sym2(), n_orbs),
                        U=blocked_integrals([sym2() for _ in n_orbs], n_orbs, lead_blocks=True), V=blocked_integrals(V, n_orbs))
    nuc = rng.standard_normal((len(n_orbs), len(n_orbs)))
    return symm, (nuc + nuc.T) / 2


def make_det_densities(n_orb, n_states_bra, n_states_ket, rng, ops=OPS_ORDER0, n_elec_ref=N_ELEC_REF):
    """Density dict of a fragment whose bras (or kets) are determinants (the "bra_det" / "ket_det" options of
    Be-states/densities.py:96-99,214-231): rectangular [N_bra, N_ket, n...] blocks, every operator string drawn
    independently (such densities are not symmetric, :205), plus KetCoeffs[(chg,chg)] = [N_states, N_dets]."""
    rho = {}
    for op in ops:
        k, d = len(op), op_dchg(op)
        rho[op] = {}
        for ci in n_states_bra:
            cj = ci - d
            if cj not in n_states_ket:
                continue
            t = rng.standard_normal((n_states_bra[ci], n_states_ket[cj]) + (n_orb,) * k) * n_orb ** (-k / 2)
            rho[op][(ci, cj)] = antisymmetrize(t, op)
    rho["n_elec"] = {chg: n_elec_ref - chg for chg in n_states_ket}
    rho["n_states"] = dict(n_states_ket)
    rho["n_states_bra"] = dict(n_states_bra)
    rho["KetCoeffs"] = {}
    for chg in n_states_ket:
        n_det, n_st = max(n_states_bra[chg], n_states_ket[chg]), min(n_states_bra[chg], n_states_ket[chg])
        rho["KetCoeffs"][(chg, chg)] = rng.standard_normal((n_st, n_det))
    return rho


def make_det_system(which, n_orb=6, n_states=None, n_dets=None, seed=23):
    """2-fragment toy for the bra_det / ket_det variants: fragment 0 has determinant bras (which="bra") or kets
    (which="ket"), fragment 1 is an ordinary fragment."""
    n_states = n_states or {0: 3, +1: 2, -1: 2}
    n_dets = n_dets or {0: 4, +1: 3, -1: 3}
    rng = numpy.random.default_rng(seed)
    symm, bior, nuc = make_integrals(2, n_orb, rng, with_bior=True)
    if which == "bra":
        rho0 = make_det_densities(n_orb, n_dets, n_states, rng)
    else:
        rho0 = make_det_densities(n_orb, n_states, n_dets, rng)
    rho1 = make_densities(n_orb, n_states, rng)
    return dict(n_frag=2, n_orb=n_orb, n_states=dict(n_states), charges=list(n_states), densities=[rho0, rho1],
                symm=symm, bior=bior, nuc=nuc)


def make_densities(n_orb, n_states, rng, ops=OPS_ORDER0, n_elec_ref=N_ELEC_REF, dtype=numpy.float64):
    """One fragment's density dict in the hermitian-path format (Be-states/densities.py:210-231)."""
    charges = list(n_states)
    rho = {}
    primary = [op for op in ops if op_dchg(op) > 0 or (op_dchg(op) == 0)]
    for op in primary:
        k = len(op)
        d = op_dchg(op)
        rho.setdefault(op, {})
        for ci in charges:
            cj = ci - d
            if cj not in n_states:
                continue
            t = rng.standard_normal((n_states[ci], n_states[cj]) + (n_orb,) * k) * n_orb ** (-k / 2)
            t = antisymmetrize(t, op)
            if d == 0:
                t = (t + _reverse_conjugate(t)) / 2        # rho[i,j,p..] = rho[j,i,..p]
            rho[op][(ci, cj)] = numpy.ascontiguousarray(t.astype(dtype, copy=False))
            if d != 0:
                cop = conjugate_op(op)
                if cop in ops:
                    rho.setdefault(cop, {})[(cj, ci)] = _reverse_conjugate(rho[op][(ci, cj)])
    rho["n_elec"] = {chg: n_elec_ref - chg for chg in charges}
    rho["n_states"] = dict(n_states)
    rho["n_states_bra"] = rho["n_states"]
    rho["KetCoeffs"] = {(chg, chg): numpy.eye(n_states[chg]) for chg in charges}
    return rho


class fragment(object):
    """What general-XRCC/build_H.py reads from a supersystem member (build_H.py:39-40) plus
    the state ordering test_H.py iterates (Be631g.py:81-85)."""
    def __init__(self, rho, n_elec_ref, state_indices):
        self.rho = rho
        self.n_elec_ref = n_elec_ref
        self.state_indices = state_indices


def _held_states(where, n_states, held, chg):
    """state indices of charge chg selected by a bra range: (lo, hi) matrix positions or {charge: (i_lo, i_hi)}"""
    if held is None:
        return list(range(n_states[chg]))
    if isinstance(held, dict):
        i_lo, i_hi = held.get(chg, (0, 0))
        return list(range(i_lo, i_hi))
    return [i for i in range(n_states[chg]) if held[0] <= where[(chg, i)] < held[1]]


def slab_fragment(frag, held, n_states):
    """The same fragment holding only the densities whose BRA state the range ``held`` selects -- (lo, hi) matrix
    positions or {charge: (i_lo, i_hi)} -- the per-rank input of a sharded build: build_matrix_elements(..., held={m: held})."""
    where = {state: p for p, state in enumerate(frag.state_indices)}
    keep = {chg: _held_states(where, n_states, held, chg) for chg in n_states}
    rho = {}
    for op, blocks in frag.rho.items():
        rho[op] = {}
        for (ci, cj), block in blocks.items():
            arr = numpy.asarray(block)
            rho[op][(ci, cj)] = numpy.ascontiguousarray(arr[keep[ci]]) if keep[ci] else arr[:0]
    out = fragment(rho, frag.n_elec_ref, frag.state_indices)
    out.n_states = dict(n_states)
    return out


def make_device_slab_fragments(n_frag, n_orb, n_states, held, torch_device, seed=0, ops=("a", "c", "aa", "cc", "ca", "caa", "cca"),
                               n_elec_ref=N_ELEC_REF):
    """Fragments whose densities are drawn ON THE DEVICE and only for the bra states held[m] selects (see slab_fragment)
    (BASELINE configs[4]: the full densities of 1000 states x 48 orbitals are ~450 GB per fragment and only ever exist
    sharded).  Block (ci,cj) of op is a CUDA tensor [n_held_ci, N_cj, n, ...]; the random stream of a bra state depends
    on (seed, fragment, op, sector, state) alone, so any sharding draws the same numbers."""
    import torch
    order = general_state_indices(n_states)
    where = {state: p for p, state in enumerate(order)}
    charges = list(n_states)
    frags = []
    for m in range(n_frag):
        rho = {}
        for o, op in enumerate(ops):
            k, dchg = len(op), op_dchg(op)
            rho[op] = {}
            for ci in charges:
                cj = ci - dchg
                if cj not in n_states:
                    continue
                rows = _held_states(where, n_states, held.get(m), ci)
                block = torch.empty((len(rows), n_states[cj]) + (n_orb,) * k, dtype=torch.float64, device=torch_device)
                gen = torch.Generator(device=torch_device)
                for r, i in enumerate(rows):
                    gen.manual_seed(((seed * 64 + m) * 16 + o) * 1000003 + (ci + 8) * 100003 + i)
                    # drawn in float32 (fast generator path), widened exactly: 110 GB per GPU at configs[4]
                    draw = torch.empty(block[r].shape, dtype=torch.float32, device=torch_device)
                    block[r].copy_(draw.normal_(0.0, n_orb ** (-k / 2), generator=gen))
                rho[op][(ci, cj)] = block
        f = fragment(rho, n_elec_ref, order)
        f.n_states = dict(n_states)
        frags.append(f)
    return frags


def general_state_indices(n_states, ref_state=(0, 0)):
    """general-XRCC/Be631g.py:81-85: reference state first, rest of its charge, then other charges."""
    ref_chg, ref_idx = ref_state
    order = [(ref_chg, ref_idx)]
    order += [(ref_chg, i) for i in range(n_states[ref_chg]) if i != ref_idx]
    for chg in n_states:
        if chg != ref_chg:
            order += [(chg, i) for i in range(n_states[chg])]
    return order


def general_rho_from_hermitian(rho, Vmmmm):
    """general-XRCC view of the same numbers: keeps a,c,aa,cc,ca,caa,cca as [N,N,n..] arrays
    (``arr[i][j]`` is the per-state-pair ndarray the reference indexes) and replaces ccaa by
    the scalar sum V[p,q,r,s]*ccaa[p,q,s,r] (build_density_tensors.py:129-132,
    H_contractions.c:61-76)."""
    out = {}
    for op in OPS_GENERAL:
        if op in rho:
            out[op] = dict(rho[op])
    if "ccaa" in rho:
        out["ccaa"] = {key: numpy.einsum("pqrs,ijpqsr->ij", Vmmmm, t) for key, t in rho["ccaa"].items()}
    return out


def make_system(name=None, n_frag=None, n_orb=None, n_states=None, seed=None, ops=OPS_ORDER0,
                with_bior=False, general_ccaa="contract"):
    """Returns a dict with densities (hermitian format), integrals, nuc, and general-path fragments.

    general_ccaa: "contract" builds the ccaa tensors and contracts them with V[m,m,m,m] for the
    general path; "random" draws the general path's ccaa scalars directly (large configs: the
    ccaa tensor would be N^2 n^4 doubles) and omits ccaa from the hermitian dict.
    """
    if name is not None:
        cfg = CONFIGS[name]
        n_frag = cfg["n_frag"] if n_frag is None else n_frag
        n_orb = cfg["n_orb"] if n_orb is None else n_orb
        n_states = cfg["n_states"] if n_states is None else n_states
        seed = SEEDS[name] if seed is None else seed
    rng = numpy.random.default_rng(seed)
    ints = make_integrals(n_frag, n_orb, rng, with_bior=with_bior)
    if with_bior:
        symm, bior, nuc = ints
    else:
        (symm, nuc), bior = ints, None
    if general_ccaa == "random":
        ops = tuple(op for op in ops if op != "ccaa")
    orb_of = (lambda m: n_orb) if isinstance(n_orb, int) else (lambda m: n_orb[m])
    dens = [make_densities(orb_of(m), n_states, rng, ops=ops) for m in range(n_frag)]
    frags = []
    for m in range(n_frag):
        grho = general_rho_from_hermitian(dens[m], symm.V[m, m, m, m])
        if general_ccaa == "random":
            grho["ccaa"] = {}
            for c in n_states:
                s = rng.standard_normal((n_states[c], n_states[c]))
                grho["ccaa"][(c, c)] = (s + s.T) / 2
        frags.append(fragment(grho, N_ELEC_REF, general_state_indices(n_states)))
    return dict(n_frag=n_frag, n_orb=n_orb, n_states=dict(n_states), charges=list(n_states),
                densities=dens, symm=symm, bior=bior, nuc=nuc, fragments=frags)
