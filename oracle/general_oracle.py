"""CPU oracle for the general-XRCC Hamiltonian build.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this;
the product path (qodeapplications_b200/) never does.

Parity status: PINNED against the reference itself -- tests/golden/general_*.npz were
produced by running the reference's unmodified general-XRCC/build_H.py + H_contractions.c
in the build container (oracle/gen_golden.py); tests/test_oracle.py checks both restatements
below against them.  (The reference has no tests or golden vectors of its own for this path,
SURVEY.md 8(c).)

Two restatements:

* ``element_oracle``  -- a port of build_H.build_matrix_elements' per-element control flow
  (general-XRCC/build_H.py:42-188) that calls a C library with the H_contractions ABI: either
  oracle/_ref/libH_contractions_ref.so (the reference's own C, compiled where it lies) or
  oracle/libxr_oracle.so (our plain-C restatement).  This is what the CPU baseline times.
* ``block_*``         -- NumPy einsum block form of the same formulas,
  whole charge-blocks at a time; used to check the CUDA path at sizes where per-element
  evaluation would take hours.
"""
import ctypes
import itertools
import os
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
_c_double_p = ctypes.POINTER(ctypes.c_double)

_SIGNATURES = {           # name -> number of leading PyInt arguments, number of Double* arguments
    "monomer": (1, 4), "monomer_1e": (1, 2), "monomer_2e": (1, 2), "monomer_extPot": (1, 2),
    "dimer_2min2pls": (2, 3), "dimer_1min1pls_1e": (2, 3), "dimer_1min1pls_2e": (2, 6), "dimer_ExEx": (2, 3),
    "trimer_2min1pls1pls": (3, 4), "trimer_2pls1min1min": (3, 4), "trimer_Ex1min1pls": (3, 4),
}


class c_contractions(object):
    """ctypes view of a library exporting the 11 H_contractions.c symbols
    (general-XRCC/H_contractions.c:22,48,61,80,95,116,129,163,184,208,232)."""
    def __init__(self, which="ref"):
        path = {"ref": os.path.join(HERE, "_ref", "libH_contractions_ref.so"),
                "port": os.path.join(HERE, "libxr_oracle.so")}[which]
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle`)")
        self.kind = "reference" if which == "ref" else "port"
        self._lib = ctypes.CDLL(path)
        for name, (n_int, n_ptr) in _SIGNATURES.items():
            fn = getattr(self._lib, name)
            fn.restype = ctypes.c_double
            fn.argtypes = [ctypes.c_int64] * n_int + [_c_double_p] * n_ptr
            setattr(self, name, self._wrap(fn, n_int))
    @staticmethod
    def _wrap(fn, n_int):
        def call(*args):
            conv = [int(a) for a in args[:n_int]]
            keep = []
            for a in args[n_int:]:
                a = numpy.ascontiguousarray(a, dtype=numpy.float64)
                keep.append(a)
                conv.append(a.ctypes.data_as(_c_double_p))
            return fn(*conv)
        return call


def parity(n_elec_ref, chg):
    """build_H.py:68,113-114: (-1)**((n_elec[m]-chg) % 2)"""
    return -1 if (n_elec_ref - chg) % 2 else +1


class element_oracle(object):
    """Per-element evaluation with the control flow of build_H.py:42-188."""
    def __init__(self, fragments, integrals, nuc_repulsion, contract):
        self.rho = [f.rho for f in fragments]
        self.n_elec = [f.n_elec_ref for f in fragments]
        self.T, self.U, self.V = integrals.T, integrals.U, integrals.V
        self.nuc = nuc_repulsion
        self.c = contract
        self._h = {}
    def _R(self, m, op, ci, cj, i, j):
        return self.rho[m][op][(ci, cj)][i][j]
    def monomer(self, m, I, J):                                   # build_H.py:42-55
        (ci, i), (cj, j) = I, J
        if ci != cj:
            return 0
        n = self.T[m, m].shape[0]
        key = ("m", m)
        if key not in self._h:
            self._h[key] = self.T[m, m] + self.U[m, m, m]
        E = self.nuc[m, m] if i == j else 0
        return E + self._R(m, "ccaa", ci, cj, i, j) + self.c.monomer_1e(n, self._R(m, "ca", ci, cj, i, j), self._h[key])
    def dimer(self, frags, I, J):                                 # build_H.py:56-102
        m1, m2 = frags
        (c1, i1), (c2, i2) = I
        (d1, j1), (d2, j2) = J
        if c1 + c2 != d1 + d2 or abs(c1 - d1) > 2:
            return 0
        n1, n2 = self.T[m1, m1].shape[0], self.T[m2, m2].shape[0]
        s2 = parity(self.n_elec[m2], c2)
        delta = c1 - d1
        if delta == -2:
            return self.c.dimer_2min2pls(n1, n2, self._R(m1, "cc", c1, d1, i1, j1), self._R(m2, "aa", c2, d2, i2, j2),
                                         self.V[m1, m1, m2, m2])
        if delta == +2:
            return self.c.dimer_2min2pls(n2, n1, self._R(m2, "cc", c2, d2, i2, j2), self._R(m1, "aa", c1, d1, i1, j1),
                                         self.V[m2, m2, m1, m1])
        if delta == 0:
            R1, R2 = self._R(m1, "ca", c1, d1, i1, j1), self._R(m2, "ca", c2, d2, i2, j2)
            val = self.c.dimer_ExEx(n1, n2, R1, R2, self.V[m1, m2, m1, m2])
            if i1 == j1:
                val += self.c.monomer_extPot(n2, R2, self.U[m1, m2, m2])
            if i2 == j2:
                val += self.c.monomer_extPot(n1, R1, self.U[m2, m1, m1])
            if i1 == j1 and i2 == j2:
                val += self.nuc[m1, m2]
            return val
        # |delta| == 1: x creates (delta_x = -1), y annihilates
        if delta == -1:
            x, y, sign = m1, m2, +s2
            X, Y = (c1, d1, i1, j1), (c2, d2, i2, j2)
        else:
            x, y, sign = m2, m1, -s2
            X, Y = (c2, d2, i2, j2), (c1, d1, i1, j1)
        nx, ny = self.T[x, x].shape[0], self.T[y, y].shape[0]
        h = self.T[x, y] + self.U[m1, x, y] + self.U[m2, x, y]     # re-added per element, as the reference does
        Rc, Rcca = self._R(x, "c", *X), self._R(x, "cca", *X)
        Ra, Rcaa = self._R(y, "a", *Y), self._R(y, "caa", *Y)
        val = self.c.dimer_1min1pls_1e(nx, ny, Rc, Ra, h)
        val += self.c.dimer_1min1pls_2e(nx, ny, Rc, Rcca, Ra, Rcaa, self.V[x, x, x, y], self.V[x, y, y, y])
        return sign * val
    def trimer(self, frags, I, J):                                # build_H.py:103-188
        m = frags
        chg_i = [I[k][0] for k in range(3)]
        chg_j = [J[k][0] for k in range(3)]
        idx_i = [I[k][1] for k in range(3)]
        idx_j = [J[k][1] for k in range(3)]
        d = [chg_i[k] - chg_j[k] for k in range(3)]
        if sorted(d) not in ([-2, 1, 1], [-1, 0, 1], [-1, -1, 2]):
            return 0
        n = [self.T[mk, mk].shape[0] for mk in m]
        s2, s3 = parity(self.n_elec[m[1]], chg_i[1]), parity(self.n_elec[m[2]], chg_i[2])
        base = {0: s3, 1: s2 * s3, 2: s2}
        R = lambda k, op: self._R(m[k], op, chg_i[k], chg_j[k], idx_i[k], idx_j[k])
        for k in range(3):
            o0, o1 = [o for o in range(3) if o != k]
            if d[k] == -2:
                return base[k] * self.c.trimer_2min1pls1pls(n[k], n[o0], n[o1], R(k, "cc"), R(o0, "a"), R(o1, "a"),
                                                            self.V[m[k], m[k], m[o0], m[o1]])
        for k in range(3):
            o0, o1 = [o for o in range(3) if o != k]
            if d[k] == +2:
                return base[k] * self.c.trimer_2pls1min1min(n[k], n[o0], n[o1], R(k, "aa"), R(o0, "c"), R(o1, "c"),
                                                            self.V[m[o0], m[o1], m[k], m[k]])
        k = d.index(0)
        c, a = d.index(-1), d.index(+1)
        sign = base[k] if c < a else -base[k]
        val = self.c.trimer_Ex1min1pls(n[k], n[c], n[a], R(k, "ca"), R(c, "c"), R(a, "a"), self.V[m[k], m[c], m[k], m[a]])
        if idx_i[k] == idx_j[k]:
            val += self.c.dimer_1min1pls_1e(n[c], n[a], R(c, "c"), R(a, "a"), self.U[m[k], m[c], m[a]])
        return sign * val


# ------------------------------------------------------------------------------ block form

def _positions(state_indices):
    pos = {}
    for p, (chg, idx) in enumerate(state_indices):
        pos[(chg, idx)] = p
    return pos


def _sector_positions(state_indices, n_states):
    """{chg: int array of matrix positions of states 0..N_chg-1}"""
    pos = _positions(state_indices)
    return {chg: numpy.array([pos[(chg, i)] for i in range(n)], dtype=numpy.int64) for chg, n in n_states.items()}


def _n_states(fragment):
    n = {}
    for chg, idx in fragment.state_indices:
        n[chg] = max(n.get(chg, 0), idx + 1)
    return n


def block_monomer(fragments, integrals, nuc, m):
    """H1[m] dense in state_indices order (build_H.py:42-55)."""
    f = fragments[m]
    ns = _n_states(f)
    sp = _sector_positions(f.state_indices, ns)
    dim = len(f.state_indices)
    H = numpy.zeros((dim, dim))
    h = integrals.T[m, m] + integrals.U[m, m, m]
    for c in ns:
        blk = numpy.asarray(f.rho["ccaa"][(c, c)]) + numpy.einsum("ijpq,pq->ij", f.rho["ca"][(c, c)], h)
        blk = blk + nuc[m, m] * numpy.eye(ns[c])
        H[numpy.ix_(sp[c], sp[c])] = blk
    return H


def dimer_charge_block(fragments, integrals, nuc, m1, m2, bra, ket):
    """[N_i1, N_i2, N_j1, N_j2] block for bra charges (c1,c2), ket charges (d1,d2), or None if
    forbidden (build_H.py:65)."""
    (c1, c2), (d1, d2) = bra, ket
    if c1 + c2 != d1 + d2 or abs(c1 - d1) > 2:
        return None
    r1, r2 = fragments[m1].rho, fragments[m2].rho
    T, U, V = integrals.T, integrals.U, integrals.V
    delta = c1 - d1
    if delta == -2:
        return numpy.einsum("pqrs,ijpq,klsr->ikjl", V[m1, m1, m2, m2], r1["cc"][(c1, d1)], r2["aa"][(c2, d2)], optimize=True)
    if delta == +2:
        return numpy.einsum("pqrs,klpq,ijsr->ikjl", V[m2, m2, m1, m1], r2["cc"][(c2, d2)], r1["aa"][(c1, d1)], optimize=True)
    if delta == 0:
        ca1, ca2 = r1["ca"][(c1, d1)], r2["ca"][(c2, d2)]
        blk = 4 * numpy.einsum("pqrs,ijpr,klqs->ikjl", V[m1, m2, m1, m2], ca1, ca2, optimize=True)
        e1, e2 = numpy.eye(ca1.shape[0]), numpy.eye(ca2.shape[0])
        blk += numpy.einsum("ij,kl->ikjl", e1, numpy.einsum("klqs,qs->kl", ca2, U[m1, m2, m2]))
        blk += numpy.einsum("ij,kl->ikjl", numpy.einsum("ijpr,pr->ij", ca1, U[m2, m1, m1]), e2)
        blk += nuc[m1, m2] * numpy.einsum("ij,kl->ikjl", e1, e2)
        return blk
    s2 = parity(fragments[m2].n_elec_ref, c2)
    if delta == -1:
        x, y, sign = m1, m2, +s2
        Rc, Rcca, Ra, Rcaa = r1["c"][(c1, d1)], r1["cca"][(c1, d1)], r2["a"][(c2, d2)], r2["caa"][(c2, d2)]
        out = "ikjl"
    else:
        x, y, sign = m2, m1, -s2
        Rc, Rcca, Ra, Rcaa = r2["c"][(c2, d2)], r2["cca"][(c2, d2)], r1["a"][(c1, d1)], r1["caa"][(c1, d1)]
        out = "kilj"
    h = T[x, y] + U[m1, x, y] + U[m2, x, y]
    blk = numpy.einsum("pq,ijp,klq->" + out, h, Rc, Ra, optimize=True)
    blk += 2 * numpy.einsum("pqrs,ijqpr,kls->" + out, V[x, x, x, y], Rcca, Ra, optimize=True)
    blk += 2 * numpy.einsum("pqrs,ijp,klqsr->" + out, V[x, y, y, y], Rc, Rcaa, optimize=True)
    return sign * blk


def block_dimer(fragments, integrals, nuc, m1, m2):
    """H2[m1][m2] dense, ordering of test_H.py:101-107."""
    f1, f2 = fragments[m1], fragments[m2]
    ns1, ns2 = _n_states(f1), _n_states(f2)
    sp1, sp2 = _sector_positions(f1.state_indices, ns1), _sector_positions(f2.state_indices, ns2)
    dim1, dim2 = len(f1.state_indices), len(f2.state_indices)
    H = numpy.zeros((dim1, dim2, dim1, dim2))
    for c1, c2, d1, d2 in itertools.product(ns1, ns2, ns1, ns2):
        blk = dimer_charge_block(fragments, integrals, nuc, m1, m2, (c1, c2), (d1, d2))
        if blk is not None:
            H[numpy.ix_(sp1[c1], sp2[c2], sp1[d1], sp2[d2])] = blk
    return H.reshape(dim1 * dim2, dim1 * dim2)


def trimer_charge_block(fragments, integrals, ms, bra, ket):
    """[N_i1,N_i2,N_i3,N_j1,N_j2,N_j3] block or None (build_H.py:103-188)."""
    d = [bra[k] - ket[k] for k in range(3)]
    if sorted(d) not in ([-2, 1, 1], [-1, 0, 1], [-1, -1, 2]):
        return None
    rho = [fragments[m].rho for m in ms]
    U, V = integrals.U, integrals.V
    s2, s3 = parity(fragments[ms[1]].n_elec_ref, bra[1]), parity(fragments[ms[2]].n_elec_ref, bra[2])
    base = {0: s3, 1: s2 * s3, 2: s2}
    R = lambda k, op: rho[k][op][(bra[k], ket[k])]
    bra_l, ket_l = "abc", "xyz"
    out = "abcxyz"
    def sub(k):
        return bra_l[k] + ket_l[k]
    if -2 in d or 2 in d:
        k = d.index(-2) if -2 in d else d.index(2)
        o0, o1 = [o for o in range(3) if o != k]
        if d[k] == -2:
            expr = "pqrs,%sqp,%sr,%ss->%s" % (sub(k), sub(o0), sub(o1), out)
            val = numpy.einsum(expr, V[ms[k], ms[k], ms[o0], ms[o1]], R(k, "cc"), R(o0, "a"), R(o1, "a"), optimize=True)
        else:
            expr = "rspq,%sr,%ss,%sqp->%s" % (sub(o0), sub(o1), sub(k), out)
            val = numpy.einsum(expr, V[ms[o0], ms[o1], ms[k], ms[k]], R(o0, "c"), R(o1, "c"), R(k, "aa"), optimize=True)
        return base[k] * 2 * val
    k, c, a = d.index(0), d.index(-1), d.index(+1)
    sign = base[k] if c < a else -base[k]
    expr = "prqs,%spq,%sr,%ss->%s" % (sub(k), sub(c), sub(a), out)
    val = 4 * numpy.einsum(expr, V[ms[k], ms[c], ms[k], ms[a]], R(k, "ca"), R(c, "c"), R(a, "a"), optimize=True)
    eye = numpy.eye(R(k, "ca").shape[0])
    expr = "%s,rs,%sr,%ss->%s" % (sub(k), sub(c), sub(a), out)
    val += numpy.einsum(expr, eye, U[ms[k], ms[c], ms[a]], R(c, "c"), R(a, "a"), optimize=True)
    return sign * val


def block_trimer(fragments, integrals, ms):
    """H3[m1][m2][m3] dense, ordering of test_H.py:113-126."""
    fs = [fragments[m] for m in ms]
    ns = [_n_states(f) for f in fs]
    sp = [_sector_positions(f.state_indices, n) for f, n in zip(fs, ns)]
    dims = [len(f.state_indices) for f in fs]
    H = numpy.zeros(dims + dims)
    for bra in itertools.product(*ns):
        for ket in itertools.product(*ns):
            blk = trimer_charge_block(fragments, integrals, ms, bra, ket)
            if blk is not None:
                H[numpy.ix_(sp[0][bra[0]], sp[1][bra[1]], sp[2][bra[2]], sp[0][ket[0]], sp[1][ket[1]], sp[2][ket[2]])] = blk
    D = dims[0] * dims[1] * dims[2]
    return H.reshape(D, D)


def trimer_class_moments(W, beta, gamma):
    """Exact sum and sum of squares of T[a,b,c] = sum_rs W[a,r,s] beta[b,r] gamma[c,s] from Gram
    matrices (size independent: O((Pa+Pb+Pc) n^4)); the full-size property check for streamed
    trimer tiles."""
    n = beta.shape[1]
    total = numpy.einsum("rs,r,s->", W.sum(axis=0), beta.sum(axis=0), gamma.sum(axis=0))
    Wf = W.reshape(W.shape[0], n * n)
    GW = (Wf.T @ Wf).reshape(n, n, n, n)            # [r,s,r',s']
    Gb, Gg = beta.T @ beta, gamma.T @ gamma
    sumsq = numpy.einsum("rsuv,ru,sv->", GW, Gb, Gg, optimize=True)
    return float(total), float(sumsq)


# ------------------------------------------------------------------------------- consumer: supersystem matrix

def supersystem_matrix(dims, H):
    """NumPy restatement of general-XRCC/hamiltonian.py:21-84 (braket_loops; hermitian-XRCC/hamiltonian.py when H has no
    trimer member), allowing unequal fragment dimensions.  Every block is embedded with Kronecker deltas on the spectator
    fragments; trimer couplings are read only where at least two of their fragments change state (hamiltonian.py:44-56)."""
    import itertools
    F = len(dims)
    D = int(numpy.prod(dims))
    Hmat = numpy.zeros(tuple(dims) + tuple(dims))
    eyes = [numpy.eye(d) for d in dims]
    letters_b, letters_k = "abcdefgh"[:F], "ijklmnop"[:F]

    def embed(frags, block):
        block = numpy.asarray(block, dtype=numpy.float64).reshape([dims[m] for m in frags] * 2)
        operands, subs = [block], ["".join(letters_b[m] for m in frags) + "".join(letters_k[m] for m in frags)]
        for m in range(F):
            if m not in frags:
                operands.append(eyes[m])
                subs.append(letters_b[m] + letters_k[m])
        return numpy.einsum(",".join(subs) + "->" + letters_b + letters_k, *operands)

    for M in range(F):
        Hmat += embed((M,), H[0][M])
    for M, N in itertools.combinations(range(F), 2):
        Hmat += embed((M, N), H[1][M][N])
    if len(H) > 2:
        for M, N, O in itertools.combinations(range(F), 3):
            sub = [dims[M], dims[N], dims[O]]
            block = numpy.array(H[2][M][N][O], dtype=numpy.float64).reshape(sub * 2)
            idx = numpy.indices(sub * 2)
            changed = sum((idx[t] != idx[3 + t]).astype(int) for t in range(3))
            block[changed < 2] = 0.0
            Hmat += embed((M, N, O), block)
    return Hmat.reshape(D, D)
