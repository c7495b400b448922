"""Matrix-free action of the excitonic Hamiltonian, y = Hmat . v, from the FACTORS of the build (the
consumer side of general-XRCC -- ``hamiltonian.braket_loops`` expands H1/H2/H3 into Hmat and ``excitonic.fci`` hands it to
an eigen-solver, hamiltonian.py:21-84 / excitonic.py:60-85).

Nothing of size dim^2 is formed: a dimer class is  H2[(i1 i2),(j1 j2)] = sum_k A[(i1 j1),k] B[(i2 j2),k]  and a trimer
class  H3[(ik ib ic),(jk jb jc)] = alpha sum_rs W[(ik jk),r,s] beta[(ib jb),r] gamma[(ic jc),s]  (build_H.py:69-188 after the
rho x V precontractions), so

    y[ik ib ic ..] += alpha sum_{jk r s} W[ik jk r s] sum_{jb} beta[ib jb r] sum_{jc} gamma[ic jc s] v[jk jb jc ..]

is three small GEMMs per (class, charge-sector triple) -- O(N^4 n^2) work per trimer instead of the O(N^6 n) of streaming
every element -- all through xr_gemm_scatter / xr_permute_copy (hermitian.tensor.Contractor does the index bookkeeping).
v is a device tensor over the product basis, one axis per fragment in ``state_indices`` order (test_H.py:101-126 layout
when flattened), optionally with trailing axes (several vectors at once).
"""
import itertools
import numpy
import torch

from ..hermitian.tensor import Contractor, DeviceTensor


class xr_operator(object):
    def __init__(self, engine, fragments=None, monomers=True, dimers=True, trimers=True):
        self.eng = engine
        self.dev = engine.dev
        self.frags = list(range(len(engine._supersystem))) if fragments is None else list(fragments)
        self.info = [engine._frag(m) for m in self.frags]
        self.dims = [f.dim for f in self.info]
        self.contractor = Contractor(self.dev)
        self.terms = []           # (factors [(DeviceTensor, labels)], alpha, bra ranges {axis: (lo, hi)}, ket ranges)
        self._prepare(monomers, dimers, trimers)

    # -------------------------------------------------------------------------------- preparation
    def _range(self, info, chg, i_lo=0, i_hi=None):
        """matrix positions [lo, hi) of states i_lo..i_hi of one charge sector (sectors are contiguous in state_indices)"""
        pos = info.pos[chg]
        i_hi = len(pos) if i_hi is None else i_hi
        if not numpy.array_equal(pos, pos[0] + numpy.arange(len(pos))):
            raise NotImplementedError("states of charge %r are not contiguous in state_indices" % (chg,))
        return int(pos[0]) + i_lo, int(pos[0]) + i_hi

    def _sector_block(self, buf, ld, off, rows, cols, shape):
        """compact copy of rows [off, off+rows) x columns [0, cols) of a class factor buffer"""
        out = self.dev.empty((rows, cols))
        self.dev.ctx.copy2d_scaled(out, cols, buf.data_ptr() + 8 * off * ld, ld, rows, cols, 1.0)
        return DeviceTensor(out.reshape(shape), self.dev)

    def _prepare(self, monomers, dimers, trimers):
        eng, F = self.eng, len(self.frags)
        for x, m in enumerate(self.frags if monomers else []):
            H1 = DeviceTensor(eng.H1_device(m), self.dev)
            self.terms.append(([(H1, [("i", x), ("j", x)])], 1.0, {x: (0, self.dims[x])}, {x: (0, self.dims[x])}))
        for x1, x2 in itertools.combinations(range(F) if dimers else [], 2):
            m1, m2 = self.frags[x1], self.frags[x2]
            f1, f2 = self.info[x1], self.info[x2]
            for d1, c1, c2, A, B, K, ld in eng._dimer_class_factors(m1, m2):
                for ci1, cj1, lo1, hi1, off1 in c1.sectors:
                    Nj1 = f1.n_states[cj1]
                    As = self._sector_block(A, ld, off1, (hi1 - lo1) * Nj1, K, (hi1 - lo1, Nj1, K))
                    for ci2, cj2, lo2, hi2, off2 in c2.sectors:
                        Nj2 = f2.n_states[cj2]
                        Bs = self._sector_block(B, ld, off2, (hi2 - lo2) * Nj2, K, (hi2 - lo2, Nj2, K))
                        factors = [(As, [("i", x1), ("j", x1), "k"]), (Bs, [("i", x2), ("j", x2), "k"])]
                        bra = {x1: self._range(f1, ci1, lo1, hi1), x2: self._range(f2, ci2, lo2, hi2)}
                        ket = {x1: self._range(f1, cj1), x2: self._range(f2, cj2)}
                        self.terms.append((factors, 1.0, bra, ket))
        if not trimers:
            return
        for xs in itertools.combinations(range(F), 3):
            ms = tuple(self.frags[x] for x in xs)
            for cl in eng._trimer_classes(ms):
                fac = eng._trimer_factors(ms, cl)
                if fac is None:
                    continue
                roles = [cl["k"], cl["b"], cl["c"]]
                xk, xb, xc = (xs[r] for r in roles)
                fk, fb, fc = (self.info[x] for x in (xk, xb, xc))
                nb = nc = fac["n"]          # both padded to the larger orbital count by _trimer_factors
                ldb, ldc = fac["beta"].shape[1], fac["gamma"].shape[1]
                for cik, cjk, lok, hik, offk in fac["ck"].sectors:
                    Njk = fk.n_states[cjk]
                    Ws = self._sector_block(fac["W"], fac["ldw"], offk, (hik - lok) * Njk, nb * nc, (hik - lok, Njk, nb, nc))
                    for cib, cjb, lob, hib, offb in fac["cb"].sectors:
                        Njb = fb.n_states[cjb]
                        Bs = self._sector_block(fac["beta"], ldb, offb, (hib - lob) * Njb, nb, (hib - lob, Njb, nb))
                        for cic, cjc, loc, hic, offc in fac["cc"].sectors:
                            Njc = fc.n_states[cjc]
                            Gs = self._sector_block(fac["gamma"], ldc, offc, (hic - loc) * Njc, nc, (hic - loc, Njc, nc))
                            factors = [(Ws, [("i", xk), ("j", xk), "r", "s"]), (Bs, [("i", xb), ("j", xb), "r"]),
                                       (Gs, [("i", xc), ("j", xc), "s"])]
                            bra = {xk: self._range(fk, cik, lok, hik), xb: self._range(fb, cib, lob, hib),
                                   xc: self._range(fc, cic, loc, hic)}
                            ket = {xk: self._range(fk, cjk), xb: self._range(fb, cjb), xc: self._range(fc, cjc)}
                            self.terms.append((factors, fac["alpha"], bra, ket))

    # -------------------------------------------------------------------------------- application
    def apply(self, v, out=None):
        """y = Hmat . v; v: contiguous CUDA float64 tensor of shape dims (+ trailing axes).  Returns y (same shape)."""
        F = len(self.dims)
        if tuple(v.shape[:F]) != tuple(self.dims) or v.dtype != torch.float64 or not v.is_contiguous():
            raise ValueError("v must be a contiguous float64 tensor of shape %r (+ trailing axes)" % (tuple(self.dims),))
        extra = list(v.shape[F:])
        y = out if out is not None else self.dev.zeros(tuple(v.shape))
        shape = list(v.shape)
        strides = [int(numpy.prod(shape[a + 1:], dtype=numpy.int64)) for a in range(len(shape))]
        axes = list(range(F)) + ["x%d" % e for e in range(len(extra))]           # labels of non-contracted axes
        slices, group = {}, None
        for factors, alpha, bra, ket in self.terms:
            if tuple(sorted(ket)) != group:        # terms come grouped by fragment tuple: drop the previous group's slices of v
                slices, group = {}, tuple(sorted(ket))
            key = tuple(sorted(ket.items()))
            if key not in slices:           # compact copy of the ket-sector slice of v (shared by every term on that sector)
                sub = [ket[a][1] - ket[a][0] if a in ket else shape[a] for a in range(len(shape))]
                first = sum(ket[a][0] * strides[a] for a in ket)
                if sub == shape:
                    slices[key] = DeviceTensor(v, self.dev)
                else:
                    buf = self.dev.empty(tuple(sub))
                    self.dev.ctx.permute_copy(buf, v.data_ptr() + 8 * first, sub, strides, 1.0)
                    slices[key] = DeviceTensor(buf, self.dev)
            v_labels = [("j", a) if a in ket else ("s", a) for a in range(F)] + axes[F:]
            y_labels = [("i", a) if a in bra else ("s", a) for a in range(F)] + axes[F:]
            out_strides = dict(zip(y_labels, strides))
            offset = sum(bra[a][0] * strides[a] for a in bra)
            self.contractor.multi_contract(list(factors) + [(slices[key], v_labels)], y_labels, alpha, out=y, out_offset=offset,
                                           out_strides=out_strides, accumulate=True)
        return y

    def dense(self):
        """Hmat itself (small systems / tests): the operator applied to the identity"""
        D = int(numpy.prod(self.dims, dtype=numpy.int64))
        eye = torch.eye(D, dtype=torch.float64, device=self.dev.torch_device).reshape(tuple(self.dims) + (D,))
        return self.apply(eye.contiguous()).reshape(D, D)
