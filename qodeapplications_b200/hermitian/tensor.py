"""Device tensors and the pairwise contraction planner of the hermitian path.

What the reference reaches through ``XR_tensor.raw`` -> qode.math.tensornet -> tensorly/opt_einsum
-> numpy/BLAS (XRbase/XR_tensor.py:49-58) is here ONE device primitive used two ways:

    contract(A, idxA, B, idxB, out, idx_out)   out[idx_out] (+)= alpha * sum_shared A[idxA] * B[idxB]

planned as  (optional xr_permute_copy of an operand so its contracted indices trail) ->
xr_gemm_scatter with int64 offset tables that place every (row, column) of the GEMM at its final
address, so any output index order -- and accumulation straight into a block of a bigger matrix --
costs nothing extra.  No arithmetic happens in torch or numpy.
"""
import itertools
import numpy
import torch

from ..device import Device

_default_device = None


def default_device():
    global _default_device
    if _default_device is None:
        _default_device = Device()
    return _default_device


def as_host(x):
    """ndarray view of whatever the caller wrapped its data in (ndarray, XR_tensor.init object,
    qode tensornet primitive, torch tensor)."""
    if isinstance(x, numpy.ndarray):
        return x
    if isinstance(x, (DeviceTensor, FactoredTensor)):
        return x.host()
    if hasattr(x, "factors") or hasattr(x, "terms"):          # a lazy XR_tensor expression
        return DeviceStore(default_device()).get(x).host()
    for attr in ("array", "data", "_raw_tensor"):
        inner = getattr(x, attr, None)
        if isinstance(inner, numpy.ndarray):
            return inner
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return numpy.asarray(x, dtype=numpy.float64)


class DeviceTensor(object):
    """Contiguous FP64 tensor in HBM (a torch tensor used purely as a buffer)."""
    def __init__(self, buf, dev):
        self.buf = buf
        self.dev = dev
    @property
    def shape(self):
        return tuple(self.buf.shape)
    @property
    def ndim(self):
        return self.buf.dim()
    def host(self):
        return self.dev.download(self.buf)
    def __array__(self, dtype=None, copy=None):
        out = self.host()
        return out if dtype is None else out.astype(dtype)
    def __len__(self):
        return self.buf.shape[0]


class FactoredTensor(object):
    """A tensor given as a SUM of PRODUCTS of device tensors and never formed: terms = [(scalar, [(DeviceTensor, labels)])],
    where an int label k says "axis k of the tensor this stands for" and a string label is summed inside the term.  This
    is what a lazy XR_tensor expression (e.g. the decomposed rank-6..8 densities of frag-states/decomps.py, sums of outer
    products of a core and a valence density) becomes on the device; Contractor splices the factors of every term into
    the product it is asked for, so a diagram with such operands is contracted factor by factor."""
    def __init__(self, terms, dev):
        self.terms, self.dev = terms, dev
        extent = {}
        for T, labels in terms[0][1]:
            extent.update((l, e) for l, e in zip(labels, T.shape) if isinstance(l, (int, numpy.integer)))
        if sorted(extent) != list(range(len(extent))):
            raise ValueError("a factored tensor needs integer labels 0..rank-1 on its free axes, got %r" % sorted(extent))
        self._shape = tuple(extent[k] for k in range(len(extent)))
    @property
    def shape(self):
        return self._shape
    @property
    def ndim(self):
        return len(self._shape)
    def __len__(self):
        return self._shape[0]
    def spliced(self, labels, tag):
        """[(scalar, factors)] with axis k labelled labels[k] and the internal labels made private to operand `tag`"""
        out = []
        for t, (scalar, factors) in enumerate(self.terms):
            out.append((scalar, [(T, [labels[l] if isinstance(l, (int, numpy.integer)) else "%s.%d:%s" % (tag, t, l) for l in ls])
                                 for T, ls in factors]))
        return out
    def appended(self, labels, extra, axes_out):
        """the factored tensor of  sum_{contracted} self[labels] * extra[0][extra[1]]  with result axes axes_out (labels):
        nothing is contracted, the extra factor just joins every term (a lazy precontraction)"""
        def back(l, t):
            return axes_out.index(l) if l in axes_out else "pc%d:%s" % (t, l)
        terms = []
        for t, (scalar, factors) in enumerate(self.spliced(labels, "r")):
            terms.append((scalar, [(T, [back(l, t) for l in ls]) for T, ls in factors + [extra]]))
        return FactoredTensor(terms, self.dev)
    def dense(self, contractor=None):
        contractor = contractor or Contractor(self.dev)
        axes = list(range(self.ndim))
        return contractor.multi_contract([(self, axes)], axes)
    def host(self):
        return self.dense().host()
    def __array__(self, dtype=None, copy=None):
        out = self.host()
        return out if dtype is None else out.astype(dtype)


class DeviceStore(object):
    """host ndarray -> device copy, uploaded once (keyed by the host buffer's identity)."""
    def __init__(self, dev=None):
        self.dev = dev or default_device()
        self._by_id = {}
    def get(self, obj):
        if isinstance(obj, (DeviceTensor, FactoredTensor)):
            return obj
        if hasattr(obj, "terms") or hasattr(obj, "factors"):       # lazy XR_tensor sum / product: keep it factored
            hit = self._by_id.get(id(obj))
            if hit is None or hit[0] is not obj:
                products = obj.terms if hasattr(obj, "terms") else [obj]
                terms = [(float(t.scalar), [(self.get(T), tuple(ls)) for T, ls in t.factors]) for t in products]
                for _, factors in terms:
                    if any(isinstance(T, FactoredTensor) for T, _ in factors):
                        raise NotImplementedError("nested factored tensors")
                hit = (obj, FactoredTensor(terms, self.dev))
                self._by_id[id(obj)] = hit
            return hit[1]
        host = as_host(obj)
        key = id(host) if isinstance(obj, numpy.ndarray) else id(obj)
        hit = self._by_id.get(key)
        if hit is None or hit[0] is not obj:
            dt = DeviceTensor(self.dev.upload(host), self.dev)
            self._by_id[key] = (obj, dt)     # keeps obj alive so the id stays unique
            return dt
        return hit[1]


def _strides(shape):
    s, acc = [], 1
    for extent in reversed(shape):
        s.append(acc)
        acc *= extent
    return list(reversed(s))


def _offset_table(extents, strides):
    """int64 table over the row-major product of `extents`: sum_d i_d * strides[d]"""
    table = numpy.zeros(1, dtype=numpy.int64)
    for extent, stride in zip(extents, strides):
        table = (table[:, None] + (numpy.arange(extent, dtype=numpy.int64) * stride)[None, :]).reshape(-1)
    return table


_TABLES = {}        # id(device) -> {(extents, strides): device int64 table}; pure functions of the shapes, kept across calls


class Contractor(object):
    def __init__(self, dev=None):
        self.dev = dev or default_device()
        self._tables = _TABLES.setdefault(id(self.dev), {"device": self.dev})
        self.flops = 0.0

    def _table(self, extents, strides):
        key = (tuple(extents), tuple(strides))
        if key not in self._tables:
            if len(self._tables) > 4096:        # bound the cache (an optimiser loop reuses a few hundred shapes)
                for stale in [k for k in self._tables if k != "device"]:
                    del self._tables[stale]
            self._tables[key] = self.dev.upload(_offset_table(extents, strides), dtype=numpy.int64)
        return self._tables[key]

    def _arranged(self, T, idx, lead, tail):
        """T viewed as a row-major matrix [prod(lead), prod(tail)]; copies through xr_permute_copy only
        if the index order differs."""
        want = list(lead) + list(tail)
        idx = list(idx)
        if want == idx:
            return T.buf
        strides = _strides(T.shape)
        shape = [T.shape[idx.index(l)] for l in want]
        src = [strides[idx.index(l)] for l in want]
        out = self.dev.empty(tuple(shape))
        self.dev.ctx.permute_copy(out, T.buf, shape, src, 1.0)
        return out

    STREAM_MIN_ELEMENTS = 1 << 20

    def _stream_plan(self, A, idxA, rowsA, M, B, idxB, colsB, N, shared, K, extent):
        """A large operand whose free indices are INTERLEAVED with contracted ones (a density rho[i,j,p,X,s,r] contracted over
        p,s,r) used to be re-ordered by xr_permute_copy -- a read and a write of the whole tensor -- before the GEMM could
        read it.  When the other operand is small (<= 32 output columns) xr_gemm_stream reads the large one where it
        lies: its index list, in memory order, must be [free run R1][contracted run K1][free run R2][contracted run K2].
        Returns (large tensor, small operand arranged as [its free labels, contracted labels in the large one's order], free
        labels of the large one in memory order, free labels of the small one, (E1, s1, E2, s2, EK1, sk1, EK2)) or None."""
        for T, idxT, freeT, nT, S, idxS, freeS, nS in ((A, idxA, rowsA, M, B, idxB, colsB, N), (B, idxB, colsB, N, A, idxA, rowsA, M)):
            if nS > 32 or nT * K < self.STREAM_MIN_ELEMENTS or K < 256 or K % 2:
                continue
            runs = []                                   # [(is_contracted, [labels])] in memory order
            for l in idxT:
                kind = l in shared
                if runs and runs[-1][0] == kind:
                    runs[-1][1].append(l)
                else:
                    runs.append((kind, [l]))
            if [k for k, _ in runs] != [False, True, False, True]:
                continue
            size = lambda labels: int(numpy.prod([extent[l] for l in labels]))
            R1, K1, R2, K2 = (labels for _, labels in runs)
            E1, EK1, E2, EK2 = size(R1), size(K1), size(R2), size(K2)
            s2, sk1, s1 = EK2, E2 * EK2, EK1 * E2 * EK2
            if s2 % 2 or T.buf.data_ptr() % 16:
                continue
            small2 = self._arranged(S, idxS, freeS, K1 + K2)
            return T, small2, R1 + R2, list(freeS), (E1, s1, E2, s2, EK1, sk1, EK2)
        return None

    def contract(self, A, idxA, B, idxB, idx_out, alpha=1.0, out=None, out_offset=0, out_strides=None, accumulate=False):
        """out[idx_out] (+)= alpha * sum_{shared} A[idxA] * B[idxB].

        idx* are sequences of hashable labels.  Labels present in both operands and absent from idx_out are
        summed.  `out` may be a bigger buffer: then out_strides[label] gives the element stride of every
        output label inside it and out_offset the element offset of the block (this is how a diagram block is
        accumulated in place into the packed XR matrix)."""
        idxA, idxB, idx_out = list(idxA), list(idxB), list(idx_out)
        if isinstance(A, FactoredTensor) or isinstance(B, FactoredTensor):
            return self.multi_contract([(A, idxA), (B, idxB)], idx_out, alpha, out, out_offset, out_strides, accumulate)
        shared = [l for l in idxA if l in idxB and l not in idx_out]
        rows = [l for l in idxA if l not in shared]
        cols = [l for l in idxB if l not in shared]
        if sorted(map(str, rows + cols)) != sorted(map(str, idx_out)):
            raise ValueError("contract: output labels %r do not match free labels %r" % (idx_out, rows + cols))
        extent = {}
        for T, idx in ((A, idxA), (B, idxB)):
            for l, e in zip(idx, T.shape):
                if extent.setdefault(l, e) != e:
                    raise ValueError("contract: label %r has extents %d and %d" % (l, extent[l], e))
        M = int(numpy.prod([extent[l] for l in rows])) if rows else 1
        N = int(numpy.prod([extent[l] for l in cols])) if cols else 1
        K = int(numpy.prod([extent[l] for l in shared])) if shared else 1
        streamed = self._stream_plan(A, idxA, rows, M, B, idxB, cols, N, shared, K, extent)
        if streamed is None:
            A2 = self._arranged(A, idxA, rows, shared)
            B2 = self._arranged(B, idxB, cols, shared)
        if out is None:
            out_buf = self.dev.empty(tuple(extent[l] for l in idx_out))
            stride_of = dict(zip(idx_out, _strides([extent[l] for l in idx_out])))
            accumulate = False
        else:
            out_buf = out.buf if isinstance(out, DeviceTensor) else out
            stride_of = out_strides if out_strides is not None else dict(zip(idx_out, _strides([extent[l] for l in idx_out])))
        row_strides = [stride_of[l] for l in rows]
        col_strides = [stride_of[l] for l in cols]
        plain = (out is None or out_strides is None) and idx_out == rows + cols
        base = out_buf.data_ptr() + 8 * int(out_offset)
        if streamed is not None:
            big, small2, big_free, small_free, geometry = streamed
            offM = self._table([extent[l] for l in big_free], [stride_of[l] for l in big_free])
            offN = self._table([extent[l] for l in small_free], [stride_of[l] for l in small_free])
            n_small = int(numpy.prod([extent[l] for l in small_free])) if small_free else 1
            if self.dev.ctx.gemm_stream(*geometry, n_small, alpha, big.buf, small2, K, base, offM, 0, offN, accumulate):
                self.flops += 2.0 * M * N * K
                return DeviceTensor(out_buf, self.dev) if out is None else out
            A2 = self._arranged(A, idxA, rows, shared)      # the kernel declined (alignment): the re-ordering path
            B2 = self._arranged(B, idxB, cols, shared)
        if plain:
            self.dev.ctx.gemm_scatter(M, N, K, alpha, A2, K, B2, K, base, None, N, None, accumulate)
        else:
            offM = self._table([extent[l] for l in rows], row_strides)
            offN = self._table([extent[l] for l in cols], col_strides)
            self.dev.ctx.gemm_scatter(M, N, K, alpha, A2, K, B2, K, base, offM, 0, offN, accumulate)
        self.flops += 2.0 * M * N * K
        return DeviceTensor(out_buf, self.dev) if out is None else out


def _multi_contract(self, factors, idx_out, alpha=1.0, out=None, out_offset=0, out_strides=None, accumulate=False):
    """Product of any number of tensors: out[idx_out] (+)= alpha * sum over every label not in idx_out.
    factors = [(DeviceTensor, labels), ...].  Pairs are contracted greedily (smallest intermediate first, the
    choice opt_einsum's greedy path makes for these chains); the last pairwise contraction writes into `out`."""
    factors = [(T, list(idx)) for T, idx in factors]
    idx_out = list(idx_out)
    if len(factors) == 1:
        T, idx = factors[0]
        one = DeviceTensor(self.dev.upload(numpy.ones((1,))), self.dev)
        if out is None:       # scaled (and possibly permuted) copy: a GEMM against the 1x1 identity
            res = self.contract(T, idx, one, ["__one"], idx_out + ["__one"], alpha)
            return DeviceTensor(res.buf.reshape(res.shape[:-1]), self.dev)
        strides = None if out_strides is None else dict(out_strides, __one=0)
        if strides is None:
            strides = dict(zip(idx_out, _strides([T.shape[idx.index(l)] for l in idx_out])), __one=0)
        return self.contract(T, idx, one, ["__one"], idx_out + ["__one"], alpha, out, out_offset, strides, accumulate)
    while len(factors) > 2:
        extent = {}
        for T, idx in factors:
            extent.update(zip(idx, T.shape))
        best = None
        for a in range(len(factors)):
            for b in range(a + 1, len(factors)):
                la, lb = factors[a][1], factors[b][1]
                shared = [l for l in la if l in lb]
                if not shared:
                    continue
                elsewhere = set(idx_out)
                for k, (_, ls) in enumerate(factors):
                    if k not in (a, b):
                        elsewhere.update(ls)
                if any(l in elsewhere for l in shared):
                    raise NotImplementedError("label shared by more than two factors")
                keep = [l for l in la if l not in shared] + [l for l in lb if l not in shared]
                size = 1
                for l in keep:
                    size *= extent[l]
                if best is None or size < best[0]:
                    best = (size, a, b, keep)
        if best is None:
            # no two factors share a label (outer products of a core and a valence density, frag-states/decomps.py): take
            # the outer product of the two smallest factors -- a K = 1 GEMM -- and go on
            order = sorted(range(len(factors)), key=lambda k: int(numpy.prod(factors[k][0].shape)))
            a, b = sorted(order[:2])
            best = (0, a, b, factors[a][1] + factors[b][1])
        _, a, b, keep = best
        merged = self.contract(factors[a][0], factors[a][1], factors[b][0], factors[b][1], keep)
        factors = [f for k, f in enumerate(factors) if k not in (a, b)] + [(merged, keep)]
    (A, idxA), (B, idxB) = factors
    return self.contract(A, idxA, B, idxB, idx_out, alpha, out, out_offset, out_strides, accumulate)


def _multi_contract_factored(self, factors, idx_out, alpha=1.0, out=None, out_offset=0, out_strides=None, accumulate=False):
    """multi_contract with FactoredTensor operands: the product is expanded over the terms of every factored operand and
    each combination is contracted factor by factor, all of them accumulated into the one output."""
    if not any(isinstance(T, FactoredTensor) for T, _ in factors):
        return _multi_contract(self, factors, idx_out, alpha, out, out_offset, out_strides, accumulate)
    idx_out = list(idx_out)
    choices = []
    for k, (T, idx) in enumerate(factors):
        choices.append(T.spliced(list(idx), "f%d" % k) if isinstance(T, FactoredTensor) else [(1.0, [(T, list(idx))])])
    if out is None:
        extent = {}
        for T, idx in factors:
            extent.update(zip(idx, T.shape))
        out = DeviceTensor(self.dev.zeros(tuple(extent[l] for l in idx_out)), self.dev)
        accumulate = True
    first = True
    for combo in itertools.product(*choices):
        scalar, spliced = alpha, []
        for s, fs in combo:
            scalar *= s
            spliced += fs
        _multi_contract(self, spliced, idx_out, scalar, out, out_offset, out_strides, accumulate or not first)
        first = False
    return out


Contractor.multi_contract = _multi_contract_factored
