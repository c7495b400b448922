"""Device tensors and the pairwise contraction planner of the hermitian path.

What the reference reaches through ``XR_tensor.raw`` -> qode.math.tensornet -> tensorly/opt_einsum
-> numpy/BLAS (XRbase/XR_tensor.py:49-58) is here ONE device primitive used two ways:

    contract(A, idxA, B, idxB, out, idx_out)   out[idx_out] (+)= alpha * sum_shared A[idxA] * B[idxB]

planned as  (optional xr_permute_copy of an operand so its contracted indices trail) ->
xr_gemm_scatter with int64 offset tables that place every (row, column) of the GEMM at its final
address, so any output index order -- and accumulation straight into a block of a bigger matrix --
costs nothing extra.  No arithmetic happens in torch or numpy.
"""
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
    if isinstance(x, DeviceTensor):
        return x.host()
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


class DeviceStore(object):
    """host ndarray -> device copy, uploaded once (keyed by the host buffer's identity)."""
    def __init__(self, dev=None):
        self.dev = dev or default_device()
        self._by_id = {}
    def get(self, obj):
        if isinstance(obj, DeviceTensor):
            return obj
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

    def contract(self, A, idxA, B, idxB, idx_out, alpha=1.0, out=None, out_offset=0, out_strides=None, accumulate=False):
        """out[idx_out] (+)= alpha * sum_{shared} A[idxA] * B[idxB].

        idx* are sequences of hashable labels.  Labels present in both operands and absent from idx_out are
        summed.  `out` may be a bigger buffer: then out_strides[label] gives the element stride of every
        output label inside it and out_offset the element offset of the block (this is how a diagram block is
        accumulated in place into the packed XR matrix)."""
        idxA, idxB, idx_out = list(idxA), list(idxB), list(idx_out)
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
        A2 = self._arranged(A, idxA, rows, shared)
        B2 = self._arranged(B, idxB, cols, shared)
        M = int(numpy.prod([extent[l] for l in rows])) if rows else 1
        N = int(numpy.prod([extent[l] for l in cols])) if cols else 1
        K = int(numpy.prod([extent[l] for l in shared])) if shared else 1
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
            raise NotImplementedError("outer products of unconnected factors")
        _, a, b, keep = best
        merged = self.contract(factors[a][0], factors[a][1], factors[b][0], factors[b][1], keep)
        factors = [f for k, f in enumerate(factors) if k not in (a, b)] + [(merged, keep)]
    (A, idxA), (B, idxB) = factors
    return self.contract(A, idxA, B, idxB, idx_out, alpha, out, out_offset, out_strides, accumulate)


Contractor.multi_contract = _multi_contract
