"""Dependency analysis of a recorded launch sequence (lib.Context.begin_trace) and its placement on several streams.
Used by hermitian/plan.py (recorded get_xr_H) and general/distributed.recorded_step (recorded general-XRCC build).

A recorded get_xr_H is ~800 launches issued on one stream, most of them 10-microsecond GEMMs that do not depend on each
other (different diagrams, different charge sectors).  Captured as a chain, the GPU runs them one after the other.  This
module works out, from the ARGUMENTS of the recorded calls alone, which calls really depend on which -- every call reads
and writes address ranges of device buffers; scattered outputs (offset tables) are compared element by element -- and
spreads the sequence over a few streams with events only where a dependency crosses streams, so that the CUDA graph
captured from it has the true dependency structure and independent kernels overlap.

Conservative by construction: a call this module does not know, or an operand it cannot size, is a barrier.
"""
import numpy

EXACT_LIMIT = 1 << 18          # scattered outputs up to this many elements are compared element by element


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, (int, numpy.integer)):
        return int(x)
    return int(x.data_ptr())


class _tables(object):
    """host copies of the int64 offset tables the calls refer to (device tensors), by address; a table passed as a raw
    pointer (base + offset into a bigger table) is resolved through the tensors the caller names as `known`"""
    def __init__(self, known=()):
        self._host = {}
        self._known = [(int(t.data_ptr()), int(t.data_ptr()) + t.numel() * 8, t) for t in known
                       if hasattr(t, "data_ptr") and getattr(t, "dtype", None) is not None and str(t.dtype).endswith("int64") and t.numel()]

    def _copy(self, tensor):
        key = (int(tensor.data_ptr()), int(tensor.numel()))
        if key not in self._host:
            self._host[key] = tensor.detach().cpu().numpy().reshape(-1).astype(numpy.int64)
        return self._host[key]

    def get(self, table, count):
        if table is None:
            return None
        if hasattr(table, "data_ptr"):
            return self._copy(table)[:count]
        ptr = int(table)
        for lo, hi, tensor in self._known:
            if lo <= ptr and ptr + 8 * count <= hi:
                first = (ptr - lo) // 8
                return self._copy(tensor)[first:first + count]
        raise KeyError("offset table passed as a raw pointer into an unknown buffer")


class footprint(object):
    """reads / writes of one call: lists of (lo, hi, exact) with byte addresses [lo, hi) and, for scattered outputs, the
    sorted element addresses"""
    def __init__(self):
        self.reads, self.writes, self.barrier = [], [], False

    def read(self, ptr, elements):
        if ptr is not None and elements > 0:
            self.reads.append((ptr, ptr + 8 * int(elements), None))

    def write(self, ptr, elements):
        if ptr is not None and elements > 0:
            self.writes.append((ptr, ptr + 8 * int(elements), None))

    def write_scattered(self, ptr, om, on):
        lo = ptr + 8 * int(om.min() + on.min())
        hi = ptr + 8 * int(om.max() + on.max()) + 8
        exact = None
        if len(om) * len(on) <= EXACT_LIMIT:
            exact = numpy.unique((ptr + 8 * (om[:, None] + on[None, :])).reshape(-1))
        self.writes.append((lo, hi, exact))


def _overlap(a, b):
    for lo1, hi1, ex1 in a:
        for lo2, hi2, ex2 in b:
            if lo1 < hi2 and lo2 < hi1:
                if ex1 is not None and ex2 is not None:
                    if len(numpy.intersect1d(ex1, ex2, assume_unique=True)):
                        return True
                elif ex1 is not None:
                    k = numpy.searchsorted(ex1, lo2)
                    if k < len(ex1) and ex1[k] < hi2:
                        return True
                elif ex2 is not None:
                    k = numpy.searchsorted(ex2, lo1)
                    if k < len(ex2) and ex2[k] < hi1:
                        return True
                else:
                    return True
    return False


def _gemm_output(fp, tables, C, M, N, offM, ldc, offN):
    pC = _ptr(C)
    if offM is None and offN is None:
        fp.write(pC, (M - 1) * ldc + N)
        return
    om = tables.get(offM, M) if offM is not None else numpy.arange(M, dtype=numpy.int64) * ldc
    on = tables.get(offN, N) if offN is not None else numpy.arange(N, dtype=numpy.int64)
    fp.write_scattered(pC, om, on)


def footprint_of(name, args, kwargs, tables):
    fp = footprint()
    try:
        if kwargs:
            raise KeyError("keyword arguments")
        if name == "gemm_scatter":
            M, N, K, alpha, A, lda, B, ldb, C = args[:9]
            offM, ldc, offN = (list(args[9:12]) + [None, 0, None])[:3] if len(args) < 12 else args[9:12]
            fp.read(_ptr(A), (M - 1) * lda + K)
            fp.read(_ptr(B), (N - 1) * ldb + K)
            _gemm_output(fp, tables, C, M, N, offM, ldc, offN)
        elif name == "gemm_stream":
            E1, s1, E2, s2, EK1, sk1, EK2, N, alpha, A, B, ldb, C = args[:13]
            offM, ldc, offN = args[13:16] if len(args) >= 16 else (None, 0, None)
            fp.read(_ptr(A), (E1 - 1) * s1 + (E2 - 1) * s2 + (EK1 - 1) * sk1 + EK2)
            fp.read(_ptr(B), (N - 1) * ldb + EK1 * EK2)
            _gemm_output(fp, tables, C, E1 * E2, N, offM, ldc, offN)
        elif name == "permute_copy":
            dst, src, shape, strides = args[:4]
            total, span = 1, 1
            for extent, stride in zip(shape, strides):
                total *= int(extent)
                span += (int(extent) - 1) * int(stride)
            fp.read(_ptr(src), span)
            fp.write(_ptr(dst), total)
        elif name == "memset_zero":
            fp.write(_ptr(args[0]), (int(args[1]) + 7) // 8)
        elif name == "scatter_const":
            C, idx, count = args[:3]
            at = tables.get(idx, count)
            fp.write_scattered(_ptr(C), at, numpy.zeros(1, dtype=numpy.int64))
        elif name == "gemm_dd":
            M, N, K, A, lda, B, ldb, C0, ldc0, sign, out, ldo = args[:12]
            fp.read(_ptr(A), (M - 1) * lda + K)
            fp.read(_ptr(B), (K - 1) * ldb + N)
            if C0 is not None:
                fp.read(_ptr(C0), (M - 1) * ldc0 + N)
            fp.write(_ptr(out), (M - 1) * ldo + N)
        elif name == "trimer_stream":
            n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end, mode = args[:14]
            rest = list(args[14:]) + [None] * 5
            moments, C, offA, offB, offC = rest[:5]
            fp.read(_ptr(W), (Pa - 1) * ldw + n * n)
            fp.read(_ptr(beta), (Pb - 1) * ldbeta + n)
            fp.read(_ptr(gamma), (Pc - 1) * ldgamma + n)
            if mode == 0:
                fp.write(_ptr(moments), 2)
            else:
                oa, ob, oc = tables.get(offA, Pa)[a_begin:a_end], tables.get(offB, Pb), tables.get(offC, Pc)
                lo = int(oa.min() + ob.min() + oc.min())
                hi = int(oa.max() + ob.max() + oc.max()) + 1
                fp.writes.append((_ptr(C) + 8 * lo, _ptr(C) + 8 * hi, None))
        elif name == "copy2d_scaled":
            dst, dst_ld, src, src_ld, rows, cols = args[:6]
            fp.read(_ptr(src), (rows - 1) * src_ld + cols)
            fp.write(_ptr(dst), (rows - 1) * dst_ld + cols)
        else:
            fp.barrier = True
    except (KeyError, TypeError, ValueError, AttributeError):
        fp = footprint()
        fp.barrier = True
    return fp


def dependencies(trace, known=()):
    """deps[i] = sorted indices j < i that call i must run after (true, anti and output dependencies); transitively implied
    ones are dropped where that is cheap to see (a dependency on a call that itself depends on the other).  known: device
    tensors that raw-pointer offset tables may point into."""
    tables = _tables(known)
    prints = [footprint_of(call.__name__, args, kwargs, tables) for call, args, kwargs in trace]
    deps = []
    last_barrier = -1
    for i, fi in enumerate(prints):
        mine = set()
        if fi.barrier:
            mine.update(range(max(last_barrier, 0), i))
            last_barrier = i
        else:
            if last_barrier >= 0:
                mine.add(last_barrier)
            for j in range(i - 1, last_barrier, -1):
                fj = prints[j]
                if _overlap(fj.writes, fi.reads) or _overlap(fj.writes, fi.writes) or _overlap(fj.reads, fi.writes):
                    mine.add(j)
        implied = set()
        for j in mine:
            implied.update(d for d in deps[j] if d in mine)
        deps.append(sorted(mine - implied))
    return deps


def assign_streams(deps, n_streams):
    """(stream of every call, cross-stream dependencies of every call): a call follows its latest dependency onto that
    dependency's stream when it is still the last call there, else it takes the stream that has been idle longest"""
    stream_of, cross = [], []
    last_on = [-1] * n_streams
    for i, d in enumerate(deps):
        choice = None
        for j in reversed(d):
            if last_on[stream_of[j]] == j:
                choice = stream_of[j]
                break
        if choice is None:
            choice = min(range(n_streams), key=lambda s: last_on[s])
        stream_of.append(choice)
        cross.append([j for j in d if stream_of[j] != choice])
        last_on[choice] = i
    return stream_of, cross
