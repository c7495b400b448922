"""TEST-ONLY stand-in for the device layer, so the HOST logic of the product (offset tables, sign
folding, class bookkeeping, diagram planning) can be exercised by the CPU test-suite.

It implements the semantics of the C ABI (include/xr_b200.h) on raw host pointers with NumPy and is
injected by the tests in place of qodeapplications_b200.device.Device.  It lives under tests/ and is
never imported by the product: the product has no CPU path.  The CUDA kernels themselves are checked
against the oracle by the `-m gpu` tests.
"""
import ctypes
import numpy
import torch


def _view(ptr, count, dtype=numpy.float64):
    if hasattr(ptr, "data_ptr"):
        ptr = ptr.data_ptr()
    if count == 0:
        return numpy.zeros(0, dtype=dtype)
    ctype = ctypes.c_double if dtype == numpy.float64 else ctypes.c_int64
    return numpy.ctypeslib.as_array((ctype * int(count)).from_address(int(ptr)))


def _recorded(method):
    def call(self, *args, **kwargs):
        if self._trace is not None:
            self._trace.append((call, args, kwargs))
        return method(self, *args, **kwargs)
    call.__name__ = method.__name__
    return call


class FakeContext(object):
    def __init__(self):
        self.launches = 0
        self._trace = None

    def begin_trace(self):
        self._trace = []

    def end_trace(self):
        trace, self._trace = self._trace, None
        return trace

    def replay(self, trace):
        for call, args, kwargs in trace:
            call(self, *args, **kwargs)

    @_recorded
    def memset_zero(self, ptr, nbytes):
        _view(ptr, nbytes // 8)[...] = 0.0

    def launch_count(self):
        return self.launches

    def sync(self):
        pass

    @_recorded
    def gemm_scatter(self, M, N, K, alpha, A, lda, B, ldb, C, offM=None, ldc=0, offN=None, accumulate=False):
        self.launches += 1
        if M <= 0 or N <= 0:
            return
        a = _view(A, (M - 1) * lda + K).copy() if K else numpy.zeros(0)
        b = _view(B, (N - 1) * ldb + K).copy() if K else numpy.zeros(0)
        Am = numpy.lib.stride_tricks.as_strided(a, (M, K), (8 * lda, 8)) if K else numpy.zeros((M, 0))
        Bm = numpy.lib.stride_tricks.as_strided(b, (N, K), (8 * ldb, 8)) if K else numpy.zeros((N, 0))
        om = _view(offM, M, numpy.int64) if offM is not None else numpy.arange(M, dtype=numpy.int64) * ldc
        on = _view(offN, N, numpy.int64) if offN is not None else numpy.arange(N, dtype=numpy.int64)
        at = (om[:, None] + on[None, :]).reshape(-1)
        c = _view(C, int(at.max()) + 1)
        val = (alpha * (Am @ Bm.T)).reshape(-1)
        if accumulate:
            numpy.add.at(c, at, val)
        else:
            c[at] = val

    @_recorded
    def gemm_stream(self, E1, s1, E2, s2, EK1, sk1, EK2, N, alpha, A, B, ldb, C, offM=None, ldc=0, offN=None, accumulate=False):
        self.launches += 1
        if N > 32 or (E2 > 1 and s2 % 2) or (EK1 > 1 and sk1 % 2) or (E1 > 1 and s1 % 2) or ldb % 2:
            return False
        span = (E1 - 1) * s1 + (E2 - 1) * s2 + (EK1 - 1) * sk1 + EK2
        a = numpy.lib.stride_tricks.as_strided(_view(A, span), (E1, E2, EK1, EK2), (8 * s1, 8 * s2, 8 * sk1, 8))
        M, K = E1 * E2, EK1 * EK2
        a = numpy.ascontiguousarray(a).reshape(M, K)
        b = numpy.lib.stride_tricks.as_strided(_view(B, (N - 1) * ldb + K), (N, K), (8 * ldb, 8))
        prod = alpha * (a @ b.T)
        om = _view(offM, M, numpy.int64) if offM is not None else numpy.arange(M, dtype=numpy.int64) * ldc
        on = _view(offN, N, numpy.int64) if offN is not None else numpy.arange(N, dtype=numpy.int64)
        at = (om[:, None] + on[None, :]).reshape(-1)
        c = _view(C, int(at.max()) + 1)
        if accumulate:
            numpy.add.at(c, at, prod.reshape(-1))
        else:
            c[at] = prod.reshape(-1)
        return True

    @_recorded
    def gemm_reduce(self, M, N, K, alpha, A, lda, B, ldb, moments):
        self.launches += 1
        if M <= 0 or N <= 0:
            return
        assert lda % 2 == 0 and ldb % 2 == 0, "xr_gemm_reduce needs even leading dimensions"
        Am = numpy.lib.stride_tricks.as_strided(_view(A, (M - 1) * lda + K).copy(), (M, K), (8 * lda, 8))
        Bm = numpy.lib.stride_tricks.as_strided(_view(B, (N - 1) * ldb + K).copy(), (N, K), (8 * ldb, 8))
        C = alpha * (Am @ Bm.T)
        m = _view(moments, 2)
        m[0] += C.sum()
        m[1] += (C * C).sum()

    @_recorded
    def copy2d_scaled(self, dst, dst_ld, src, src_ld, rows, cols, alpha=1.0):
        self.launches += 1
        if rows <= 0 or cols <= 0:
            return
        s = _view(src, (rows - 1) * src_ld + cols)
        d = _view(dst, (rows - 1) * dst_ld + cols)
        S = numpy.lib.stride_tricks.as_strided(s, (rows, cols), (8 * src_ld, 8))
        D = numpy.lib.stride_tricks.as_strided(d, (rows, cols), (8 * dst_ld, 8))
        D[...] = alpha * S

    def density_tensor(self, ops, rho, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket, n_configs_ket, ket_masks,
                       n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate=False):
        """semantics of xr_density_tensor through the (test-only) Python restatement in oracle/density_oracle.py"""
        import os, sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import density_oracle as do
        self.launches += 1
        assert 2 * n_orbs <= 64 and 1 <= len(ops) <= 4
        dim = 2 * n_orbs

        class states(object):
            pass
        bra, ket = states(), states()
        bra.coeffs = _view(z_bra, n_bra_states * n_configs_bra).reshape(n_bra_states, n_configs_bra)
        bra.configs = numpy.zeros((n_configs_bra, n_elec_bra), dtype=numpy.int64)
        ket.coeffs = _view(z_ket, n_ket_states * n_configs_ket).reshape(n_ket_states, n_configs_ket)
        masks = _view(ket_masks, n_configs_ket, numpy.int64)
        ket.configs = numpy.array([[i for i in range(dim) if (int(m) >> i) & 1] for m in masks], dtype=numpy.int64).reshape(n_configs_ket, n_elec_ket)
        out = do.tensor(ops, {"bra": bra, "ket": ket}, "bra", "ket", n_orbs, n_core)
        if accumulate:
            _view(rho, out.size)[...] += out.reshape(-1)
        else:
            _view(rho, out.size)[...] = out.reshape(-1)

    def density_contracted(self, ops, out, weights, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket, n_configs_ket, ket_masks,
                           n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate=False):
        T = (2 * n_orbs) ** len(ops)
        rho = numpy.zeros(n_bra_states * n_ket_states * T)
        self.density_tensor(ops, rho.ctypes.data, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket, n_configs_ket, ket_masks,
                            n_elec_bra, n_elec_ket, n_orbs, n_core)
        res = rho.reshape(n_bra_states * n_ket_states, T) @ _view(weights, T)
        o = _view(out, n_bra_states * n_ket_states)
        o[...] = o + res if accumulate else res

    @_recorded
    def gemm_dd(self, M, N, K, A, lda, B, ldb, C0, ldc0, sign, out, ldo):
        self.launches += 1
        if M <= 0 or N <= 0:
            return
        a = numpy.lib.stride_tricks.as_strided(_view(A, (M - 1) * lda + K), (M, K), (8 * lda, 8)).astype(numpy.longdouble)
        b = numpy.lib.stride_tricks.as_strided(_view(B, (K - 1) * ldb + N), (K, N), (8 * ldb, 8)).astype(numpy.longdouble)
        if C0 is None:
            c = numpy.eye(M, N, dtype=numpy.longdouble)
        else:
            c = numpy.lib.stride_tricks.as_strided(_view(C0, (M - 1) * ldc0 + N), (M, N), (8 * ldc0, 8)).astype(numpy.longdouble)
        o = numpy.lib.stride_tricks.as_strided(_view(out, (M - 1) * ldo + N), (M, N), (8 * ldo, 8))
        o[...] = (c + sign * (a @ b)).astype(numpy.float64)

    @_recorded
    def embed_add(self, H, src, ld, R, Cn, S, offR, offC, offS=None, dims_sub=(), min_transitions=0, alpha=1.0):
        self.launches += 1
        if R <= 0 or Cn <= 0 or S <= 0:
            return
        assert S == 1 or offS is not None
        block = numpy.lib.stride_tricks.as_strided(_view(src, (R - 1) * ld + Cn), (R, Cn), (8 * ld, 8)).copy()
        if min_transitions > 0:
            assert 1 <= len(dims_sub) <= 4 and int(numpy.prod(dims_sub)) == R == Cn
            digits = numpy.indices(list(dims_sub)).reshape(len(dims_sub), -1)
            changed = (digits[:, :, None] != digits[:, None, :]).sum(axis=0)
            block[changed < min_transitions] = 0.0
        oR, oC = _view(offR, R, numpy.int64), _view(offC, Cn, numpy.int64)
        oS = _view(offS, S, numpy.int64) if offS is not None else numpy.zeros(1, dtype=numpy.int64)
        at = (oR[:, None, None] + oS[None, :, None] + oC[None, None, :])
        h = _view(H, int(at.max()) + 1)
        numpy.add.at(h, at.reshape(-1), numpy.broadcast_to(alpha * block[:, None, :], at.shape).reshape(-1))

    @_recorded
    def scatter_const(self, C, idx, count, value, accumulate=False):
        self.launches += 1
        if count <= 0:
            return
        at = _view(idx, count, numpy.int64)
        c = _view(C, int(at.max()) + 1)
        if accumulate:
            numpy.add.at(c, at, value)
        else:
            c[at] = value

    @_recorded
    def permute_copy(self, dst, src, shape, src_strides, alpha=1.0):
        self.launches += 1
        assert 1 <= len(shape) <= 12, "xr_permute_copy supports 1..12 dimensions"      # same limit as the CUDA kernel
        total = int(numpy.prod(shape))
        if total == 0:
            return
        span = 1 + sum((e - 1) * abs(s) for e, s in zip(shape, src_strides))
        s = _view(src, span)
        S = numpy.lib.stride_tricks.as_strided(s, tuple(shape), tuple(8 * x for x in src_strides))
        _view(dst, total)[...] = (alpha * S).reshape(-1)

    @_recorded
    def trimer_stream(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end, mode,
                      moments=None, C=None, offA=None, offB=None, offC=None):
        self.launches += 1
        if a_begin == a_end or Pb <= 0 or Pc <= 0:
            return
        w = numpy.lib.stride_tricks.as_strided(_view(W, (Pa - 1) * ldw + n * n), (Pa, n * n), (8 * ldw, 8))[a_begin:a_end]
        b = numpy.lib.stride_tricks.as_strided(_view(beta, (Pb - 1) * ldbeta + n), (Pb, n), (8 * ldbeta, 8))
        g = numpy.lib.stride_tricks.as_strided(_view(gamma, (Pc - 1) * ldgamma + n), (Pc, n), (8 * ldgamma, 8))
        T = alpha * numpy.einsum("ars,br,cs->abc", w.reshape(-1, n, n), b, g, optimize=True)
        if mode == 0:
            m = _view(moments, 2)
            m[0] += T.sum()
            m[1] += (T * T).sum()
        else:
            oa = _view(offA, Pa, numpy.int64)[a_begin:a_end]
            ob = _view(offB, Pb, numpy.int64)
            oc = _view(offC, Pc, numpy.int64)
            at = (oa[:, None, None] + ob[None, :, None] + oc[None, None, :]).reshape(-1)
            _view(C, int(at.max()) + 1)[at] = T.reshape(-1)

    def _trimer_T(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end):
        w = numpy.lib.stride_tricks.as_strided(_view(W, (Pa - 1) * ldw + n * n), (Pa, n * n), (8 * ldw, 8))[a_begin:a_end]
        b = numpy.lib.stride_tricks.as_strided(_view(beta, (Pb - 1) * ldbeta + n), (Pb, n), (8 * ldbeta, 8))
        g = numpy.lib.stride_tricks.as_strided(_view(gamma, (Pc - 1) * ldgamma + n), (Pc, n), (8 * ldgamma, 8))
        return alpha * numpy.einsum("ars,br,cs->abc", w.reshape(-1, n, n), b, g, optimize=True)

    def trimer_threshold(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end, tau, offA, offB, offC,
                         capacity, idx_out, val_out, count):
        self.launches += 1
        cnt = _view(count, 1, numpy.int64)
        cnt[0] = 0
        if a_begin == a_end or Pb <= 0 or Pc <= 0:
            return
        T = self._trimer_T(n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end)
        oa = _view(offA, Pa, numpy.int64)[a_begin:a_end]
        at = oa[:, None, None] + _view(offB, Pb, numpy.int64)[None, :, None] + _view(offC, Pc, numpy.int64)[None, None, :]
        keep = numpy.abs(T) > tau
        cnt[0] = int(keep.sum())
        stored = min(int(cnt[0]), capacity)
        if stored:
            _view(idx_out, stored, numpy.int64)[...] = at[keep][:stored]
            _view(val_out, stored)[...] = T[keep][:stored]

    def trimer_sample(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, abc_host, out):
        self.launches += 1
        if len(abc_host) == 0:
            return
        T = self._trimer_T(n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, 0, Pa)
        abc = numpy.asarray(abc_host)
        _view(out, len(abc))[...] = T[abc[:, 0], abc[:, 1], abc[:, 2]]


class FakeDevice(object):
    """same surface as qodeapplications_b200.device.Device, on CPU torch tensors"""
    def __init__(self, index=0):
        self.index = index
        self.torch_device = torch.device("cpu")
        self.ctx = FakeContext()
        self.h2d_bytes = self.d2h_bytes = 0
        self.largest_allocation = 0          # elements of the biggest tensor ever created on this "device"
        self._keep = None

    def begin_trace(self):
        self._keep = []
        self.ctx.begin_trace()

    def end_trace(self):
        keep, self._keep = self._keep, None
        return self.ctx.end_trace(), keep

    @property
    def tracing(self):
        return self._keep is not None

    def _made(self, tensor):
        self.largest_allocation = max(self.largest_allocation, tensor.numel())
        if self._keep is not None:
            self._keep.append(tensor)
        return tensor

    def empty(self, shape, dtype=torch.float64):
        return self._made(torch.zeros(shape, dtype=dtype))

    def zeros(self, shape, dtype=torch.float64):
        out = self._made(torch.zeros(shape, dtype=dtype))
        if self._keep is not None and out.numel():
            self.ctx.memset_zero(out, out.numel() * out.element_size())
        return out

    def upload(self, array, dtype=numpy.float64):
        array = numpy.array(array, dtype=dtype, order="C", copy=True)
        self.h2d_bytes += array.nbytes
        return self._made(torch.from_numpy(array))

    def download(self, tensor):
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.numpy().copy()

    def sync(self):
        pass
