"""ctypes binding of libxr_b200.so (include/xr_b200.h) -- the only way Python reaches the GPU here.

There is NO CPU fallback: if the shared library is missing, or no sm_100 device is present,
``load()`` / ``Context()`` raise.  Importing this module does not need a GPU (the CPU test-suite
checks that the library loads and exports every declared symbol).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# XR_B200_LIB: load another BUILD of the same library (tools/trimer_variants.py compiles kernel variants side by side)
LIB_PATH = os.environ.get("XR_B200_LIB") or os.path.join(HERE, "libxr_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "xr_b200.h")

_i64 = ctypes.c_int64
_dbl = ctypes.c_double
_ptr = ctypes.c_void_p
_int = ctypes.c_int

LEGACY_SYMBOLS = {      # name -> (number of PyInt args, number of Double* args); H_contractions.c ABI
    "monomer": (1, 4), "monomer_1e": (1, 2), "monomer_2e": (1, 2), "monomer_extPot": (1, 2),
    "dimer_2min2pls": (2, 3), "dimer_1min1pls_1e": (2, 3), "dimer_1min1pls_2e": (2, 6), "dimer_ExEx": (2, 3),
    "trimer_2min1pls1pls": (3, 4), "trimer_2pls1min1min": (3, 4), "trimer_Ex1min1pls": (3, 4),
}

# density_tensors.c ABI: (storage, bra, ket, n_elec[], n_states[], z_list[], n_configs[], configs[], n_orbs, n_core, combinatorics[], n_threads)
LEGACY_DENSITY_SYMBOLS = ("a_tensor", "c_tensor", "ca_tensor", "aa_tensor", "cc_tensor", "caa_tensor", "cca_tensor", "ccaa_tensor")

_PROTOTYPES = {
    "xr_last_error": (ctypes.c_char_p, []),
    "xr_version": (ctypes.c_char_p, []),
    "xr_ctx_create": (_int, [_int, _ptr, _int, ctypes.POINTER(_ptr)]),
    "xr_ctx_destroy": (_int, [_ptr]),
    "xr_ctx_set_stream": (_int, [_ptr, _ptr]),
    "xr_sync": (_int, [_ptr]),
    "xr_launch_count": (_int, [_ptr, ctypes.POINTER(_i64)]),
    "xr_device_info": (_int, [_ptr, ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int),
                              ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "xr_probe_fp64": (_int, [_ptr, _dbl, ctypes.POINTER(_dbl)]),
    "xr_malloc": (_int, [_ptr, ctypes.c_size_t, ctypes.POINTER(_ptr)]),
    "xr_free": (_int, [_ptr, _ptr]),
    "xr_memset_zero": (_int, [_ptr, _ptr, ctypes.c_size_t]),
    "xr_upload": (_int, [_ptr, _ptr, _ptr, ctypes.c_size_t]),
    "xr_download": (_int, [_ptr, _ptr, _ptr, ctypes.c_size_t]),
    "xr_gemm_scatter": (_int, [_ptr, _i64, _i64, _i64, _dbl, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _int]),
    "xr_gemm_stream": (_int, [_ptr, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _dbl, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _int]),
    "xr_gemm_reduce": (_int, [_ptr, _i64, _i64, _i64, _dbl, _ptr, _i64, _ptr, _i64, _ptr]),
    "xr_copy2d_scaled": (_int, [_ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _dbl]),
    "xr_scatter_const": (_int, [_ptr, _ptr, _ptr, _i64, _dbl, _int]),
    "xr_density_tensor": (_int, [_ptr, ctypes.c_char_p, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _int]),
    "xr_density_contracted": (_int, [_ptr, ctypes.c_char_p, _ptr, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _int]),
    "xr_gemm_dd": (_int, [_ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _dbl, _ptr, _i64]),
    "xr_embed_add": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr, _int, _ptr, _int, _dbl]),
    "xr_permute_copy": (_int, [_ptr, _ptr, _ptr, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64), _dbl]),
    "xr_trimer_stream": (_int, [_ptr, _int, _i64, _i64, _i64, _dbl, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _int,
                                _ptr, _ptr, _ptr, _ptr, _ptr]),
    "xr_trimer_threshold": (_int, [_ptr, _int, _i64, _i64, _i64, _dbl, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _dbl,
                                   _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "xr_trimer_sample": (_int, [_ptr, _int, _i64, _i64, _i64, _dbl, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _ptr, _ptr]),
}

_lib = None


class XRError(RuntimeError):
    pass


def declared_symbols(header_path=HEADER_PATH):
    """Every function name include/xr_b200.h declares (used by the CPU test that the ABI is complete)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text)
    return sorted(set(n for n in names if n not in ("defined",)))


def load():
    """dlopen libxr_b200.so and attach prototypes.  Raises (loudly) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XRError("%s not found: build it with `python -m qodeapplications_b200.build` "
                      "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    for name, (n_int, n_ptr) in LEGACY_SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = _dbl
        fn.argtypes = [_i64] * n_int + [_ptr] * n_ptr
    for name in LEGACY_DENSITY_SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = None
        fn.argtypes = [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _ptr, _i64]
    _lib = lib
    return lib


def last_error():
    msg = load().xr_last_error()
    return msg.decode() if msg else ""


def check(rc, what=""):
    if rc != 0:
        raise XRError("%s failed (status %d): %s" % (what or "xr call", rc, last_error()))


def _p(x):
    """device/host pointer from int, None, torch tensor or numpy array"""
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return ctypes.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):
        return ctypes.c_void_p(x.ctypes.data)
    raise TypeError("cannot take a pointer of %r" % type(x))


def _recorded(method):
    """kernel method of a Context: while a trace is being recorded (Context.begin_trace) the call is also appended to it,
    so that the whole launch sequence of a build can be replayed without the Python planning that produced it"""
    def call(self, *args, **kwargs):
        if self._trace is not None:
            self._trace.append((call, args, kwargs))
        return method(self, *args, **kwargs)
    call.__name__, call.__doc__ = method.__name__, method.__doc__
    return call


class Context(object):
    """One xr_ctx: a device plus the CUDA stream all xr kernels of this context are launched on."""
    def __init__(self, device=0, stream=None, own_stream=False):
        """stream: a cudaStream_t handle (int) to borrow; None/0 is the CUDA default stream.
        own_stream=True lets the context create its own non-blocking stream instead."""
        self.lib = load()
        handle = ctypes.c_void_p()
        check(self.lib.xr_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None,
                                     1 if own_stream else 0, ctypes.byref(handle)), "xr_ctx_create")
        self.handle = handle
        self.device = int(device)
        self._trace = None

    # ---- launch-sequence recording (hermitian/plan.py): the data-independent launch list of a build, replayable as is
    def begin_trace(self):
        if self._trace is not None:
            raise XRError("a trace is already being recorded on this context")
        self._trace = []

    def end_trace(self):
        trace, self._trace = self._trace, None
        return trace

    def replay(self, trace):
        """re-issue a recorded launch sequence (same device buffers, same arguments) on the context's current stream"""
        for call, args, kwargs in trace:
            call(self, *args, **kwargs)

    @_recorded
    def memset_zero(self, ptr, nbytes):
        check(self.lib.xr_memset_zero(self.handle, _p(ptr), int(nbytes)), "xr_memset_zero")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.xr_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        check(self.lib.xr_ctx_set_stream(self.handle, ctypes.c_void_p(stream) if stream else None), "xr_ctx_set_stream")

    def sync(self):
        check(self.lib.xr_sync(self.handle), "xr_sync")

    def launch_count(self):
        n = ctypes.c_int64()
        check(self.lib.xr_launch_count(self.handle, ctypes.byref(n)), "xr_launch_count")
        return n.value

    def device_info(self):
        sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        free, total = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.lib.xr_device_info(self.handle, ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor),
                                      ctypes.byref(free), ctypes.byref(total)), "xr_device_info")
        return dict(sm_count=sm.value, cc=(major.value, minor.value), free_bytes=free.value, total_bytes=total.value)

    def probe_fp64(self, seconds=0.5):
        """sustained DMMA.8x8x4 rate of this device in TFLOP/s, measured now (the tensor-roofline denominator)"""
        out = ctypes.c_double()
        check(self.lib.xr_probe_fp64(self.handle, float(seconds), ctypes.byref(out)), "xr_probe_fp64")
        return out.value

    # ---- kernels -------------------------------------------------------------------------------
    @_recorded
    def gemm_scatter(self, M, N, K, alpha, A, lda, B, ldb, C, offM=None, ldc=0, offN=None, accumulate=False):
        check(self.lib.xr_gemm_scatter(self.handle, M, N, K, float(alpha), _p(A), lda, _p(B), ldb, _p(C), _p(offM),
                                       ldc, _p(offN), 1 if accumulate else 0), "xr_gemm_scatter")

    @_recorded
    def gemm_stream(self, E1, s1, E2, s2, EK1, sk1, EK2, N, alpha, A, B, ldb, C, offM=None, ldc=0, offN=None, accumulate=False):
        """returns False (nothing launched) when the operands cannot be addressed by the stream kernel"""
        rc = self.lib.xr_gemm_stream(self.handle, E1, s1, E2, s2, EK1, sk1, EK2, N, float(alpha), _p(A), _p(B), ldb, _p(C), _p(offM),
                                     ldc, _p(offN), 1 if accumulate else 0)
        if rc == -4:            # XR_ERR_UNSUPPORTED
            return False
        check(rc, "xr_gemm_stream")
        return True

    @_recorded
    def gemm_reduce(self, M, N, K, alpha, A, lda, B, ldb, moments):
        check(self.lib.xr_gemm_reduce(self.handle, M, N, K, float(alpha), _p(A), lda, _p(B), ldb, _p(moments)), "xr_gemm_reduce")

    @_recorded
    def copy2d_scaled(self, dst, dst_ld, src, src_ld, rows, cols, alpha=1.0):
        check(self.lib.xr_copy2d_scaled(self.handle, _p(dst), dst_ld, _p(src), src_ld, rows, cols, float(alpha)),
              "xr_copy2d_scaled")

    @_recorded
    def permute_copy(self, dst, src, shape, src_strides, alpha=1.0):
        nd = len(shape)
        arr = _i64 * nd
        check(self.lib.xr_permute_copy(self.handle, _p(dst), _p(src), nd, arr(*[int(x) for x in shape]),
                                       arr(*[int(x) for x in src_strides]), float(alpha)), "xr_permute_copy")

    def density_tensor(self, ops, rho, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket, n_configs_ket, ket_masks,
                       n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate=False):
        check(self.lib.xr_density_tensor(self.handle, ops.encode(), _p(rho), n_bra_states, n_ket_states, _p(z_bra), n_configs_bra,
                                         _p(z_ket), n_configs_ket, _p(ket_masks), n_elec_bra, n_elec_ket, n_orbs, n_core,
                                         1 if accumulate else 0),
              "xr_density_tensor")

    def density_contracted(self, ops, out, weights, n_bra_states, n_ket_states, z_bra, n_configs_bra, z_ket, n_configs_ket, ket_masks,
                           n_elec_bra, n_elec_ket, n_orbs, n_core, accumulate=False):
        check(self.lib.xr_density_contracted(self.handle, ops.encode(), _p(out), _p(weights), n_bra_states, n_ket_states, _p(z_bra),
                                             n_configs_bra, _p(z_ket), n_configs_ket, _p(ket_masks), n_elec_bra, n_elec_ket, n_orbs,
                                             n_core, 1 if accumulate else 0), "xr_density_contracted")

    @_recorded
    def gemm_dd(self, M, N, K, A, lda, B, ldb, C0, ldc0, sign, out, ldo):
        check(self.lib.xr_gemm_dd(self.handle, M, N, K, _p(A), lda, _p(B), ldb, _p(C0), ldc0, float(sign), _p(out), ldo), "xr_gemm_dd")

    @_recorded
    def embed_add(self, H, src, ld, R, Cn, S, offR, offC, offS=None, dims_sub=(), min_transitions=0, alpha=1.0):
        dims = (ctypes.c_int64 * max(1, len(dims_sub)))(*[int(d) for d in dims_sub])
        check(self.lib.xr_embed_add(self.handle, _p(H), _p(src), ld, R, Cn, S, _p(offR), _p(offC), _p(offS), len(dims_sub),
                                    dims, int(min_transitions), float(alpha)), "xr_embed_add")

    @_recorded
    def scatter_const(self, C, idx, count, value, accumulate=False):
        check(self.lib.xr_scatter_const(self.handle, _p(C), _p(idx), count, float(value), 1 if accumulate else 0),
              "xr_scatter_const")

    @_recorded
    def trimer_stream(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end, mode,
                      moments=None, C=None, offA=None, offB=None, offC=None):
        check(self.lib.xr_trimer_stream(self.handle, n, Pa, Pb, Pc, float(alpha), _p(W), ldw, _p(beta), ldbeta,
                                        _p(gamma), ldgamma, a_begin, a_end, mode, _p(moments), _p(C), _p(offA),
                                        _p(offB), _p(offC)), "xr_trimer_stream")

    def trimer_threshold(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, a_begin, a_end, tau, offA, offB, offC,
                         capacity, idx_out, val_out, count):
        check(self.lib.xr_trimer_threshold(self.handle, n, Pa, Pb, Pc, float(alpha), _p(W), ldw, _p(beta), ldbeta, _p(gamma), ldgamma,
                                           a_begin, a_end, float(tau), _p(offA), _p(offB), _p(offC), capacity, _p(idx_out),
                                           _p(val_out), _p(count)), "xr_trimer_threshold")

    def trimer_sample(self, n, Pa, Pb, Pc, alpha, W, ldw, beta, ldbeta, gamma, ldgamma, abc_host, out):
        """abc_host: C-contiguous int64 ndarray [count, 3] on the HOST; out: device doubles [count]"""
        check(self.lib.xr_trimer_sample(self.handle, n, Pa, Pb, Pc, float(alpha), _p(W), ldw, _p(beta), ldbeta, _p(gamma), ldgamma,
                                        len(abc_host), _p(abc_host), _p(out)), "xr_trimer_sample")


TRIMER_REDUCE = 0
TRIMER_MATERIALIZE = 1
