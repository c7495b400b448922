"""Drop-in for what ``qode.util.PyC.import_C("H_contractions", flags="-O2")`` returns in the
reference (general-XRCC/build_H.py:20-29): an object whose attributes are the eleven contraction
functions, each callable with (ints..., float64 ndarrays...) and carrying ``.return_type``.
Here the attributes are bound to the GPU implementations exported, under the same names and
C ABI, by libxr_b200.so (include/xr_b200.h group A)."""
import ctypes
import numpy

from .. import lib as _lib


class _function(object):
    def __init__(self, name, cfunc, n_int, n_ptr):
        self._name, self._cfunc, self._n_int, self._n_ptr = name, cfunc, n_int, n_ptr
    def return_type(self, py_type):
        if py_type is not float:
            raise TypeError("%s returns a C double" % self._name)
    def __call__(self, *args):
        if len(args) != self._n_int + self._n_ptr:
            raise TypeError("%s takes %d arguments" % (self._name, self._n_int + self._n_ptr))
        conv, keep = [int(a) for a in args[:self._n_int]], []
        for a in args[self._n_int:]:
            a = numpy.ascontiguousarray(a, dtype=numpy.float64)
            keep.append(a)
            conv.append(ctypes.c_void_p(a.ctypes.data))
        value = self._cfunc(*conv)
        if value != value:
            raise _lib.XRError("%s failed on the GPU: %s" % (self._name, _lib.last_error()))
        return value


class _module(object):
    def __init__(self):
        lib = _lib.load()
        for name, (n_int, n_ptr) in _lib.LEGACY_SYMBOLS.items():
            setattr(self, name, _function(name, getattr(lib, name), n_int, n_ptr))


_contract = None


def import_C(name="H_contractions", flags=""):
    global _contract
    if name != "H_contractions":
        raise ImportError("only H_contractions is provided by libxr_b200")
    if _contract is None:
        _contract = _module()
    return _contract
