"""qode.util.PyC stand-in: import_C compiles <caller dir>/<name>.c (where it lies) with
the shim PyC_types.h and exposes its functions through ctypes the way build_H.py uses
them (general-XRCC/build_H.py:18-29): attributes are callables with .return_type(float);
ints go in as int64, C-contiguous float64 ndarrays as double*."""
import ctypes
import hashlib
import inspect
import os
import subprocess
import tempfile
import numpy

_SHIM_INCLUDE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "shim")


class Double(object):
    numpy = numpy.float64
    ctypes = ctypes.c_double
    tensorly = numpy    # XRbase/XR_tensor.py:54 passes this as a dtype


class BigInt(object):
    numpy = numpy.int64
    ctypes = ctypes.c_int64


class _function(object):
    def __init__(self, cfunc):
        self._cfunc = cfunc
        self._cfunc.restype = ctypes.c_int64
    def return_type(self, py_type):
        self._cfunc.restype = {float: ctypes.c_double, int: ctypes.c_int64}[py_type]
    def __call__(self, *args):
        converted = []
        for arg in args:
            if isinstance(arg, numpy.ndarray):
                if arg.dtype == numpy.float64:
                    arg = numpy.ascontiguousarray(arg)
                    converted.append(arg.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
                else:
                    arg = numpy.ascontiguousarray(arg, dtype=numpy.int64)
                    converted.append(arg.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
            elif isinstance(arg, (list, tuple)):      # list of per-sector arrays -> Double*[] / BigInt*[] (build_density_tensors.py:88)
                arrays = [numpy.ascontiguousarray(a) for a in arg]
                ctype = ctypes.c_double if arrays[0].dtype == numpy.float64 else ctypes.c_int64
                self._keep = arrays
                converted.append((ctypes.POINTER(ctype) * len(arrays))(*[a.ctypes.data_as(ctypes.POINTER(ctype)) for a in arrays]))
            elif isinstance(arg, (int, numpy.integer)):
                converted.append(ctypes.c_int64(int(arg)))
            elif isinstance(arg, (float, numpy.floating)):
                converted.append(ctypes.c_double(float(arg)))
            else:
                raise TypeError("cannot pass {} to C".format(type(arg)))
        return self._cfunc(*converted)


class _module(object):
    def __init__(self, lib):
        self._lib = lib
        self._functions = {}
    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._functions:
            self._functions[name] = _function(getattr(self._lib, name))
        return self._functions[name]


def import_C(name, flags="", include=None, cc="gcc"):
    caller_dir = os.path.dirname(os.path.abspath(inspect.stack()[1].filename))
    source = os.path.join(caller_dir, name + ".c")
    tag = hashlib.sha1((source + flags).encode()).hexdigest()[:12]
    out = os.path.join(tempfile.gettempdir(), "qode_shim_{}_{}.so".format(name, tag))
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(source):
        cmd = [cc] + flags.split() + ["-shared", "-fPIC", "-I", _SHIM_INCLUDE, "-o", out, source]
        subprocess.check_call(cmd)
    return _module(ctypes.CDLL(out))
