"""qode.math stand-ins (hermitian-XRCC/get_xr_result.py:165,246)."""
import numpy
from . import tensornet


def precise_numpy_inverse(M):
    """Inverse refined by one Newton-Schulz step in extended precision (the real Qode
    routine's exact recipe is not visible from /root/reference; any inverse accurate to
    double precision satisfies the call site)."""
    M = numpy.asarray(M, dtype=numpy.float64)
    X = numpy.linalg.inv(M)
    ML = M.astype(numpy.longdouble)
    XL = X.astype(numpy.longdouble)
    R = numpy.eye(M.shape[0], dtype=numpy.longdouble) - ML @ XL
    XL = XL + XL @ R
    return numpy.asarray(XL, dtype=numpy.float64)
