"""The lowest seam of the reference: the backend function table Qode's tensornet takes
(``tensornet.primitive_tensor_factory(functions)``, XRbase/meta_backend.py:27-72: shape, scalar_value, copy_data, zeros,
scalar_tensor, increment, mult, element, str, contract).  With Qode on the path,

    import qode.math.tensornet as tensornet
    from qodeapplications_b200.hermitian.meta_backend import xr_functions, xr_wrapper
    xr_tensor = tensornet.primitive_tensor_factory(xr_functions)          # as XRbase/meta_backend.py:72 does
    A = xr_tensor.init(xr_wrapper.from_host(ndarray))                      # data live in HBM from here on

gives tensornet tensors whose every contraction runs through xr_gemm_scatter / xr_permute_copy on the GPU.  Nothing here
imports Qode; the table only follows the calling convention visible at XRbase/meta_backend.py:29-69:
``contract(*factors)`` receives scalars and ``(tensor, *indices)`` tuples, integer indices are the free axes of the
result (ascending), any other label is summed.
"""
import numpy

from .tensor import Contractor, DeviceTensor, default_device

_contractors = {}


def _contractor(dev):
    if id(dev) not in _contractors:
        _contractors[id(dev)] = Contractor(dev)
    return _contractors[id(dev)]


class xr_wrapper(object):
    """a device tensor as tensornet's primitive (the role of meta_wrapper, XRbase/meta_backend.py:22-24)"""
    def __init__(self, internal):
        self.internal = internal          # tensor.DeviceTensor
    @staticmethod
    def from_host(array, dev=None):
        dev = dev or default_device()
        return xr_wrapper(DeviceTensor(dev.upload(numpy.ascontiguousarray(array, dtype=numpy.float64)), dev))
    def host(self):
        return self.internal.host()


class xr_functions(object):
    @staticmethod
    def shape(tensor):
        return tuple(tensor.internal.shape)
    @staticmethod
    def scalar_value(tensor):
        if tensor.internal.buf.numel() != 1:
            raise ValueError("scalar_value of a tensor with shape %r" % (tuple(tensor.internal.shape),))
        return float(tensor.internal.host().reshape(-1)[0])
    @staticmethod
    def copy_data(tensor):
        src = tensor.internal
        out = src.dev.empty(tuple(src.shape))
        n = int(numpy.prod(src.shape)) if src.shape else 1
        src.dev.ctx.copy2d_scaled(out, n, src.buf, n, 1, n, 1.0)
        return xr_wrapper(DeviceTensor(out, src.dev))
    @staticmethod
    def zeros(shape=()):
        dev = default_device()
        return xr_wrapper(DeviceTensor(dev.zeros(tuple(shape)), dev))
    @staticmethod
    def scalar_tensor(scalar):
        return xr_wrapper.from_host(numpy.array(float(scalar)))
    @staticmethod
    def increment(tensor, delta):
        """in place: tensor += delta (XRbase/meta_backend.py:44-49)"""
        dst, src = tensor.internal, delta.internal
        if tuple(dst.shape) != tuple(src.shape):
            raise ValueError("increment: shapes %r and %r differ" % (tuple(dst.shape), tuple(src.shape)))
        n = int(numpy.prod(dst.shape)) if dst.shape else 1
        one = _contractor(dst.dev).dev.upload(numpy.ones((1, 2)))
        # dst[i] += src[i] * 1: a rank-1 GEMM with the accumulate epilogue
        dst.dev.ctx.gemm_scatter(n, 1, 1, 1.0, src.buf, 1, one, 2, dst.buf, None, 1, None, True)
    @staticmethod
    def mult(scalar, tensor):
        src = tensor.internal
        out = src.dev.empty(tuple(src.shape))
        n = int(numpy.prod(src.shape)) if src.shape else 1
        src.dev.ctx.copy2d_scaled(out, n, src.buf, n, 1, n, float(scalar))
        return xr_wrapper(DeviceTensor(out, src.dev))
    @staticmethod
    def element(tensor, indices):
        return float(tensor.internal.buf[tuple(indices)].item())
    @staticmethod
    def str(tensor):
        return "xr_b200 device tensor, shape %r" % (tuple(tensor.internal.shape),)
    @staticmethod
    def contract(*tensor_factors):
        scalar, factors, free = 1.0, [], set()
        for factor in tensor_factors:
            try:
                tensor, *indices = factor
            except TypeError:
                scalar *= factor
            else:
                if len(set(indices)) != len(indices):
                    raise NotImplementedError("a label repeated inside one tensor (partial trace)")
                factors.append((tensor.internal, list(indices)))
                free.update(i for i in indices if isinstance(i, (int, numpy.integer)))
        if not factors:
            return xr_functions.scalar_tensor(scalar)
        dev = factors[0][0].dev
        return xr_wrapper(_contractor(dev).multi_contract(factors, sorted(free), alpha=scalar))
