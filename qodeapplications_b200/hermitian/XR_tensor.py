"""The tensor seam of the reference (XRbase/XR_tensor.py:53-58): ``init`` wraps primitive data,
``raw`` evaluates an index-labelled expression to an ndarray.

The expression syntax of qode.math.tensornet that the reference's diagram files are written in is kept:

    raw( A(0, 1, "p", "q") @ B(2, 3, "q", "p") )         # ints = free axes of the result, in ascending order;
    2.0 * raw( A(0, 1, "p") @ B(2, 3, "q") @ S("p", "q") )   strings = contracted labels (each appears twice)

but the product is evaluated on the GPU: the operands are uploaded once (tensor.DeviceStore) and contracted
pairwise by xr_gemm_scatter (tensor.Contractor.multi_contract).
"""
import numpy

from .tensor import Contractor, DeviceStore, DeviceTensor, as_host, default_device

_engine = {}


def _default_engine():
    dev = default_device()
    if dev not in _engine:
        _engine[dev] = (DeviceStore(dev), Contractor(dev))
    return _engine[dev]


class _product(object):
    """scalar * product of index-labelled tensors (a tensornet expression)"""
    def __init__(self, factors, scalar=1.0):
        self.factors, self.scalar = factors, scalar
    def __matmul__(self, other):
        other = other() if isinstance(other, xr_tensor) else other
        return _product(self.factors + other.factors, self.scalar * other.scalar)
    def __mul__(self, scalar):
        return _product(self.factors, self.scalar * scalar)
    __rmul__ = __mul__
    def __neg__(self):
        return _product(self.factors, -self.scalar)


class xr_tensor(object):
    def __init__(self, array):
        self.array = numpy.ascontiguousarray(array, dtype=numpy.float64)
    @property
    def shape(self):
        return self.array.shape
    def __call__(self, *labels):
        if len(labels) != self.array.ndim:
            raise ValueError("tensor of rank %d indexed with %d labels" % (self.array.ndim, len(labels)))
        return _product([(self, tuple(labels))])
    def __matmul__(self, other):
        return self() @ other


def init(raw_tensor):
    return xr_tensor(raw_tensor)


def zeros():
    return xr_tensor(numpy.zeros(()))


def evaluate(expr, engine=None):
    """device tensor of an expression; free (int) labels in ascending order"""
    store, contractor = engine or _default_engine()
    factors, free = [], set()
    for tensor, labels in expr.factors:
        if len(set(labels)) != len(labels):
            raise NotImplementedError("a label repeated inside one tensor (partial trace)")
        factors.append((store.get(tensor), list(labels)))
        free.update(l for l in labels if isinstance(l, (int, numpy.integer)))
    return contractor.multi_contract(factors, sorted(free), alpha=expr.scalar)


def raw(tensor, engine=None):
    if isinstance(tensor, _product):
        return evaluate(tensor, engine).host()
    if isinstance(tensor, DeviceTensor):
        return tensor.host()
    return as_host(tensor)
