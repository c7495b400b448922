"""The tensor seam of the reference (XRbase/XR_tensor.py:53-58): ``init`` wraps primitive data,
``raw`` evaluates an index-labelled expression to an ndarray.

The expression syntax of qode.math.tensornet that the reference's diagram files are written in is kept:

    raw( A(0, 1, "p", "q") @ B(2, 3, "q", "p") )         # ints = free axes of the result, in ascending order;
    2.0 * raw( A(0, 1, "p") @ B(2, 3, "q") @ S("p", "q") )   strings = contracted labels (each appears twice)

but the product is evaluated on the GPU: the operands are uploaded once (tensor.DeviceStore) and contracted
pairwise by xr_gemm_scatter (tensor.Contractor.multi_contract).  Sums of such products are supported the way the
reference's density code writes them (frag-states/compress_frags.py:91-99):

    temp = XR_tensor.zeros()                      # takes its shape from the summed terms
    temp += rho(0, 1, 2, 3);  temp -= rho(1, 0, 2, 3)
    raw(temp)  /  temp(0, 1, "p", "q") @ ...      # the terms are accumulated into ONE buffer by the GEMM epilogues

Expressions stay LAZY, as in tensornet: calling a product or a sum with labels re-indexes its free axes (free axis k, in
ascending order of the current integer labels, gets labels[k]) without evaluating anything, and `@` distributes over sums.
That is what lets the decomposed high-rank densities of frag-states/decomps.py:71-166 -- antisymmetrised sums of OUTER
products such as  anti(caC(p,u) @ cccaV(i,j,q,r,s,t), ...)  -- be handed to the diagram engine as they are: a density or
integral entry may be such an expression, and every diagram then contracts its factors directly (tensor.FactoredTensor);
the n^6..n^8 tensors they stand for are never formed.
"""
import numpy

from .tensor import Contractor, DeviceStore, DeviceTensor, as_host, default_device

_engine = {}


def _default_engine():
    dev = default_device()
    if dev not in _engine:
        _engine[dev] = (DeviceStore(dev), Contractor(dev))
    return _engine[dev]


_unique = [0]


def _relabelled(factors, free, labels):
    """factors with free axis free[k] renamed to labels[k]; the product's own contracted (string) labels are made unique
    so that they cannot collide with string labels the caller brings"""
    if len(labels) != len(free):
        raise ValueError("expression with %d free axes indexed with %d labels" % (len(free), len(labels)))
    _unique[0] += 1
    rename = dict(zip(free, labels))
    def new(l):
        if isinstance(l, (int, numpy.integer)):
            return rename[l]
        return l if str(l).startswith("_x") else "_x%d_%s" % (_unique[0], l)
    return [(tensor, tuple(new(l) for l in ls)) for tensor, ls in factors]


class _product(object):
    """scalar * product of index-labelled tensors (a tensornet expression)"""
    def __init__(self, factors, scalar=1.0):
        self.factors, self.scalar = factors, scalar
    def __matmul__(self, other):
        other = other() if isinstance(other, xr_tensor) else other
        if isinstance(other, _sum):
            return _sum([self @ t for t in other.terms])
        return _product(self.factors + other.factors, self.scalar * other.scalar)
    def __mul__(self, scalar):
        return _product(self.factors, self.scalar * scalar)
    __rmul__ = __mul__
    def __neg__(self):
        return _product(self.factors, -self.scalar)
    def __call__(self, *labels):
        """the same product with its free axes re-indexed (lazy, like a tensornet tensor)"""
        return _product(_relabelled(self.factors, _free(self), labels), self.scalar)
    @property
    def shape(self):
        extent = {}
        for tensor, labels in self.factors:
            extent.update(zip(labels, tensor.shape))
        return tuple(extent[l] for l in _free(self))


class _sum(object):
    """sum of tensornet expressions with the same free (integer) labels; evaluated into one device buffer"""
    def __init__(self, terms):
        self.terms = list(terms)
    def __add__(self, other):
        return _sum(self.terms + _terms(other))
    __iadd__ = __add__
    def __sub__(self, other):
        return _sum(self.terms + [-t for t in _terms(other)])
    __isub__ = __sub__
    def __mul__(self, scalar):
        return _sum([t * scalar for t in self.terms])
    __rmul__ = __mul__
    def __neg__(self):
        return _sum([-t for t in self.terms])
    def __call__(self, *labels):
        """the same sum with the free axes of every term re-indexed (lazy)"""
        return _sum([t(*labels) for t in self.terms])
    def __matmul__(self, other):
        other = other() if isinstance(other, xr_tensor) else other
        return _sum([t @ other for t in self.terms])
    @property
    def shape(self):
        return self.terms[0].shape


def _terms(expr):
    if isinstance(expr, _sum):
        return list(expr.terms)
    if isinstance(expr, xr_tensor):
        return [] if expr.is_empty_sum else [expr(*range(len(expr.shape)))]
    return [expr]


def _free(product):
    return sorted({l for _, labels in product.factors for l in labels if isinstance(l, (int, numpy.integer))})


_product.__add__ = lambda self, other: _sum([self] + _terms(other))
_product.__sub__ = lambda self, other: _sum([self] + [-t for t in _terms(other)])


class _device_tensor(object):
    """an evaluated expression (tensor.DeviceTensor) that can be indexed again"""
    def __init__(self, device_tensor):
        self.device_tensor = device_tensor
    @property
    def shape(self):
        return self.device_tensor.shape
    def __call__(self, *labels):
        if len(labels) != len(self.shape):
            raise ValueError("tensor of rank %d indexed with %d labels" % (len(self.shape), len(labels)))
        return _product([(self.device_tensor, tuple(labels))])
    def __getitem__(self, key):
        """basic slicing (ints, slices, Ellipsis) of an evaluated tensor: a contiguous copy in HBM, indexable again"""
        sliced = self.device_tensor.buf[key].contiguous()
        return _device_tensor(DeviceTensor(sliced, self.device_tensor.dev))


class xr_tensor(object):
    is_empty_sum = False
    def __add__(self, other):
        return _sum(_terms(self) + _terms(other))
    __iadd__ = __add__
    def __sub__(self, other):
        return _sum(_terms(self) + [-t for t in _terms(other)])
    __isub__ = __sub__
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
    def __getitem__(self, key):
        """basic slicing of primitive data, as the reference's callers write it:  v_bior[:, :, sl1, :](p, q, 1, s) @ ...
        (StateSpaceOptimizer/orb_grads.py:67)"""
        return xr_tensor(self.array[key])


def init(raw_tensor):
    return xr_tensor(raw_tensor)


def zeros():
    """the empty sum: takes its shape from the terms added to it (XRbase/XR_tensor.py:55-56)"""
    out = xr_tensor(numpy.zeros(()))
    out.is_empty_sum = True
    return out


def evaluate(expr, engine=None):
    """device tensor of an expression; free (int) labels in ascending order"""
    if isinstance(expr, _sum):
        if not expr.terms:
            raise ValueError("cannot evaluate an empty sum")
        out = evaluate(expr.terms[0], engine)
        store, contractor = engine or _default_engine()
        for term in expr.terms[1:]:
            if _free(term) != _free(expr.terms[0]):
                raise ValueError("terms of a sum must have the same free indices")
            factors = [(store.get(tensor), list(labels)) for tensor, labels in term.factors]
            contractor.multi_contract(factors, _free(term), alpha=term.scalar, out=out, accumulate=True)
        return out
    store, contractor = engine or _default_engine()
    factors, free = [], set()
    for tensor, labels in expr.factors:
        if len(set(labels)) != len(labels):
            raise NotImplementedError("a label repeated inside one tensor (partial trace)")
        factors.append((store.get(tensor), list(labels)))
        free.update(l for l in labels if isinstance(l, (int, numpy.integer)))
    return contractor.multi_contract(factors, sorted(free), alpha=expr.scalar)


def raw(tensor, engine=None):
    if isinstance(tensor, (_product, _sum)):
        return evaluate(tensor, engine).host()
    if isinstance(tensor, _device_tensor):
        return tensor.device_tensor.host()
    if isinstance(tensor, DeviceTensor):
        return tensor.host()
    return as_host(tensor)
