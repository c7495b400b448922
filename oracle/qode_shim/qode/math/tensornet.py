"""einsum-backed stand-in for qode.math.tensornet, covering what the hot path uses:
primitive tensors called with index labels (ints = free output axes, sorted ascending in
the result; strings = contracted, also between the two factors of a product, which is how
diagram_hack.state_indices traces ket/bra pairs), `@` products, scalar `*`, evaluate(),
raw().  Call sites: precontract.py:75, every diagrams/*.py, XRbase/XR_tensor.py:53-58."""
import numpy

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


class ContractionError(Exception):
    pass


class _product(object):
    def __init__(self, factors, scalar=1.):
        self.factors = factors      # list of (ndarray, labels)
        self.scalar = scalar
    def __matmul__(self, other):
        other = _as_product(other)
        return _product(self.factors + other.factors, self.scalar * other.scalar)
    def __mul__(self, scalar):
        return _product(self.factors, self.scalar * scalar)
    __rmul__ = __mul__
    def __neg__(self):
        return _product(self.factors, -self.scalar)


class primitive(object):
    def __init__(self, data):
        self.data = numpy.asarray(data, dtype=numpy.float64)
    @property
    def shape(self):
        return self.data.shape
    def __call__(self, *labels):
        if len(labels) != self.data.ndim:
            raise ContractionError("tensor of rank {} called with {} indices".format(self.data.ndim, len(labels)))
        return _product([(self.data, tuple(labels))])
    def __matmul__(self, other):      # rank-0 convenience
        return self() @ other
    def __mul__(self, scalar):
        return primitive(self.data * scalar)
    __rmul__ = __mul__


def _as_product(obj):
    if isinstance(obj, _product):
        return obj
    if isinstance(obj, primitive):
        return obj()
    raise ContractionError("cannot contract {}".format(type(obj)))


def _evaluate_ndarray(expr):
    expr = _as_product(expr)
    letters = {}
    def letter(label):
        if label not in letters:
            letters[label] = _LETTERS[len(letters)]
        return letters[label]
    terms = []
    operands = []
    free = set()
    for data, labels in expr.factors:
        terms.append("".join(letter(label) for label in labels))
        operands.append(data)
        free.update(label for label in labels if isinstance(label, (int, numpy.integer)))
    output = "".join(letters[label] for label in sorted(free))
    result = numpy.einsum(",".join(terms) + "->" + output, *operands, optimize=True)
    return expr.scalar * result


def evaluate(expr):
    return primitive(_evaluate_ndarray(expr))


def raw(expr):
    if isinstance(expr, primitive):
        return expr.data
    return numpy.asarray(_evaluate_ndarray(expr))


def scalar_value(expr):
    return float(raw(expr))


def shape(tensor):
    return tensor.shape
