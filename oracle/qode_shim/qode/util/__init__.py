"""qode.util stand-ins inferred from call sites:
struct      -- hermitian-XRCC/diagrammatic_expansion.py:113-114, get_xr_result.py:64
timer       -- diagrams/build_diagram.py:48-50, precontract.py:74-76, XR_term.py:56-58
recursive_looper, compound_range -- hermitian-XRCC/XR_term.py:41,69-70,94
"""
import itertools
import time
import numpy


class struct(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)
    def update(self, **kwargs):
        self.__dict__.update(kwargs)
        return self
    def __call__(self, *names):
        return struct(**{name: self.__dict__[name] for name in names})
    def __contains__(self, name):
        return name in self.__dict__


class timer(object):
    def __init__(self):
        self._t0 = None
        self.totals = {}
        self.counts = {}
    def start(self):
        self._t0 = time.perf_counter()
    def record(self, label):
        dt = time.perf_counter() - self._t0
        self.totals[label] = self.totals.get(label, 0.) + dt
        self.counts[label] = self.counts.get(label, 0) + 1
    def print(self, title=""):
        print(title)
        for label in sorted(self.totals, key=self.totals.get, reverse=True):
            print("  {:30s} {:10.4f} s  ({} calls)".format(label, self.totals[label], self.counts[label]))


def recursive_looper(loops, kernel):
    """loops = [(level, iterable), ...]; calls kernel(i_0, i_1, ...) over the full nest."""
    ranges = [list(r) for _, r in loops]
    for indices in itertools.product(*ranges):
        kernel(*indices)


def compound_range(ranges, inactive=()):
    """Yield fresh index *lists* over the product of the active ranges; positions listed
    in `inactive` are left as None placeholders for the caller to fill."""
    inactive = set(inactive)
    active = [list(r) if m not in inactive else [None] for m, r in enumerate(ranges)]
    for indices in itertools.product(*active):
        yield list(indices)


def sort_eigen(eig):
    vals, vecs = eig
    order = numpy.argsort(vals.real)
    return vals[order], vecs[:, order]
