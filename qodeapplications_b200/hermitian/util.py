"""Small stand-ins for the qode.util helpers the reference's public signatures pass around
(struct: diagrammatic_expansion.py:113-114; timer: get_xr_result.py:46-48).  Any object with the same
methods (e.g. the real qode.util.timer) works in their place."""
import time


class struct(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)


class timer(object):
    def __init__(self):
        self._t0 = None
        self.totals, self.counts = {}, {}
    def start(self):
        self._t0 = time.perf_counter()
    def record(self, label):
        dt = time.perf_counter() - (self._t0 if self._t0 is not None else time.perf_counter())
        self.totals[label] = self.totals.get(label, 0.0) + dt
        self.counts[label] = self.counts.get(label, 0) + 1
    def print(self, title=""):
        print(title)
        for label in sorted(self.totals, key=self.totals.get, reverse=True):
            print("  {:32s} {:10.4f} s  ({} calls)".format(label, self.totals[label], self.counts[label]))
