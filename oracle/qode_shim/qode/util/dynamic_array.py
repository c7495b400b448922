"""qode.util.dynamic_array stand-in: lazy n-d mapping, arr[i,j] -> rule(i,j).
Call sites: hermitian-XRCC/precontract.py:80-94, diagrams/build_diagram.py:140-196."""


class dynamic_array(object):
    def __init__(self, rule, ranges):
        self._rule = rule
        self.ranges = ranges
    def __getitem__(self, indices):
        if not isinstance(indices, tuple):
            indices = (indices,)
        if len(indices) != len(self.ranges):
            raise IndexError("dynamic_array expects {} indices, got {}".format(len(self.ranges), len(indices)))
        for index, allowed in zip(indices, self.ranges):
            if allowed is not None and index not in allowed:
                raise KeyError(index)
        return self._rule(*indices)


def cached(rule):
    memory = {}
    def cached_rule(*indices):
        if indices not in memory:
            memory[indices] = rule(*indices)
        return memory[indices]
    return cached_rule


def wrap(array, wrappers):
    def rule(*indices):
        value = array[indices if len(indices) > 1 else indices[0]]
        for wrapper in wrappers:
            value = wrapper(value)
        return value
    return dynamic_array(rule, array.ranges)
