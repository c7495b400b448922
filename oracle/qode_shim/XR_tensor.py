"""Stand-in for XRbase/XR_tensor.py (which needs tensorly + opt_einsum, both absent):
same three entry points, XR_tensor.py:53-58, on the einsum-backed tensornet shim."""
import qode.math.tensornet as tensornet

def init(raw_tensor):
    return tensornet.primitive(raw_tensor)
def zeros():
    return tensornet.primitive(0.)
def raw(tensor):
    return tensornet.raw(tensor)
