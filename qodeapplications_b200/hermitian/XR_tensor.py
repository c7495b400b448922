"""The tensor seam of the reference (XRbase/XR_tensor.py:53-58): ``init`` wraps primitive data,
``raw`` gives it back as an ndarray.  Here ``init`` simply tags a host ndarray; it is uploaded to the
GPU once, the first time a contraction needs it (tensor.DeviceStore)."""
import numpy

from .tensor import DeviceTensor, as_host


class xr_tensor(object):
    def __init__(self, array):
        self.array = numpy.ascontiguousarray(array, dtype=numpy.float64)
    @property
    def shape(self):
        return self.array.shape


def init(raw_tensor):
    return xr_tensor(raw_tensor)


def zeros():
    return xr_tensor(numpy.zeros(()))


def raw(tensor):
    if isinstance(tensor, DeviceTensor):
        return tensor.host()
    return as_host(tensor)
