"""Device-memory plumbing (PyTorch tensors as HBM buffers, torch's current stream as the launch
stream).  torch is used for allocation, host<->device copies and torch.distributed only; every
arithmetic kernel of the hot path is in libxr_b200.so."""
import numpy
import torch

from . import lib as _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.XRError("no CUDA device visible: qodeapplications_b200 has no CPU fallback")


class Device(object):
    """A GPU + an xr context bound to torch's current stream on it."""
    def __init__(self, index=None):
        require_cuda()
        if index is None:
            index = torch.cuda.current_device()
        self.index = int(index)
        self.torch_device = torch.device("cuda", self.index)
        with torch.cuda.device(self.index):
            stream = torch.cuda.current_stream(self.index).cuda_stream
        self.ctx = _lib.Context(self.index, stream)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.torch_device)

    def zeros(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype, device=self.torch_device)

    def upload(self, array, dtype=numpy.float64):
        """host ndarray -> device tensor on the current stream"""
        array = numpy.ascontiguousarray(array, dtype=dtype)
        host = torch.from_numpy(array)
        self.h2d_bytes += array.nbytes
        # pinned sources (bench.py pins its inputs) go asynchronously; pageable ones are copied by the driver's
        # own staging -- allocating a pinned bounce buffer per call costs more than it saves
        return host.to(self.torch_device, non_blocking=host.is_pinned())

    def download(self, tensor):
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.cpu().numpy()

    def sync(self):
        torch.cuda.synchronize(self.index)
