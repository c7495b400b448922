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
        self._ring, self._ring_at = None, 0      # pinned staging ring for small pageable uploads

    def empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.torch_device)

    def zeros(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype, device=self.torch_device)

    STAGING_BYTES = 64 << 20

    def upload(self, array, dtype=numpy.float64):
        """host ndarray -> device tensor on the current stream.  Pinned sources (bench.py pins its inputs) are copied
        asynchronously as they are; small pageable ones (integral blocks, offset tables: hundreds per get_xr_H call) are
        staged through a pinned ring so the host never waits for the stream; large pageable ones go through the driver's
        own staging."""
        array = numpy.ascontiguousarray(array, dtype=dtype)
        host = torch.from_numpy(array)
        self.h2d_bytes += array.nbytes
        if host.is_pinned():
            return host.to(self.torch_device, non_blocking=True)
        if array.nbytes == 0 or array.nbytes > self.STAGING_BYTES // 4:
            return host.to(self.torch_device)
        if self._ring is None:
            self._ring = torch.empty(self.STAGING_BYTES, dtype=torch.uint8, pin_memory=True)
        size = (array.nbytes + 255) & ~255
        if self._ring_at + size > self.STAGING_BYTES:
            torch.cuda.current_stream(self.index).synchronize()      # every copy staged so far has left the ring
            self._ring_at = 0
        staged = self._ring[self._ring_at:self._ring_at + array.nbytes].view(host.dtype).reshape(host.shape)
        self._ring_at += size
        staged.copy_(host)
        return staged.to(self.torch_device, non_blocking=True)

    def download(self, tensor):
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.cpu().numpy()

    def sync(self):
        torch.cuda.synchronize(self.index)
