"""Device-memory plumbing (PyTorch tensors as HBM buffers, torch's current stream as the launch
stream).  torch is used for allocation, host<->device copies and torch.distributed only; every
arithmetic kernel of the hot path is in libxr_b200.so."""
import contextlib
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
        self._stream = stream          # the cudaStream_t every xr kernel of this device is launched on
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._ring, self._ring_at = None, 0      # pinned staging ring for small pageable uploads
        self._keep = None                        # while a launch trace is recorded: every buffer it touches stays alive

    def empty(self, shape, dtype=torch.float64):
        out = torch.empty(shape, dtype=dtype, device=self.torch_device)
        if self._keep is not None:
            self._keep.append(out)
        return out

    def zeros(self, shape, dtype=torch.float64):
        if self._keep is None:
            return torch.zeros(shape, dtype=dtype, device=self.torch_device)
        # recording a launch trace: the zero fill must be part of it (a replay accumulates into the same buffer again)
        out = self.empty(shape, dtype)
        if out.numel():
            self.ctx.memset_zero(out, out.numel() * out.element_size())
        return out

    @property
    def tracing(self):
        return self._keep is not None

    def begin_trace(self):
        """start recording every xr launch (and zero fill) issued through this device; buffers created meanwhile are kept"""
        self._keep = []
        self.ctx.begin_trace()

    def end_trace(self):
        """-> (launch trace, buffers the trace refers to)"""
        keep, self._keep = self._keep, None
        return self.ctx.end_trace(), keep

    def _check_stream(self):
        """torch allocates, copies and synchronises on ITS current stream; the xr kernels run on the stream the context
        was bound to.  If the two differ, uploads race with the kernels that read them -- refuse instead."""
        current = torch.cuda.current_stream(self.index).cuda_stream
        if current != self._stream:
            raise _lib.XRError("torch's current stream (0x%x) is not the stream this Device launches on (0x%x): "
                               "wrap the work in `with device.use_stream(stream):`" % (current, self._stream))

    @contextlib.contextmanager
    def use_stream(self, stream):
        """run xr kernels AND torch copies/allocations on `stream` (a torch.cuda.Stream) inside the block"""
        previous = self._stream
        with torch.cuda.stream(stream):
            self.ctx.set_stream(stream.cuda_stream)
            self._stream = stream.cuda_stream
            try:
                yield self
            finally:
                self.ctx.set_stream(previous)
                self._stream = previous

    STAGING_BYTES = 64 << 20

    def upload(self, array, dtype=numpy.float64):
        """host ndarray -> device tensor on the current stream.  Pinned sources (bench.py pins its inputs) are copied
        asynchronously as they are; small pageable ones (integral blocks, offset tables: hundreds per get_xr_H call) are
        staged through a pinned ring so the host never waits for the stream; large pageable ones go through the driver's
        own staging."""
        self._check_stream()
        array = numpy.ascontiguousarray(array, dtype=dtype)
        host = torch.from_numpy(array)
        self.h2d_bytes += array.nbytes
        if self._keep is not None:
            out = host.to(self.torch_device)
            self._keep.append(out)
            return out
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
        self._check_stream()
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.cpu().numpy()

    def sync(self):
        torch.cuda.synchronize(self.index)
