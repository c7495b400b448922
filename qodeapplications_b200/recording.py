"""A recorded launch sequence (lib.Context.begin_trace / Device.begin_trace) turned into ONE CUDA graph.

Both host paths issue hundreds of small launches per build whose arguments do not depend on the VALUES of the inputs.
`launch_graph` takes such a recorded sequence and captures it: either as a chain on one stream, or -- streams > 1 -- with
every call placed on one of several streams according to its real data dependencies (schedule.py), each stream launching
through an xr context of its own (no shared scratch), so that the captured graph has the true dependency structure and
independent kernels overlap.  Where graph capture is not available the recorded calls are re-issued one by one.
"""
import torch

from . import lib as _lib
from . import schedule


class launch_graph(object):
    def __init__(self, dev, trace, streams=32, graph=True, known=()):
        self.dev, self.trace = dev, trace
        self.graph, self.n_streams = None, 1
        self.graph_error = self.streams_error = None
        if graph and dev.torch_device.type == "cuda":
            if streams > 1:
                self._capture_streams(streams, known)
            if self.graph is None:
                self._capture_chain()

    def run(self):
        """launch the recorded sequence on the current stream"""
        if self.graph is not None:
            self.graph.replay()
        else:
            self.dev.ctx.replay(self.trace)

    def single_stream(self):
        """drop the multi-stream graph for the plain chain (taken when a multi-stream replay did not reproduce the eager result)"""
        self.graph, self.n_streams = None, 1
        if self.dev.torch_device.type == "cuda":
            self._capture_chain()

    def _capture_chain(self):
        dev = self.dev
        try:
            stream = torch.cuda.Stream(device=dev.torch_device)
            stream.wait_stream(torch.cuda.current_stream(dev.torch_device))
            g = torch.cuda.CUDAGraph()
            with dev.use_stream(stream):
                with torch.cuda.graph(g, stream=stream):
                    dev.ctx.replay(self.trace)
            torch.cuda.current_stream(dev.torch_device).wait_stream(stream)
            self.graph = g
        except Exception as exc:          # the recorded calls can still be re-issued one by one
            self.graph, self.graph_error = None, repr(exc)

    def _issue_on_streams(self, streams, contexts, stream_of, cross):
        """the recorded calls, each on its stream, with an event wait for every dependency that crosses streams; the first
        stream forks the others and joins them at the end"""
        needed = set(j for c in cross for j in c)
        start = torch.cuda.Event()
        start.record(streams[0])
        for s in streams[1:]:
            s.wait_event(start)
        events = {}
        for i, (call, args, kwargs) in enumerate(self.trace):
            s = streams[stream_of[i]]
            for j in cross[i]:
                s.wait_event(events[j])
            call(contexts[stream_of[i]], *args, **kwargs)
            if i in needed:
                events[i] = torch.cuda.Event()
                events[i].record(s)
        for s in streams[1:]:
            done = torch.cuda.Event()
            done.record(s)
            streams[0].wait_event(done)

    def _capture_streams(self, n_streams, known):
        dev = self.dev
        try:
            deps = schedule.dependencies(self.trace, known)
            stream_of, cross = schedule.assign_streams(deps, n_streams)
            streams = [torch.cuda.Stream(device=dev.torch_device) for _ in range(n_streams)]
            contexts = [_lib.Context(dev.index, s.cuda_stream) for s in streams]
            current = torch.cuda.current_stream(dev.torch_device)
            streams[0].wait_stream(current)
            self._issue_on_streams(streams, contexts, stream_of, cross)      # sizes every context's scratch before the capture
            current.wait_stream(streams[0])
            torch.cuda.synchronize(dev.index)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[0]):
                self._issue_on_streams(streams, contexts, stream_of, cross)
            self.graph, self.n_streams = g, n_streams
            self._stream_contexts = contexts                                   # their scratch buffers belong to the graph
            self.cross_stream_dependencies = sum(len(c) for c in cross)
        except Exception as exc:
            self.graph, self.streams_error = None, repr(exc)
