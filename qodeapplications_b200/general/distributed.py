"""One-process-per-GPU sharding of the general-XRCC H build.

* dimers: rank r builds the rows of H2[m1][m2] whose fragment-m1 bra state lies in its slab (equal
  slabs of ceil(dim1/world) states, so the assemble step is ONE all_gather_into_tensor per dimer,
  in place: every rank writes its slab directly into its slice of the full matrix);
* trimers: rank r streams its slab of every class's leading pair index; the only exchange is the sum
  of 24 doubles per trimer (moments).  No data-path collective exists anywhere else.

Assemble step.  The H2 all-gather moves 8 B per element against 2K flop per element of compute: at n = 18 it is
~20x the dimer compute of a rank, and it sits in front of a trimer phase that keeps every SM busy with a persistent
FP64 kernel (all 64 K registers of each SM).  Two ways to run it, both off the launch stream:

* "nccl": one all_gather_into_tensor per dimer, issued asynchronously so that it overlaps the next dimer's kernels, joined
  before the trimer phase (NCCL's kernels need SMs: side by side with the trimer stream they make the step longer);
* "ce" (what "auto", the default, takes when it can): the H2 buffers live in symmetric memory (torch.distributed._symmetric_memory: every rank maps every peer's buffer
  over NVLink), and after a device-side barrier each rank PULLS the other ranks' slabs with plain device-to-device copies
  on a side stream -- the copy engines move the data through the NVSwitch while the SMs stream trimer tiles, so the
  gather costs the step nothing (N = 2: 5798 ms with it, 5796 ms without; 732 GB/s per rank when timed alone = 95 % of
  the measured 770 GB/s peer copy).  Falls back to "nccl" when symmetric memory cannot be set up on the box.

torch.distributed is plumbing here (NCCL on GPUs, gloo in the CPU tests); the arithmetic is in
libxr_b200.so.
"""
import numpy
import torch
import torch.distributed as dist


def slab_bounds(dim, rank, world):
    """(lo, hi, per): bra states [lo, hi) of this rank, per = padded slab size shared by all ranks"""
    per = -(-dim // world)
    return min(rank * per, dim), min((rank + 1) * per, dim), per


def balanced_shard(n_states, rank, world):
    """{charge: (i_lo, i_hi)}: an equal share of the states of every charge sector -- the bra range of a rank in the
    streamed dimer build, where every rank then has the same pair count in every charge-transfer class"""
    out = {}
    for chg, n in n_states.items():
        base, extra = divmod(n, world)
        lo = rank * base + min(rank, extra)
        out[chg] = (lo, lo + base + (1 if rank < extra else 0))
    return out


class _peer_gather(object):
    """Copy-engine all-gather of row slabs over NVLink peer memory (see the module docstring)."""
    def __init__(self, dev, shapes, rank, world, group):
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world = rank, world
        self.local, self.handle = {}, {}
        group = group if group is not None else dist.group.WORLD
        for key, shape in shapes.items():
            t = symm.empty(shape, dtype=torch.float64, device=dev.torch_device)
            self.local[key] = t
            self.handle[key] = symm.rendezvous(t, group=group)
        self.stream = torch.cuda.Stream(device=dev.torch_device)
        self.done = None

    def gather(self, key, rows_per_rank):
        """called on the launch stream right after this rank's slab of `key` was queued"""
        ready = torch.cuda.Event()
        ready.record()
        full, handle = self.local[key], self.handle[key]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            handle.barrier(channel=0)                       # every rank's slab is complete
            for step in range(1, self.world):               # staggered: at any time each rank is read by one peer
                p = (self.rank + step) % self.world
                theirs = handle.get_buffer(p, tuple(full.shape), torch.float64)
                rows = slice(p * rows_per_rank, (p + 1) * rows_per_rank)
                full[rows].copy_(theirs[rows], non_blocking=True)
            handle.barrier(channel=1)                       # nobody rewrites its slab while a peer still reads it
            self.done = torch.cuda.Event()
            self.done.record()

    def join(self):
        if self.done is not None:
            torch.cuda.current_stream().wait_event(self.done)
            self.done = None


class sharded_build(object):
    """Holds the output buffers of a (possibly multi-rank) build and runs one build step."""
    def __init__(self, engine, dimers, trimers, rank=0, world=1, group=None, assemble="auto"):
        self.eng, self.dimers, self.trimers = engine, list(dimers), list(trimers)
        self.rank, self.world, self.group = rank, world, group
        self.dims = [engine._frag(m).dim for m in range(len(engine._supersystem))]
        self.H1, self.H2, self.H3_moments = {}, {}, {}
        dev = engine.dev
        shapes = {}
        for m1, m2 in self.dimers:
            lo, hi, per = slab_bounds(self.dims[m1], rank, world)
            # rows padded to world*per bra states so that all slabs have equal size; rows >= dim1*dim2 are unused
            shapes[(m1, m2)] = (per * world * self.dims[m2], self.dims[m1] * self.dims[m2])
        self.peer, self.assemble, self.assemble_note = None, "nccl", None
        if assemble in ("ce", "auto") and world > 1 and dev.torch_device.type == "cuda":
            ok = torch.ones(1, device=dev.torch_device)
            try:
                self.peer = _peer_gather(dev, shapes, rank, world, group)
            except Exception as exc:         # symmetric memory is not available on every box: NCCL then
                self.assemble_note = "symmetric memory unavailable (%s): NCCL all-gather" % (repr(exc)[:200],)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if float(ok.item()) == 0.0:
                self.peer = None
            else:
                self.assemble = "ce"
        for key, shape in shapes.items():
            self.H2[key] = self.peer.local[key] if self.peer is not None else dev.empty(shape)

    def my_rows(self, m1, m2):
        """view of this rank's slab of H2[m1][m2] (only the rows that exist)"""
        lo, hi, per = slab_bounds(self.dims[m1], self.rank, self.world)
        d2 = self.dims[m2]
        return self.H2[(m1, m2)][self.rank * per * d2:(self.rank * per + (hi - lo)) * d2]

    def full(self, m1, m2):
        """the assembled H2[m1][m2] (valid on every rank after step(gather=True))"""
        return self.H2[(m1, m2)][:self.dims[m1] * self.dims[m2]]

    def step(self, gather=True, after_dimers=None, trimers=None, overlap=True):
        """One build pass: every H1, every dimer block, and the trimers named by `trimers` (default: all of them; bench.py
        passes one per step, round-robin, so that len(self.trimers) consecutive steps make one full build).

        With world > 1 each dimer's all-gather is issued asynchronously (overlap=True): NCCL runs it on its own stream
        once the slab's kernels have finished, while this rank already builds the next dimer and then streams its
        trimer slabs; the launch stream only joins the gathers at the end of the step.  after_dimers(), if given, is called
        once every H1/H2 launch of the step has been queued and before the trimer streams are: the point where a caller
        can start reading ITS rows of the dimer blocks back on a second stream while the (much longer) trimer phase runs
        (the rows of the other ranks are only guaranteed after step() returns)."""
        eng, rank, world = self.eng, self.rank, self.world
        pending = []
        for m in range(len(self.dims)):
            self.H1[m] = eng.H1_device(m)
            if world > 1 and m in eng._held:       # each rank built the rows of the bra slab it holds
                dist.all_reduce(self.H1[m], group=self.group)
        for m1, m2 in self.dimers:
            lo, hi, per = slab_bounds(self.dims[m1], rank, world)
            eng.H2_device(m1, m2, bra_range=(lo, hi) if world > 1 else None, out=self.my_rows(m1, m2))
            if world > 1 and gather:
                d2 = self.dims[m2]
                full = self.H2[(m1, m2)]
                if self.peer is not None:
                    self.peer.gather((m1, m2), per * d2)      # gathers queue up in order on the peer stream
                    if not overlap:
                        self.peer.join()
                    elif self.peer not in pending:
                        pending.append(self.peer)
                    continue
                work = dist.all_gather_into_tensor(full, full[rank * per * d2:(rank + 1) * per * d2], group=self.group,
                                                   async_op=overlap)
                if overlap:
                    pending.append(work)
        if after_dimers is not None:
            after_dimers()
        # NCCL's all-gather kernels need SMs, and the trimer stream holds every SM with a persistent, statically partitioned
        # kernel: run side by side, the trimer kernels wait for the SMs NCCL took and the step gets LONGER (measured at
        # N = 2: 6364 ms overlapped vs 5864 ms one after the other, profiles/r02e).  So NCCL gathers only overlap the next
        # dimers and are joined here; the copy-engine gathers use no SM and run under the whole trimer phase.
        for work in [w for w in pending if w is not self.peer]:
            work.wait()
        for ms in (self.trimers if trimers is None else trimers):
            self.H3_moments[ms] = eng.H3_moments_device(*ms, shard=(rank, world))
        if self.peer in pending:
            self.peer.join()       # the launch stream waits for the last pull and the closing barrier (no host block)

    def recorded(self, streams=16, graph=True, **step_arguments):
        """This build's step() recorded once and replayed as ONE CUDA graph (single rank): for the small systems the
        reference loops over (Be2 / Be3 shapes: a step is a few hundred launches whose Python planning costs ten times the
        GPU work).  Returns a recorded_step; its outputs are this object's H1 / H2 / H3_moments buffers."""
        return recorded_step(self, streams, graph, step_arguments)

    def gather_bytes(self):
        """bytes this rank RECEIVES over NVLink per step for the dimer all-gathers (the other ranks' slabs)"""
        total = 0
        for m1, m2 in self.dimers:
            lo, hi, per = slab_bounds(self.dims[m1], self.rank, self.world)
            total += (self.world - 1) * per * self.dims[m2] * self.dims[m1] * self.dims[m2] * 8
        return total if self.world > 1 else 0

    def reduced_moments(self):
        """{trimer: [sum, sum of squares]} summed over classes and ranks"""
        out = {}
        for ms in self.trimers:
            t = self.H3_moments[ms].clone()
            if self.world > 1:
                dist.all_reduce(t, group=self.group)
            out[ms] = [float(x) for x in self.eng.dev.download(t).sum(axis=0)]
        return out


class recorded_step(object):
    """sharded_build.step() as a recorded launch sequence (recording.launch_graph).  Inputs are the engine's device copies
    of the densities (build_matrix_elements._rho_dev): update(fragments) copies new values into them, run() launches the
    graph, and the results are in build.H1 / build.full(m1, m2) / build.H3_moments afterwards."""
    def __init__(self, build, streams=16, graph=True, step_arguments=None):
        from ..recording import launch_graph
        if build.world != 1:
            raise NotImplementedError("recorded_step: one rank (the assemble collectives are not recorded)")
        self.build, self.arguments = build, dict(step_arguments or {})
        eng, dev = build.eng, build.eng.dev
        eng.preload()
        build.step(**self.arguments)                          # eager: sizes the context scratch, fills the table caches
        dev.begin_trace()
        try:
            build.step(**self.arguments)
        finally:
            self.trace, self._alive = dev.end_trace()
        self.launches = len(self.trace)
        known = [t for t in self._alive if t.dtype == torch.int64] + [t for t in eng._idx_dev.values()]
        self._launcher = launch_graph(dev, self.trace, streams=streams, graph=graph, known=known)
        self.graph, self.n_streams = self._launcher.graph, self._launcher.n_streams

    def update(self, fragments):
        """new density values (same shapes) for the recorded build: fragments[m].rho[op][(ci, cj)] as for build_matrix_elements"""
        eng = self.build.eng
        for (m, op, sector), slot in eng._rho_dev.items():
            block = fragments[m].rho[op][sector]
            if isinstance(block, torch.Tensor):
                if block.data_ptr() != slot.data_ptr():
                    slot.copy_(block.reshape(slot.shape))
                continue
            host = torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(block, dtype=numpy.float64)).reshape(tuple(slot.shape)))
            slot.copy_(host, non_blocking=host.is_pinned())

    def run(self):
        self._launcher.run()
        return self.build
