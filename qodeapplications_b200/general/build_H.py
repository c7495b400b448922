"""B200 drop-in for the reference's general-XRCC/build_H.py.

``build_matrix_elements(supersystem, integrals, nuc_repulsion)`` keeps the reference's constructor
and its ``monomer / dimer / trimer`` element accessors (general-XRCC/build_H.py:37-188), so
general-XRCC/test_H.py:49-58 runs unchanged -- but an element is never computed on its own.  The
first request for an element of a (fragment tuple) block builds the WHOLE block on the GPU and the
accessor then reads the cached result.  New block-level methods (``H1``, ``H2``, ``H3``,
``H3_moments``) return whole matrices in test_H.py's layout (test_H.py:65-107: state order =
``fragment.state_indices``, product basis row-major) and are what a batched caller should use.

How a block is built: flatten each fragment's (bra,ket) state pair
into one index P; for every charge-transfer class the block is ``A[P1,K] . B[P2,K]^T`` with

  class (d1)   K           A (fragment 1)                               B (fragment 2)
  0            n2^2 + 2    4 ca1.V | delta1 | U.ca1 + nuc*delta1         ca2 | U.ca2 | delta2
  -2 / +2      n^2         cc.V    (or aa raw)                           aa raw (or cc.V)
  -1 / +1      n1 + n2     (h + 2 cca.V).c | c     (or a | 2 V.caa)      sign2 * (a | 2 V.caa)  (or ...)

so each class is one ``xr_gemm_scatter`` whose offset tables write straight into H2's final
layout; the precontractions ``rho . V`` are the same kernel with K = n^2 or n^3.  Every index
permutation is applied to the tiny integral blocks on the host, never to a density.  Trimer
classes are ``sum_rs W[Pk,r,s] beta[Pb,r] gamma[Pc,s]`` streamed by ``xr_trimer_stream``.

All arithmetic runs in libxr_b200.so (hand-written sm_100a CUDA); there is no CPU path.
"""
import itertools
import numpy
import torch

from .. import lib as _lib
from ..device import Device


def _parity(n_elec_ref, chg):
    """build_H.py:68,113-114"""
    return -1.0 if (n_elec_ref - chg) % 2 else 1.0


class _FragInfo(object):
    def __init__(self, fragment, n_orb):
        self.n_orb = n_orb
        self.n_elec_ref = fragment.n_elec_ref
        rho = fragment.rho
        self.n_states = {}
        if hasattr(fragment, "n_states"):       # explicit (needed when only a slab of the densities is held)
            self.n_states = dict(fragment.n_states)
        else:
            for (ci, cj), block in rho["ca"].items():
                if ci == cj:
                    self.n_states[ci] = len(block)
        if hasattr(fragment, "state_indices"):
            self.state_indices = list(fragment.state_indices)
        else:   # general-XRCC/Be631g.py:81-85 with ref_state=(0,0)
            ref = 0 if 0 in self.n_states else sorted(self.n_states)[0]
            self.state_indices = [(ref, i) for i in range(self.n_states[ref])]
            for chg in self.n_states:
                if chg != ref:
                    self.state_indices += [(chg, i) for i in range(self.n_states[chg])]
        self.dim = len(self.state_indices)
        where = {state: p for p, state in enumerate(self.state_indices)}
        self.charges = list(self.n_states)
        self.pos = {chg: numpy.array([where.get((chg, i), -1) for i in range(n)], dtype=numpy.int64)
                    for chg, n in self.n_states.items()}
        for chg, p in self.pos.items():
            if (p < 0).any():
                raise ValueError("state_indices does not list every state of charge %r" % (chg,))
        self.where = where

    def sectors(self, delta):
        """[(bra charge, ket charge)] with bra - ket = delta, both present"""
        return [(ci, ci - delta) for ci in self.charges if (ci - delta) in self.n_states]


def _bra_states(info, chg, bra_range):
    """[i_lo, i_hi): the states of charge chg selected by a bra range, which is None (all), a pair (lo, hi) of MATRIX
    positions (a row slab of the block) or a dict {charge: (i_lo, i_hi)} of per-sector state ranges (a balanced shard:
    every rank gets the same share of every charge sector)."""
    if bra_range is None:
        return 0, info.n_states[chg]
    if isinstance(bra_range, dict):
        i_lo, i_hi = bra_range.get(chg, (0, 0))
        if not 0 <= i_lo <= i_hi <= info.n_states[chg]:
            raise ValueError("state range %r outside charge sector %r" % ((i_lo, i_hi), chg))
        return int(i_lo), int(i_hi)
    inside = numpy.nonzero((info.pos[chg] >= bra_range[0]) & (info.pos[chg] < bra_range[1]))[0]
    if len(inside) == 0:
        return 0, 0
    if int(inside[-1]) + 1 - int(inside[0]) != len(inside):
        raise NotImplementedError("bra slab is not contiguous inside charge sector %r" % (chg,))
    return int(inside[0]), int(inside[-1]) + 1


class _PairClass(object):
    """All (bra,ket) state pairs of one fragment with bra charge - ket charge = delta, optionally
    restricted to the bra states a bra range selects (_bra_states)."""
    def __init__(self, info, delta, bra_range=None):
        self.delta = delta
        self.sectors = []          # (ci, cj, i_lo, i_hi, row_offset)
        rows = 0
        for ci, cj in info.sectors(delta):
            Nj = info.n_states[cj]
            i_lo, i_hi = _bra_states(info, ci, bra_range)
            if i_hi > i_lo and Nj > 0:
                self.sectors.append((ci, cj, i_lo, i_hi, rows))
                rows += (i_hi - i_lo) * Nj
        self.P = rows

    def offsets(self, info, bra_stride, ket_stride, bra_base=0):
        """int64[P]: (pos[ci][i]-bra_base)*bra_stride + pos[cj][j]*ket_stride for every row"""
        out = numpy.empty(self.P, dtype=numpy.int64)
        for ci, cj, i_lo, i_hi, off in self.sectors:
            Nj = info.n_states[cj]
            block = ((info.pos[ci][i_lo:i_hi] - bra_base) * bra_stride)[:, None] + (info.pos[cj] * ket_stride)[None, :]
            out[off:off + (i_hi - i_lo) * Nj] = block.reshape(-1)
        return out

    def diagonal_rows(self, info):
        """rows with identical bra and ket state (only meaningful for delta == 0)"""
        rows = []
        for ci, cj, i_lo, i_hi, off in self.sectors:
            if ci == cj:
                Nj = info.n_states[cj]
                i = numpy.arange(i_lo, i_hi, dtype=numpy.int64)
                rows.append(off + (i - i_lo) * Nj + i)
        return numpy.concatenate(rows) if rows else numpy.zeros(0, dtype=numpy.int64)


def _even(k):
    return k + (k & 1)


class build_matrix_elements(object):
    def __init__(self, supersystem, integrals, nuc_repulsion, device=None, held=None):
        """held (optional, for sharded inputs): {fragment index: bra range} -- the densities of that fragment are only
        supplied for the bra states the range selects, either (lo, hi) matrix positions or {charge: (i_lo, i_hi)} state
        ranges per charge sector: rho[op][(ci,cj)] then has one leading row per HELD bra state of charge ci (and the
        fragment object must carry ``n_states``).  Density blocks may also be CUDA torch tensors already on the device."""
        n_elec = [fragment.n_elec_ref for fragment in supersystem]
        rho = [fragment.rho for fragment in supersystem]
        self.data = rho, integrals.T, integrals.U, integrals.V, nuc_repulsion, n_elec     # as build_H.py:41
        self._supersystem = supersystem
        self._device_arg = device
        self._held = dict(held or {})
        self._dev = None
        self._info = None
        self._rho_dev = {}
        self._int_dev = {}
        self._idx_dev = {}
        self.profile = None        # set to a list to collect (label, algorithmic flops, algorithmic HBM bytes, start event, end event)
        self._H1 = {}
        self._H2 = {}
        self._H3 = {}

    # ------------------------------------------------------------------ reference accessors
    def monomer(self, fragment, I, J):
        """build_H.py:42-55"""
        H = self._H1.get(fragment)
        if H is None:
            H = self._H1[fragment] = self.H1(fragment)
        info = self._frag(fragment)
        return float(H[info.where[tuple(I)], info.where[tuple(J)]])

    def dimer(self, fragments, I, J):
        """build_H.py:56-102"""
        fragments = tuple(fragments)
        H = self._H2.get(fragments)
        if H is None:
            H = self._H2[fragments] = self.H2(*fragments)
        f1, f2 = (self._frag(m) for m in fragments)
        i = f1.where[tuple(I[0])] * f2.dim + f2.where[tuple(I[1])]
        j = f1.where[tuple(J[0])] * f2.dim + f2.where[tuple(J[1])]
        return float(H[i, j])

    def trimer(self, fragments, I, J):
        """build_H.py:103-188"""
        fragments = tuple(fragments)
        H = self._H3.get(fragments)
        if H is None:
            H = self._H3[fragments] = self.H3(*fragments)
        f = [self._frag(m) for m in fragments]
        i = (f[0].where[tuple(I[0])] * f[1].dim + f[1].where[tuple(I[1])]) * f[2].dim + f[2].where[tuple(I[2])]
        j = (f[0].where[tuple(J[0])] * f[1].dim + f[1].where[tuple(J[1])]) * f[2].dim + f[2].where[tuple(J[2])]
        return float(H[i, j])

    # ----------------------------------------------------------------------------- plumbing
    @property
    def dev(self):
        if self._dev is None:
            self._dev = self._device_arg if hasattr(self._device_arg, "ctx") else Device(self._device_arg)
        return self._dev

    def _frag(self, m):
        if self._info is None:
            T = self.data[1]
            self._info = [_FragInfo(f, T[k, k].shape[0]) for k, f in enumerate(self._supersystem)]
        return self._info[m]

    def _held_range(self, m, chg):
        """[h_lo, h_hi): the bra states of charge chg of fragment m whose densities are held here"""
        return _bra_states(self._frag(m), chg, self._held.get(m))

    def _rho(self, m, op, sector):
        """device tensor [N_bra(held)*N_ket, n^k] of rho[m][op][sector]"""
        key = (m, op, sector)
        if key not in self._rho_dev:
            info = self._frag(m)
            block = self.data[0][m][op][sector]
            h_lo, h_hi = self._held_range(m, sector[0])
            rows = (h_hi - h_lo) * info.n_states[sector[1]]
            if isinstance(block, torch.Tensor):
                if block.dtype != torch.float64 or not block.is_contiguous():
                    raise TypeError("device-resident densities must be contiguous float64 tensors")
                self._rho_dev[key] = block.reshape(rows, info.n_orb ** len(op))
            else:
                arr = numpy.asarray(block, dtype=numpy.float64)
                self._rho_dev[key] = self.dev.upload(arr.reshape(rows, info.n_orb ** len(op)))
        return self._rho_dev[key]

    def _rho_rows(self, m, op, ci, cj, i_lo, i_hi):
        """(device tensor, element offset of bra state i_lo) for rows [i_lo, i_hi) of rho[m][op][(ci,cj)]"""
        src = self._rho(m, op, (ci, cj))
        h_lo, h_hi = self._held_range(m, ci)
        if i_lo < h_lo or i_hi > h_hi:
            raise ValueError("fragment %d: bra states [%d,%d) of charge %r requested but only [%d,%d) are held"
                             % (m, i_lo, i_hi, ci, h_lo, h_hi))
        return src, (i_lo - h_lo) * self._frag(m).n_states[cj] * src.shape[1]

    def _ints(self, key, make):
        """device copy of a (permuted) integral block, cached by key"""
        if key not in self._int_dev:
            self._int_dev[key] = self.dev.upload(make())
        return self._int_dev[key]

    def _index(self, array, key=None):
        """device int64 table; cached when the caller names it with a key"""
        if key is None:
            return self.dev.upload(array() if callable(array) else array, dtype=numpy.int64)
        if key not in self._idx_dev:
            self._idx_dev[key] = self.dev.upload(array() if callable(array) else array, dtype=numpy.int64)
        return self._idx_dev[key]

    class _timed(object):
        """records a pair of CUDA events on the launch stream around a kernel when profiling is on"""
        def __init__(self, owner, label, flops, nbytes=0.0):
            self.owner, self.label, self.flops, self.nbytes = owner, label, flops, nbytes
        def __enter__(self):
            if self.owner.profile is not None:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self
        def __exit__(self, *exc):
            if self.owner.profile is not None:
                self.e1.record()
                self.owner.profile.append((self.label, self.flops, self.nbytes, self.e0, self.e1))
            return False

    def drop_caches(self, densities=False):
        self._H1.clear(); self._H2.clear(); self._H3.clear()
        if densities:
            self._rho_dev.clear(); self._int_dev.clear()

    def preload(self, fragments=None, ops=("a", "c", "aa", "cc", "ca", "caa", "cca")):
        """Upload every density of the given fragments now (so later builds start HBM-resident)."""
        rho = self.data[0]
        for m in (range(len(rho)) if fragments is None else fragments):
            for op in ops:
                for sector in rho[m].get(op, {}):
                    self._rho(m, op, sector)

    def preload_distributed(self, rank, world, group=None, fragments=None, ops=("a", "c", "aa", "cc", "ca", "caa", "cca")):
        """Multi-GPU variant of preload(): every density block is uploaded over PCIe by ONE rank (blocks dealt out so that
        the ranks carry equal bytes) and reaches the others by a broadcast over NVLink -- N ranks that each need all the
        densities then move them across the host links once instead of N times (cfg4 at N = 8: 4 GB instead of 25 GB).
        Only for engines that hold full densities (no held= slabs)."""
        import torch.distributed as dist
        if world == 1:
            return self.preload(fragments, ops)
        if self._held:
            raise ValueError("preload_distributed: this engine holds bra slabs (held=); every rank uploads its own")
        rho = self.data[0]
        keys = []
        for m in (range(len(rho)) if fragments is None else fragments):
            info = self._frag(m)
            for op in ops:
                for sector in sorted(rho[m].get(op, {})):
                    block = rho[m][op][sector]
                    if isinstance(block, torch.Tensor):
                        continue                                    # already on the device
                    rows = info.n_states[sector[0]] * info.n_states[sector[1]]
                    keys.append((rows * info.n_orb ** len(op), m, op, sector, rows))
        load, owner = [0] * world, {}
        for size, m, op, sector, rows in sorted(keys, key=lambda k: (-k[0], k[1], k[2], k[3])):      # same deal on every rank
            r = min(range(world), key=lambda x: (load[x], x))
            owner[(m, op, sector)] = r
            load[r] += size
        pending = []
        for size, m, op, sector, rows in keys:
            key = (m, op, sector)
            if key in self._rho_dev:
                continue
            if owner[key] == rank:
                t = self._rho(m, op, sector)
            else:
                t = self._rho_dev[key] = self.dev.empty((rows, self._frag(m).n_orb ** len(op)))
            pending.append(dist.broadcast(t, src=owner[key], group=group, async_op=True))
        for work in pending:
            work.wait()

    # ----- factor builders: fill columns [col0, col0+F) of a [P, ld] class buffer -------------
    def _fill_raw(self, out, ld, col0, m, op, cls, sign=None):
        ctx = self.dev.ctx
        info = self._frag(m)
        for ci, cj, i_lo, i_hi, off in cls.sectors:
            Nj = info.n_states[cj]
            src, first = self._rho_rows(m, op, ci, cj, i_lo, i_hi)
            F = src.shape[1]
            alpha = 1.0 if sign is None else sign(ci)
            ctx.copy2d_scaled(out.data_ptr() + 8 * (off * ld + col0), ld, src.data_ptr() + 8 * first, F,
                              (i_hi - i_lo) * Nj, F, alpha)

    def _fill_contracted(self, out, ld, col0, m, op, cls, Wt, scale=1.0, sign=None, accumulate=False):
        """out[:, col0:col0+F] (+)= scale*sign * rho[P,K] . Wt[F,K]^T"""
        ctx = self.dev.ctx
        info = self._frag(m)
        F, K = Wt.shape
        for ci, cj, i_lo, i_hi, off in cls.sectors:
            Nj = info.n_states[cj]
            src, first = self._rho_rows(m, op, ci, cj, i_lo, i_hi)
            assert src.shape[1] == K, (op, src.shape, K)
            alpha = scale * (1.0 if sign is None else sign(ci))
            ctx.gemm_scatter((i_hi - i_lo) * Nj, F, K, alpha, src.data_ptr() + 8 * first, K, Wt, K,
                             out.data_ptr() + 8 * (off * ld + col0), None, ld, None, accumulate)

    # -------------------------------------------------------------------------------- blocks
    def H1(self, m):
        """H1[m] (dense, host ndarray) -- build_H.py:42-55 for every state pair at once."""
        return self.dev.download(self.H1_device(m))

    def H1_device(self, m):
        """H1[m] on the device.  When only a bra slab of fragment m's densities is held (held=), the rows of the held
        bra states are built and the others stay zero: the slabs of all ranks add up to the full block
        (distributed.sharded_build sums them)."""
        rho, T, U, V, nuc, n_elec = self.data
        info = self._frag(m)
        ctx, n = self.dev.ctx, info.n_orb
        H = self.dev.zeros((info.dim, info.dim))
        cls = _PairClass(info, 0, self._held.get(m))
        if cls.P == 0:
            return H
        h = self._ints(("h1", m), lambda: (T[m, m] + U[m, m, m]).reshape(1, n * n))
        off = self._index(cls.offsets(info, info.dim, 1))
        one = self._ints(("one",), lambda: numpy.ones((1, 2)))
        held_pos = []
        for ci, cj, i_lo, i_hi, row0 in cls.sectors:
            N = info.n_states[cj]
            rows = (i_hi - i_lo) * N
            ca, first = self._rho_rows(m, "ca", ci, cj, i_lo, i_hi)
            ctx.gemm_scatter(rows, 1, n * n, 1.0, ca.data_ptr() + 8 * first, n * n, h, n * n, H, off.data_ptr() + 8 * row0, 0, None, False)
            h_lo, h_hi = self._held_range(m, ci)
            scal = self._rho_dev.get((m, "ccaa", (ci, cj)))
            if scal is None:
                scal = rho[m]["ccaa"][(ci, cj)]
                if isinstance(scal, torch.Tensor):      # device-resident (general.build_density_tensors(device_result=True))
                    scal = scal.reshape((h_hi - h_lo) * N, 1)
                else:
                    scal = self.dev.upload(numpy.asarray(scal, dtype=numpy.float64).reshape((h_hi - h_lo) * N, 1))
                self._rho_dev[(m, "ccaa", (ci, cj))] = scal      # an input like every other density (drop_caches clears it)
            ctx.gemm_scatter(rows, 1, 1, 1.0, scal.data_ptr() + 8 * (i_lo - h_lo) * N, 1, one, 2, H, off.data_ptr() + 8 * row0, 0, None, True)
            held_pos.append(info.pos[ci][i_lo:i_hi])
        held_pos = numpy.concatenate(held_pos)
        diag = self._index(held_pos.astype(numpy.int64) * (info.dim + 1))
        ctx.scatter_const(H, diag, len(held_pos), float(nuc[m, m]), True)
        return H

    def H2(self, m1, m2):
        """H2[m1][m2] (dense, host ndarray) in test_H.py:101-107 ordering."""
        return self.dev.download(self.H2_device(m1, m2))

    def H2_device(self, m1, m2, bra_range=None, out=None):
        """Device tensor [(hi-lo)*dim2, dim1*dim2]: rows of H2[m1][m2] whose fragment-1 bra state has
        matrix position in bra_range=[lo,hi) (default: all).  Rows outside every charge-allowed class
        stay zero (build_H.py:65)."""
        f1, f2 = self._frag(m1), self._frag(m2)
        ctx = self.dev.ctx
        lo, hi = (0, f1.dim) if bra_range is None else bra_range
        D = f1.dim * f2.dim
        if out is None:
            out = self.dev.zeros(((hi - lo) * f2.dim, D))
        else:
            assert out.shape == ((hi - lo) * f2.dim, D)
            ctx.memset_zero(out, out.numel() * 8)        # (a library call, so that a recorded build contains it)
        for d1, c1, c2, A, B, K, ld in self._dimer_class_factors(m1, m2, (lo, hi) if bra_range is not None else None):
            off1 = self._index(lambda: c1.offsets(f1, f2.dim * D, f2.dim, bra_base=lo), ("off1", m1, m2, d1, lo, hi))
            off2 = self._index(lambda: c2.offsets(f2, D, 1), ("off2", m1, m2, d1))
            with self._timed(self, "dimer_class_d%+d" % d1, 2.0 * c1.P * c2.P * K, 8.0 * (c1.P * c2.P + (c1.P + c2.P) * K)):
                ctx.gemm_scatter(c1.P, c2.P, K, 1.0, A, ld, B, ld, out, off1, 0, off2, False)
        return out

    def H2_moments_device(self, m1, m2, shard=(0, 1), group=None, inspect=None):
        """Streamed dimer block: device tensor [5, 2] = (sum, sum of squares) of every element of each charge-transfer
        class of H2[m1][m2], formed tile by tile and consumed on chip (xr_gemm_reduce) -- for blocks that cannot be
        stored (1e12 elements at 1000 states/fragment).  With shard=(rank, world) each rank builds the factor rows of ITS
        bra states of BOTH fragments (_shard_range; so it only ever needs that part of the densities); the fragment-2 factor slabs are
        exchanged by one NCCL all-gather per class -- the one real exchange step of the path -- and every rank then
        streams its slab of rows against all columns.  Sum the results over ranks.  inspect(d1, A, B, P1, P2, K), if given,
        sees the factors of every class after the exchange (used by the tests' Gram-matrix identity at sizes where the
        block itself cannot exist)."""
        import torch.distributed as dist
        rank, world = shard
        f1, f2 = self._frag(m1), self._frag(m2)
        ctx = self.dev.ctx
        r1, r2 = self._shard_range(m1, rank, world), self._shard_range(m2, rank, world)
        moments = self.dev.zeros((5, 2))
        widest = None
        if world > 1:       # rows of the widest fragment-2 factor slab of every class (equal-size slabs for the collective)
            mine = [_PairClass(f2, -d1, r2).P for d1 in (-2, -1, 0, 1, 2)]
            widest = self.dev.upload(numpy.array(mine, dtype=numpy.int64), numpy.int64)
            dist.all_reduce(widest, op=dist.ReduceOp.MAX, group=group)
            widest = self.dev.download(widest)
        for d1, c1, c2, A, B, K, ld in self._dimer_class_factors(m1, m2, r1, r2, skip_empty=False):
            n_cols = c2.P
            if world > 1:
                rows = int(widest[d1 + 2])       # zero rows pad the narrower slabs (they add nothing to either moment)
                if rows == 0:
                    continue
                mine = B
                if c2.P != rows:
                    mine = self.dev.zeros((rows, ld))
                    if c2.P:
                        ctx.copy2d_scaled(mine, ld, B, ld, c2.P, ld, 1.0)
                B = self.dev.empty((world * rows, ld))
                dist.all_gather_into_tensor(B, mine, group=group)
                n_cols = world * rows
            if c1.P == 0 or n_cols == 0:
                continue
            if inspect is not None:
                inspect(d1, A, B, c1.P, n_cols, K)
            with self._timed(self, "dimer_stream_d%+d" % d1, 2.0 * c1.P * n_cols * K, 8.0 * (c1.P + n_cols) * K):
                ctx.gemm_reduce(c1.P, n_cols, K, 1.0, A, ld, B, ld, moments.data_ptr() + 16 * (d1 + 2))
        return moments

    def _shard_range(self, m, rank, world):
        """The bra states of fragment m this rank streams: the slab it holds when the densities are sharded (held=),
        else an equal share of EVERY charge sector (balanced: each rank gets the same pair count in every class)."""
        if m in self._held:
            return self._held[m]
        if world == 1:
            return None
        from .distributed import balanced_shard
        return balanced_shard(self._frag(m).n_states, rank, world)

    def H2_moments(self, m1, m2, shard=(0, 1)):
        out = self.dev.download(self.H2_moments_device(m1, m2, shard))
        return float(out[:, 0].sum()), float(out[:, 1].sum())

    def _dimer_class_factors(self, m1, m2, bra_range1=None, bra_range2=None, skip_empty=True):
        """Yields (d1, class of fragment 1, class of fragment 2, A [P1, ld], B [P2, ld], K, ld) for the five
        charge-transfer classes of build_H.py:69-100; H2[(i1,i2),(j1,j2)] = sum_k A[(i1,j1),k] B[(i2,j2),k]."""
        rho, T, U, V, nuc, n_elec = self.data
        f1, f2 = self._frag(m1), self._frag(m2)
        n1, n2 = f1.n_orb, f2.n_orb
        ctx = self.dev.ctx
        s2 = lambda c2: _parity(f2.n_elec_ref, c2)

        for d1 in (-2, -1, 0, 1, 2):
            c1 = _PairClass(f1, d1, bra_range1)
            c2 = _PairClass(f2, -d1, bra_range2)
            if (c1.P == 0 or c2.P == 0) and skip_empty:
                continue            # (skip_empty=False: ranks with an empty slab still take part in the factor exchange)
            if d1 == 0:
                K = n2 * n2 + 2
                ld = _even(K)
                A, B = self.dev.zeros((c1.P, ld)), self.dev.zeros((c2.P, ld))
                # 4 * sum_pr ca1[p,r] V[p,q,r,s] -> columns (q,s)
                Vt = self._ints(("ExEx", m1, m2), lambda: V[m1, m2, m1, m2].transpose(1, 3, 0, 2).reshape(n2 * n2, n1 * n1))
                self._fill_contracted(A, ld, 0, m1, "ca", c1, Vt, scale=4.0)
                u1 = self._ints(("extPot", m2, m1), lambda: U[m2, m1, m1].reshape(1, n1 * n1))
                self._fill_contracted(A, ld, n2 * n2 + 1, m1, "ca", c1, u1)
                d1rows = c1.diagonal_rows(f1)
                if len(d1rows):
                    ctx.scatter_const(A, self._index(d1rows * ld + n2 * n2), len(d1rows), 1.0, False)
                    ctx.scatter_const(A, self._index(d1rows * ld + n2 * n2 + 1), len(d1rows), float(nuc[m1, m2]), True)
                self._fill_raw(B, ld, 0, m2, "ca", c2)
                u2 = self._ints(("extPot", m1, m2), lambda: U[m1, m2, m2].reshape(1, n2 * n2))
                self._fill_contracted(B, ld, n2 * n2, m2, "ca", c2, u2)
                d2rows = c2.diagonal_rows(f2)
                if len(d2rows):
                    ctx.scatter_const(B, self._index(d2rows * ld + n2 * n2 + 1), len(d2rows), 1.0, False)
            elif d1 == -2:
                K = n2 * n2
                ld = _even(K)
                A, B = self.dev.zeros((c1.P, ld)), self.dev.zeros((c2.P, ld))
                # sum_pq cc1[p,q] V[p,q,r,s] -> column (s,r) to meet aa2[s,r]
                Vt = self._ints(("2min2pls", m1, m2), lambda: V[m1, m1, m2, m2].transpose(3, 2, 0, 1).reshape(n2 * n2, n1 * n1))
                self._fill_contracted(A, ld, 0, m1, "cc", c1, Vt)
                self._fill_raw(B, ld, 0, m2, "aa", c2)
            elif d1 == +2:
                K = n1 * n1
                ld = _even(K)
                A, B = self.dev.zeros((c1.P, ld)), self.dev.zeros((c2.P, ld))
                Vt = self._ints(("2min2pls", m2, m1), lambda: V[m2, m2, m1, m1].transpose(3, 2, 0, 1).reshape(n1 * n1, n2 * n2))
                self._fill_raw(A, ld, 0, m1, "aa", c1)
                self._fill_contracted(B, ld, 0, m2, "cc", c2, Vt)
            else:
                # x creates (delta_x = -1), y annihilates; A columns: [n_y | n_x] meeting B's [a_y-part | c_x-part]
                x, y = (m1, m2) if d1 == -1 else (m2, m1)
                fx, fy = self._frag(x), self._frag(y)
                nx, ny = fx.n_orb, fy.n_orb
                K = ny + nx
                ld = _even(K)
                A, B = self.dev.zeros((c1.P, ld)), self.dev.zeros((c2.P, ld))
                hT = self._ints(("h1e", m1, m2, x, y), lambda: (T[x, y] + U[m1, x, y] + U[m2, x, y]).T.copy())          # [q, p]
                V1 = self._ints(("1112", x, y), lambda: V[x, x, x, y].transpose(3, 1, 0, 2).reshape(ny, nx ** 3))      # [s,(q,p,r)]
                V2 = self._ints(("1222", x, y), lambda: V[x, y, y, y].transpose(0, 1, 3, 2).reshape(nx, ny ** 3))      # [p,(q,s,r)]
                sign = (lambda c: s2(c)) if d1 == -1 else (lambda c: -s2(c))
                cx, cy = (c1, c2) if d1 == -1 else (c2, c1)
                X, Y = (A, B) if d1 == -1 else (B, A)
                sx = sign if x == m2 else None      # the sign rides on fragment 2's factor rows
                sy = sign if y == m2 else None
                # creator side: [ h^T.c + 2 V1112.cca | c ]
                self._fill_contracted(X, ld, 0, x, "c", cx, hT, sign=sx)
                self._fill_contracted(X, ld, 0, x, "cca", cx, V1, scale=2.0, sign=sx, accumulate=True)
                self._fill_raw(X, ld, ny, x, "c", cx, sign=sx)
                # annihilator side: [ a | 2 V1222.caa ]
                self._fill_raw(Y, ld, 0, y, "a", cy, sign=sy)
                self._fill_contracted(Y, ld, ny, y, "caa", cy, V2, scale=2.0, sign=sy)
            yield d1, c1, c2, A, B, K, ld

    # ------------------------------------------------------------------------------ trimers
    def _trimer_classes(self, ms):
        """Yields (role positions (k,b,c), W builder, beta op, gamma op, deltas, sign info) for the 12
        charge-transfer patterns of build_H.py:116-186."""
        for k in range(3):
            o0, o1 = [o for o in range(3) if o != k]
            yield dict(kind="2min", k=k, b=o0, c=o1, dk=-2, db=+1, dc=+1, flip=False)
            yield dict(kind="2pls", k=k, b=o0, c=o1, dk=+2, db=-1, dc=-1, flip=False)
            yield dict(kind="ex", k=k, b=o0, c=o1, dk=0, db=-1, dc=+1, flip=False)     # creator before annihilator
            yield dict(kind="ex", k=k, b=o1, c=o0, dk=0, db=-1, dc=+1, flip=True)      # creator after annihilator

    def _trimer_factors(self, ms, cl, a_range=None):
        """Device factors of one trimer class: W [Pk, n_b*n_c], beta [Pb, n_b], gamma [Pc, n_c] with the
        build_H.py:113-114 signs folded into the rows, and the three pair classes."""
        rho, T, U, V, nuc, n_elec = self.data
        f = [self._frag(m) for m in ms]
        k, b, c = cl["k"], cl["b"], cl["c"]
        mk, mb, mc = ms[k], ms[b], ms[c]
        nk, nb, nc = f[k].n_orb, f[b].n_orb, f[c].n_orb
        ck, cb, cc = _PairClass(f[k], cl["dk"]), _PairClass(f[b], cl["db"]), _PairClass(f[c], cl["dc"])
        if ck.P == 0 or cb.P == 0 or cc.P == 0:
            return None
        # base = {0: s3, 1: s2*s3, 2: s2}[k]: s2 rides on the fragment at position 1, s3 on position 2
        carries = {0: (2,), 1: (1, 2), 2: (1,)}[k]
        def sign_of(pos):
            if pos in carries:
                return lambda chg: _parity(f[pos].n_elec_ref, chg)
            return None
        # the tile kernel takes ONE orbital count: fragments with different counts are padded to n = max(n_b, n_c) with zero
        # rows of the (transposed) integral block, i.e. zero columns of W, beta and gamma
        n = max(nb, nc)
        def padded(block):          # [n_b, n_c, n_k*n_k] -> [n*n, n_k*n_k]
            out = numpy.zeros((n, n, nk * nk))
            out[:nb, :nc] = block
            return out.reshape(n * n, nk * nk)
        ldw = _even(n * n)
        W = self.dev.zeros((ck.P, ldw))
        if cl["kind"] == "2min":      # 2 sum V[k,k,b,c][p,q,r,s] cc_k[q,p]
            Vt = self._ints(("t2min", mk, mb, mc), lambda: padded(V[mk, mk, mb, mc].transpose(2, 3, 1, 0).reshape(nb, nc, nk * nk)))
            self._fill_contracted(W, ldw, 0, mk, "cc", ck, Vt, scale=2.0, sign=sign_of(k))
            op_b = op_c = "a"
        elif cl["kind"] == "2pls":    # 2 sum V[b,c,k,k][r,s,p,q] aa_k[q,p]
            Vt = self._ints(("t2pls", mk, mb, mc), lambda: padded(V[mb, mc, mk, mk].transpose(0, 1, 3, 2).reshape(nb, nc, nk * nk)))
            self._fill_contracted(W, ldw, 0, mk, "aa", ck, Vt, scale=2.0, sign=sign_of(k))
            op_b = op_c = "c"
        else:                         # 4 sum V[k,cre,k,ann][p,r,q,s] ca_k[p,q] + delta_k U[k,cre,ann][r,s]
            Vt = self._ints(("tex", mk, mb, mc), lambda: padded(V[mk, mb, mk, mc].transpose(1, 3, 0, 2).reshape(nb, nc, nk * nk)))
            self._fill_contracted(W, ldw, 0, mk, "ca", ck, Vt, scale=4.0, sign=sign_of(k))
            # delta(i_k, j_k) * U[k,cre,ann][r,s] on the diagonal rows: rank-1 update through the same kernel
            drows = ck.diagonal_rows(f[k])
            if len(drows):
                sg = sign_of(k)
                dvec = numpy.zeros((ck.P, 2))
                for ci, cj, i_lo, i_hi, off in ck.sectors:
                    if ci == cj:
                        i = numpy.arange(i_lo, i_hi)
                        dvec[off + (i - i_lo) * f[k].n_states[cj] + i, 0] = 1.0 if sg is None else sg(ci)
                def padded_U():
                    out = numpy.zeros((n, n))
                    out[:nb, :nc] = U[mk, mb, mc]
                    return numpy.stack([out.reshape(n * n), numpy.zeros(n * n)], axis=1)
                Uv = self._ints(("tU", mk, mb, mc), padded_U)
                self.dev.ctx.gemm_scatter(ck.P, n * n, 1, 1.0, self.dev.upload(dvec), 2, Uv, 2, W, None, ldw, None, True)
            op_b, op_c = "c", "a"
        beta = self.dev.zeros((cb.P, _even(n)))
        gamma = self.dev.zeros((cc.P, _even(n)))
        self._fill_raw(beta, _even(n), 0, mb, op_b, cb, sign=sign_of(b))
        self._fill_raw(gamma, _even(n), 0, mc, op_c, cc, sign=sign_of(c))
        alpha = -1.0 if cl["flip"] else 1.0
        return dict(W=W, ldw=ldw, beta=beta, gamma=gamma, ck=ck, cb=cb, cc=cc, alpha=alpha, n=n, k=k, b=b, c=c)

    def H3(self, m1, m2, m3):
        """H3[m1][m2][m3] dense (host ndarray), test_H.py:113-126 ordering.  Only for sizes that fit."""
        return self.dev.download(self.H3_device(m1, m2, m3))

    def H3_device(self, m1, m2, m3):
        ms = (m1, m2, m3)
        f = [self._frag(m) for m in ms]
        D = f[0].dim * f[1].dim * f[2].dim
        if D * D * 8 > 64 * (1 << 30):
            raise MemoryError("dense H3 would need %.1f GB; use H3_moments / the tile stream" % (D * D * 8 / 2 ** 30))
        H = self.dev.zeros((D, D))
        stride = (f[1].dim * f[2].dim, f[2].dim, 1)
        ctx = self.dev.ctx
        for cl in self._trimer_classes(ms):
            fac = self._trimer_factors(ms, cl)
            if fac is None:
                continue
            offs = []
            for role, cls in ((cl["k"], fac["ck"]), (cl["b"], fac["cb"]), (cl["c"], fac["cc"])):
                offs.append(self._index(cls.offsets(f[role], stride[role] * D, stride[role])))
            ctx.trimer_stream(fac["n"], fac["ck"].P, fac["cb"].P, fac["cc"].P, fac["alpha"], fac["W"], fac["ldw"],
                              fac["beta"], fac["beta"].shape[1], fac["gamma"], fac["gamma"].shape[1], 0, fac["ck"].P,
                              _lib.TRIMER_MATERIALIZE, None, H, offs[0], offs[1], offs[2])
        return H

    def H3_moments(self, m1, m2, m3, shard=(0, 1), per_class=False):
        """Stream every element of H3[m1][m2][m3] through the on-chip reducer and return
        (sum, sum of squares) -- the consumer used when the block cannot be stored (1e13 elements at
        200 states/fragment).  shard=(rank, world) restricts to this rank's slab of each class's
        leading pair index (no communication; add the results)."""
        out = self.dev.download(self.H3_moments_device(m1, m2, m3, shard))
        if per_class:
            return out
        return float(out[:, 0].sum()), float(out[:, 1].sum())

    def H3_moments_device(self, m1, m2, m3, shard=(0, 1)):
        """device tensor [12, 2]: (sum, sum of squares) per charge-transfer class, no host sync"""
        ms = (m1, m2, m3)
        ctx = self.dev.ctx
        rank, world = shard
        moments = self.dev.zeros((12, 2))
        for idx, cl in enumerate(self._trimer_classes(ms)):
            fac = self._trimer_factors(ms, cl)
            if fac is None:
                continue
            Pa, Pb, Pc, n = fac["ck"].P, fac["cb"].P, fac["cc"].P, fac["n"]
            a_lo, a_hi = Pa * rank // world, Pa * (rank + 1) // world
            flops = 2.0 * (a_hi - a_lo) * Pb * n * n + 2.0 * (a_hi - a_lo) * Pb * Pc * n
            with self._timed(self, "trimer_stream_%s" % cl["kind"], flops, 8.0 * ((a_hi - a_lo) * n * n + (Pb + Pc) * n)):
                ctx.trimer_stream(n, Pa, Pb, Pc, fac["alpha"], fac["W"], fac["ldw"], fac["beta"], fac["beta"].shape[1],
                                  fac["gamma"], fac["gamma"].shape[1], a_lo, a_hi, _lib.TRIMER_REDUCE,
                                  moments.data_ptr() + 16 * idx, None, None, None, None)
        return moments

    # ---- consumers of the streamed trimer tiles other than the moments: sampled elements, screened sparse block ----
    def H3_elements(self, m1, m2, m3, I, J):
        """H3[m1][m2][m3] elements <I| H |J> for lists of bra / ket state triples (each state a (charge, index) pair, as in
        build_H.trimer, build_H.py:103) -- for blocks far too large to store: the requested elements are picked out of the
        streamed tiles by xr_trimer_sample.  Returns a host ndarray [len(I)]; charge-forbidden elements are 0."""
        ms = (m1, m2, m3)
        f = [self._frag(m) for m in ms]
        out = numpy.zeros(len(I))
        delta = numpy.array([[i[x][0] - j[x][0] for x in range(3)] for i, j in zip(I, J)], dtype=numpy.int64).reshape(len(I), 3)
        for cl in self._trimer_classes(ms):
            want = numpy.zeros(3, dtype=numpy.int64)
            want[cl["k"]], want[cl["b"]], want[cl["c"]] = cl["dk"], cl["db"], cl["dc"]
            mine = numpy.nonzero((delta == want).all(axis=1))[0]
            if len(mine) == 0:
                continue
            fac = self._trimer_factors(ms, cl)
            if fac is None:
                continue
            abc = numpy.empty((len(mine), 3), dtype=numpy.int64)
            for col, (role, cls) in enumerate(((cl["k"], fac["ck"]), (cl["b"], fac["cb"]), (cl["c"], fac["cc"]))):
                where = {(ci, cj): (i_lo, off) for ci, cj, i_lo, i_hi, off in cls.sectors}
                for row, t in enumerate(mine):
                    (ci, i), (cj, j) = I[t][role], J[t][role]
                    i_lo, off = where[(ci, cj)]
                    abc[row, col] = off + (i - i_lo) * f[role].n_states[cj] + j
            vals = self.dev.empty((len(mine),))
            self.dev.ctx.trimer_sample(fac["n"], fac["ck"].P, fac["cb"].P, fac["cc"].P, fac["alpha"], fac["W"], fac["ldw"],
                                       fac["beta"], fac["beta"].shape[1], fac["gamma"], fac["gamma"].shape[1], abc, vals)
            out[mine] = self.dev.download(vals)
        return out

    def H3_sparse(self, m1, m2, m3, tau, capacity=1 << 22, shard=(0, 1), classes=None, count_only=False, max_elements=1 << 28):
        """Screened H3[m1][m2][m3]: (flat indices into the dense [D, D] block in test_H.py:113-126 ordering, values) of every
        element with |H3| > tau, sorted by index (host ndarrays) -- the compaction consumer of the tile stream
        (xr_trimer_threshold).  shard=(rank, world) restricts to this rank's slab of each class's leading pair index;
        classes (indices into the 12 charge-transfer classes) restricts to those; count_only=True returns the number of
        elements above tau per class ({class index: count}) without storing anything.  A list longer than max_elements raises
        MemoryError (with the exact count) instead of trying to allocate it."""
        ms = (m1, m2, m3)
        f = [self._frag(m) for m in ms]
        D = f[0].dim * f[1].dim * f[2].dim
        stride = (f[1].dim * f[2].dim, f[2].dim, 1)
        rank, world = shard
        ctx = self.dev.ctx
        count = self.dev.zeros((1,), dtype=torch.int64)
        idx_parts, val_parts, counted = [], [], {}
        for which, cl in enumerate(self._trimer_classes(ms)):
            if classes is not None and which not in classes:
                continue
            fac = self._trimer_factors(ms, cl)
            if fac is None:
                continue
            offs = [self._index(cls.offsets(f[role], stride[role] * D, stride[role]))
                    for role, cls in ((cl["k"], fac["ck"]), (cl["b"], fac["cb"]), (cl["c"], fac["cc"]))]
            Pa = fac["ck"].P
            a_lo, a_hi = Pa * rank // world, Pa * (rank + 1) // world
            if count_only:
                ctx.trimer_threshold(fac["n"], Pa, fac["cb"].P, fac["cc"].P, fac["alpha"], fac["W"], fac["ldw"], fac["beta"],
                                     fac["beta"].shape[1], fac["gamma"], fac["gamma"].shape[1], a_lo, a_hi, tau, offs[0], offs[1],
                                     offs[2], 0, None, None, count)
                counted[which] = int(self.dev.download(count)[0])
                continue
            cap = int(capacity)
            while True:
                idx = self.dev.empty((cap,), dtype=torch.int64)
                val = self.dev.empty((cap,))
                ctx.trimer_threshold(fac["n"], Pa, fac["cb"].P, fac["cc"].P, fac["alpha"], fac["W"], fac["ldw"], fac["beta"],
                                     fac["beta"].shape[1], fac["gamma"], fac["gamma"].shape[1], a_lo, a_hi, tau, offs[0], offs[1],
                                     offs[2], cap, idx, val, count)
                kept = int(self.dev.download(count)[0])
                if kept <= cap:
                    break
                if kept > max_elements:
                    raise MemoryError("H3_sparse: %d elements of class %d are above tau = %g (max_elements = %d): raise tau"
                                      % (kept, which, tau, max_elements))
                del idx, val
                cap = kept          # the list overflowed: the count is exact, run the class again with room for all of it
            idx_parts.append(self.dev.download(idx[:kept]))
            val_parts.append(self.dev.download(val[:kept]))
        if count_only:
            return counted
        if not idx_parts:
            return numpy.zeros(0, dtype=numpy.int64), numpy.zeros(0)
        idx, val = numpy.concatenate(idx_parts), numpy.concatenate(val_parts)
        order = numpy.argsort(idx, kind="stable")
        return idx[order], val[order]

    # ------------------------------------------------------------------- flop accounting
    def algorithmic_flops(self, dimers=(), trimers=()):
        """FP64 flops of the factored algorithm: precontractions
        2*P*K*F, dimer classes 2*P1*P2*K with K = n^2 (d = 0, +-2) or 2n (d = +-1), trimer classes
        2*Pk*n^4 + 2*Pk*Pb*n^2 + 2*Pk*Pb*Pc*n.  Returns (total, {"dimer": .., "trimer": ..})."""
        tot_d = tot_t = 0.0
        for m1, m2 in dimers:
            f1, f2 = self._frag(m1), self._frag(m2)
            n1, n2 = f1.n_orb, f2.n_orb
            for d1 in (-2, -1, 0, 1, 2):
                P1, P2 = _PairClass(f1, d1).P, _PairClass(f2, -d1).P
                if P1 == 0 or P2 == 0:
                    continue
                if d1 == 0:
                    tot_d += 2.0 * P1 * P2 * n2 * n2 + 2.0 * P1 * n1 * n1 * n2 * n2
                elif abs(d1) == 2:
                    tot_d += 2.0 * P1 * P2 * n2 * n2 + 2.0 * min(P1, P2) * n1 * n1 * n2 * n2
                else:
                    tot_d += 2.0 * P1 * P2 * (n1 + n2) + 2.0 * (P1 + P2) * (n1 ** 3) * n2
        for ms in trimers:
            f = [self._frag(m) for m in ms]
            for cl in self._trimer_classes(ms):
                Pk, Pb, Pc = (_PairClass(f[cl[r]], cl[d]).P for r, d in (("k", "dk"), ("b", "db"), ("c", "dc")))
                n = max(f[cl["b"]].n_orb, f[cl["c"]].n_orb)       # the tile kernel pads both to one orbital count
                if Pk and Pb and Pc:
                    tot_t += 2.0 * Pk * n ** 4 + 2.0 * Pk * Pb * n * n + 2.0 * Pk * Pb * Pc * n
        return tot_d + tot_t, {"dimer": tot_d, "trimer": tot_t}

    def element_counts(self, dimers=(), trimers=()):
        """non-zero (charge-allowed) element counts per class, for extrapolating sampled CPU timings"""
        counts = {}
        for m1, m2 in dimers:
            f1, f2 = self._frag(m1), self._frag(m2)
            for d1 in (-2, -1, 0, 1, 2):
                counts[("dimer", d1)] = counts.get(("dimer", d1), 0) + _PairClass(f1, d1).P * _PairClass(f2, -d1).P
        for ms in trimers:
            f = [self._frag(m) for m in ms]
            for cl in self._trimer_classes(ms):
                Pk, Pb, Pc = (_PairClass(f[cl[r]], cl[d]).P for r, d in (("k", "dk"), ("b", "db"), ("c", "dc")))
                counts[("trimer", cl["kind"])] = counts.get(("trimer", cl["kind"]), 0) + Pk * Pb * Pc
        return counts
