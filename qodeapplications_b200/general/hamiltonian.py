"""B200 drop-in for general-XRCC/hamiltonian.py (and its two-body subset hermitian-XRCC/hamiltonian.py): the consumer
of the H build that expands the fragment blocks into the supersystem matrix,

    Hmat[I,J] = sum_M H1[M][b_M,k_M] d(rest) + sum_{M<N} H2[M][N][(b_M b_N),(k_M k_N)] d(rest)
              + sum_{M<N<O} H3[M][N][O][(b_M b_N b_O),(k_M k_N k_O)] d(rest)          (trimer terms only where >= 2 of M,N,O change)

with I = (b_0 .. b_{F-1}) row-major, d(rest) = all other fragments in the same state in bra and ket
(hamiltonian.py:21-84: the 0/1/2/3-transition branches are exactly these Kronecker embeddings; the trimer couplings are
never read for 0 or 1 transitions, hamiltonian.py:33-43).  On the GPU every block is one table-driven pass of
``xr_embed_add``; blocks that are already device tensors (``build_matrix_elements.H2_device`` ...) never leave HBM.

    braket_loops(Hmat, N_frag, states_per_frag, H)        # the reference's call, hamiltonian.py:82 / excitonic.py:66
    supersystem_matrix(dims, device).add(frags, block)    # the same for unequal fragment dimensions / device tensors
"""
import itertools
import numpy
import torch


def _embedding(dims, frags):
    """position of every joint state of `frags` (row-major over their dims) in the supersystem index, other digits zero"""
    strides = [int(numpy.prod(dims[m + 1:], dtype=numpy.int64)) for m in range(len(dims))]
    if not frags:
        return numpy.zeros(1, dtype=numpy.int64)
    digits = numpy.indices([dims[m] for m in frags], dtype=numpy.int64).reshape(len(frags), -1)
    return sum(digits[t] * strides[m] for t, m in enumerate(frags))


class supersystem_matrix(object):
    """Device-resident Hmat[prod(dims), prod(dims)] and the embedding of fragment blocks into it."""
    def __init__(self, dims, device=None, out=None):
        from ..device import Device
        self.dims = [int(d) for d in dims]
        self.dim = int(numpy.prod(self.dims, dtype=numpy.int64))
        self.dev = device if device is not None else Device()
        self.matrix = out if out is not None else self.dev.zeros((self.dim, self.dim))
        if tuple(self.matrix.shape) != (self.dim, self.dim) or not self.matrix.is_contiguous():
            raise ValueError("Hmat must be a contiguous (%d, %d) matrix" % (self.dim, self.dim))
        self._tables = {}

    def _offsets(self, frags):
        if frags not in self._tables:
            sub = _embedding(self.dims, frags)
            rest = tuple(m for m in range(len(self.dims)) if m not in frags)
            spect = _embedding(self.dims, rest) * (self.dim + 1)
            up = lambda a: self.dev.upload(a, numpy.int64)
            self._tables[frags] = (up(sub * self.dim), up(sub), up(spect) if len(spect) > 1 else None, len(spect))
        return self._tables[frags]

    def add(self, frags, block, min_transitions=0, alpha=1.0):
        """Hmat += alpha * block (x) identity on the other fragments; block is [prod dims[frags]]^2 (ndarray or device tensor)"""
        frags = tuple(int(m) for m in frags)
        if list(frags) != sorted(set(frags)):
            raise ValueError("fragments must be distinct and ascending: %r" % (frags,))
        d = int(numpy.prod([self.dims[m] for m in frags], dtype=numpy.int64))
        if not isinstance(block, torch.Tensor):
            block = self.dev.upload(numpy.asarray(block, dtype=numpy.float64))
        if tuple(block.shape) != (d, d) or block.stride(1) != 1:
            raise ValueError("block for fragments %r must be a (%d, %d) row-major matrix" % (frags, d, d))
        offR, offC, offS, S = self._offsets(frags)
        self.dev.ctx.embed_add(self.matrix, block, block.stride(0), d, d, S, offR, offC, offS,
                               dims_sub=[self.dims[m] for m in frags], min_transitions=min_transitions, alpha=alpha)
        return self

    def add_all(self, H):
        """H = (monomer_Hamiltonians, dimer_Couplings[, trimer_Couplings]) as hamiltonian.py:24 unpacks it"""
        F = len(self.dims)
        for M in range(F):
            self.add((M,), H[0][M])
        for M, N in itertools.combinations(range(F), 2):
            self.add((M, N), H[1][M][N])
        if len(H) > 2:
            for M, N, O in itertools.combinations(range(F), 3):
                self.add((M, N, O), H[2][M][N][O], min_transitions=2)
        return self


def braket_loops(Hmat, N_frag, states_per_frag, H, device=None):
    """hamiltonian.py:82-83: accumulate the expansion of H into Hmat (ndarray, updated in place like the reference's
    ``Hmat[I,J] +=``; or a CUDA tensor, then nothing is copied to the host)."""
    dims = [states_per_frag] * N_frag
    if isinstance(Hmat, torch.Tensor):
        supersystem_matrix(dims, device, out=Hmat).add_all(H)
        return
    big = supersystem_matrix(dims, device).add_all(H)
    Hmat += big.dev.download(big.matrix)
