"""One-process-per-GPU row sharding of the hermitian-XRCC build: "shard fragment-state bra blocks over
GPUs, with an all-gather only to assemble H".

Rank r owns a contiguous slab [lo, hi) of fragment 0's bra states in the ordering of ``monomer_charges[0]`` -- which is a
contiguous block of rows (global i0, all i1) of every dimer matrix in get_xr_H's final ordering, and of H1[0].  It builds
those rows from the densities of fragment 0 restricted to its bra states (``bra_slab``: views, nothing is copied on the
host; only the slab is uploaded) and the full densities of fragment 1, straight into its slice of the padded full matrix;
ONE in-place ``all_gather_into_tensor`` per assembled matrix is the only collective.

torch.distributed is plumbing (NCCL on GPUs, gloo in the CPU tests); the arithmetic is in libxr_b200.so.
"""
import torch.distributed as dist

from ..general.distributed import slab_bounds
from .tensor import DeviceTensor, as_host

_META = ("n_elec", "n_states", "n_states_bra", "KetCoeffs", "bra_offset")


def bra_slab(rho, held):
    """The density dict of one fragment restricted to the bra states ``held = {chg: (lo, hi)}`` of every charge sector
    (axis 0 of every tensor rho[op][(chg_bra, chg_ket)]); ``n_states_bra`` becomes the slab sizes and ``bra_offset`` the
    first held state, which is how XR_term places the spectator deltas."""
    out = {}
    for key, val in rho.items():
        if key in _META:
            out[key] = val
            continue
        sliced = {}
        for (ci, cj), t in val.items():
            lo, hi = held.get(ci, (0, 0))
            if hasattr(t, "terms") or hasattr(t, "factors"):
                raise NotImplementedError("row sharding of factored (lazy expression) densities")
            if isinstance(t, DeviceTensor):
                sliced[(ci, cj)] = DeviceTensor(t.buf[lo:hi], t.dev)
            else:
                sliced[(ci, cj)] = as_host(t)[lo:hi]
        out[key] = sliced
    out["n_states_bra"] = {c: held.get(c, (0, 0))[1] - held.get(c, (0, 0))[0] for c in rho["n_states"]}
    out["bra_offset"] = {c: held.get(c, (0, 0))[0] for c in rho["n_states"]}
    return out


class row_shard(object):
    """The slab of fragment 0's bra states this rank owns, and the gather that assembles row-sharded matrices."""
    def __init__(self, rho0, charges0, rank, world, group=None):
        self.rank, self.world, self.group = rank, world, group
        sizes = [rho0.get("n_states_bra", rho0["n_states"])[c] for c in charges0]
        self.dim0 = sum(sizes)
        self.lo, self.hi, self.per = slab_bounds(self.dim0, rank, world)
        self.held, beg = {}, 0
        for c, n in zip(charges0, sizes):
            a, b = min(max(self.lo - beg, 0), n), min(max(self.hi - beg, 0), n)
            self.held[c] = (a, b)
            beg += n
        self.densities0 = bra_slab(rho0, self.held)

    def buffer(self, dev, mult, n_cols):
        """(padded full matrix [world * per * mult, n_cols], zeroed view of this rank's (hi - lo) * mult rows of it);
        mult = rows per fragment-0 bra state (1 for H1[0], dim1 for dimer matrices)"""
        full = dev.empty((self.world * self.per * mult, n_cols))
        mine = full[self.rank * self.per * mult:(self.rank * self.per + (self.hi - self.lo)) * mult]
        mine.zero_()
        return full, mine

    def gather(self, full, mult):
        """in-place all-gather of the ranks' slabs; returns the assembled [dim0 * mult, n_cols] matrix (a view)"""
        if self.world > 1:
            step = self.per * mult
            dist.all_gather_into_tensor(full, full[self.rank * step:(self.rank + 1) * step], group=self.group)
        return full[:self.dim0 * mult]
