"""Packs diagram blocks into XR matrices -- drop-in for hermitian-XRCC/XR_term.py
(``monomer_matrix`` :96-115, ``dimer_matrix`` :117-165, same positional signatures, ndarray result in
the same charge-blocked ordering: rows for each (chg1, chg2) of ``charge_blocks``, (i1, i2) row-major).

The reference materialises every diagram block on the host, transposes it and slice-adds it (with a
Python loop over spectator states for the Kronecker deltas, :69-80).  Here the matrix lives in HBM and
every diagram is accumulated in place by the epilogue of its own GEMM; the result is downloaded once.

``bra_det`` / ``ket_det`` (the StateSpaceOptimizer gradient variants, XR_term.py:122-148): the ket (bra) state
pair of every two-fragment diagram is traced and the result is a VECTOR over the bra (ket) product states, summed
over all charge blocks of the traced side.

Extra keyword arguments (defaults reproduce the reference):
  ordering="blocked" | "final" -- "final" lays rows/columns out as (global i1, global i2), i.e. what
      get_xr_result.py:300-353 produces from the blocked matrix with a double Python loop;
  device_result=False          -- return the device tensor instead of a host ndarray;
  into=None, scale=1.0         -- accumulate scale * (this term) into an existing device matrix instead of a
      fresh zero one (how get_xr_H sums and subtracts terms without any elementwise pass).

Bra slabs (multi-GPU row sharding, hermitian/distributed.py): a density dict that holds only the bra states [lo, hi) of
every charge sector says so with ``rho["bra_offset"] = {chg: lo}`` (and ``n_states_bra`` = the slab sizes).  The rows
of the matrix are then the slab's states and the Kronecker deltas of spectator fragments (XR_term.py:69-80) connect
local bra state i with ket state lo + i.
"""
import numpy

from .tensor import DeviceTensor


def _buffer(into, shape):
    buf = into.buf if isinstance(into, DeviceTensor) else into
    if tuple(buf.shape) != tuple(shape):
        raise ValueError("into= has shape %r, expected %r" % (tuple(buf.shape), tuple(shape)))
    return buf


def _finish(dev, Matrix, device_result):
    if device_result:
        return DeviceTensor(Matrix, dev)
    return dev.download(Matrix)


def monomer_matrix(op_blocks, active_diagrams, subsys_index, charge_blocks, timings, device_result=False, into=None, scale=1.0):
    rho = op_blocks.densities[subsys_index]
    dev = op_blocks.dev
    dim_bra = sum(rho["n_states_bra"][chg] for chg in charge_blocks)
    dim_ket = sum(rho["n_states"][chg] for chg in charge_blocks)
    Matrix = dev.zeros((dim_bra, dim_ket)) if into is None else _buffer(into, (dim_bra, dim_ket))
    ld = dim_ket
    Ibeg = 0
    for chg_i in charge_blocks:
        Jbeg = 0
        for chg_j in charge_blocks:
            for frag_order, labels in active_diagrams.items():
                if frag_order != 1:
                    raise NotImplementedError("monomer_matrix with diagrams of fragment order %r" % (frag_order,))
                if chg_i != chg_j:
                    continue
                entry = op_blocks[(subsys_index,)][((chg_i, chg_j),)]
                for label in labels:
                    timings.start()
                    entry.accumulate(label, Matrix, Ibeg * ld + Jbeg, {("i", 0): ld, ("j", 0): 1}, scale)
                    timings.record("block evaluation")
            Jbeg += rho["n_states"][chg_j]
        Ibeg += rho["n_states_bra"][chg_i]
    return _finish(dev, Matrix, device_result)


def _layout(rho1, rho2, which, charge_blocks, ordering):
    """{(c1, c2): (index of the block's first product state, stride of the fragment-1 state index)}"""
    n1, n2 = rho1[which], rho2[which]
    starts = {}
    if ordering == "blocked":                       # XR_term.py:153-164
        beg = 0
        for c1, c2 in charge_blocks:
            starts[(c1, c2)] = (beg, n2[c2])
            beg += n1[c1] * n2[c2]
    elif ordering == "final":                       # get_xr_result.py:300-353
        chgs1, chgs2 = [], []
        for c1, c2 in charge_blocks:
            if c1 not in chgs1:
                chgs1.append(c1)
            if c2 not in chgs2:
                chgs2.append(c2)
        if len(chgs1) * len(chgs2) != len(charge_blocks):
            raise ValueError("ordering='final' needs charge_blocks to be a full product of monomer charges")
        off1, off2, acc = {}, {}, 0
        for c in chgs1:
            off1[c] = acc
            acc += n1[c]
        tot2 = 0
        for c in chgs2:
            off2[c] = tot2
            tot2 += n2[c]
        for c1, c2 in charge_blocks:
            starts[(c1, c2)] = (off1[c1] * tot2 + off2[c2], tot2)
    else:
        raise ValueError("ordering must be 'blocked' or 'final'")
    return starts


def dimer_matrix(op_blocks, active_diagrams, subsys_indices, charge_blocks, timings, bra_det=False, ket_det=False,
                 ordering="blocked", device_result=False, into=None, scale=1.0):
    m = tuple(subsys_indices)
    rho1, rho2 = (op_blocks.densities[x] for x in m)
    dev = op_blocks.dev
    dim_bra = sum(rho1["n_states_bra"][c1] * rho2["n_states_bra"][c2] for c1, c2 in charge_blocks)
    dim_ket = sum(rho1["n_states"][c1] * rho2["n_states"][c2] for c1, c2 in charge_blocks)
    det = "bra" if (bra_det and not ket_det) else ("ket" if (ket_det and not bra_det) else None)
    shape = (dim_bra,) if det == "bra" else ((dim_ket,) if det == "ket" else (dim_bra, dim_ket))
    Matrix = dev.zeros(shape) if into is None else _buffer(into, shape)
    ld = dim_ket
    rows = _layout(rho1, rho2, "n_states_bra", charge_blocks, ordering)
    cols = _layout(rho1, rho2, "n_states", charge_blocks, ordering)
    n_j = lambda x, chg: op_blocks.densities[m[x]]["n_states"][chg]

    def delta_range(x, chg):
        """(first ket state, count) of the states of fragment x that a Kronecker delta pairs with its local bra states"""
        rho = op_blocks.densities[m[x]]
        if "bra_offset" in rho:
            return rho["bra_offset"].get(chg, 0), rho["n_states_bra"][chg]
        return 0, n_j(x, chg)

    for chg_i in charge_blocks:
        for chg_j in charge_blocks:
            subsys_charges = [(chg_i[0], chg_j[0]), (chg_i[1], chg_j[1])]
            (Ibeg, Istride), (Jbeg, Jstride) = rows[tuple(chg_i)], cols[tuple(chg_j)]
            if det == "bra":          # vector over bra product states; every ket charge block adds into it (XR_term.py:122-133)
                base, slot = Ibeg, {("i", 0): Istride, ("i", 1): 1}
            elif det == "ket":        # vector over ket product states (XR_term.py:134-148)
                base, slot = Jbeg, {("j", 0): Jstride, ("j", 1): 1}
            else:
                base = Ibeg * ld + Jbeg
                slot = {("i", 0): Istride * ld, ("i", 1): ld, ("j", 0): Jstride, ("j", 1): 1}
            for frag_order, labels in active_diagrams.items():
                # every ascending group of `frag_order` fragments of the dimer (XR_term.py:41-53)
                groups = {0: [()], 1: [(0,), (1,)], 2: [(0, 1)]}.get(frag_order)
                if groups is None:
                    raise NotImplementedError("dimer_matrix with diagrams of fragment order %r" % (frag_order,))
                if det and frag_order != 2:
                    raise NotImplementedError("bra_det / ket_det with diagrams of fragment order %r" % (frag_order,))
                for frags in groups:
                    others = [x for x in (0, 1) if x not in frags]
                    if any(subsys_charges[x][0] != subsys_charges[x][1] for x in others):
                        continue
                    if sum(subsys_charges[x][0] for x in frags) != sum(subsys_charges[x][1] for x in frags):
                        continue
                    entry = op_blocks[tuple(m[x] for x in frags)][tuple(subsys_charges[x] for x in frags)]
                    for label in labels:
                        timings.start()
                        if frag_order == 0:
                            # diagram value (1 for "identity") on the diagonal of the block (XR_term.py:69-80)
                            value = entry[label]
                            if value is not None:
                                (lo0, n0), (lo1, n1) = delta_range(0, chg_j[0]), delta_range(1, chg_j[1])
                                a = numpy.arange(n0, dtype=numpy.int64)[:, None]
                                b = numpy.arange(n1, dtype=numpy.int64)[None, :]
                                idx = (base + lo0 * slot[("j", 0)] + lo1 * slot[("j", 1)]
                                       + a * (slot[("i", 0)] + slot[("j", 0)]) + b * (slot[("i", 1)] + slot[("j", 1)]))
                                idx = dev.upload(idx.reshape(-1), dtype=numpy.int64)
                                dev.ctx.scatter_const(Matrix, idx, n0 * n1, scale * float(value), True)
                        elif frag_order == 1:
                            x, o = frags[0], others[0]
                            lo, count = delta_range(o, chg_j[o])
                            strides = {("i", 0): slot[("i", x)], ("j", 0): slot[("j", x)],
                                       "delta": slot[("i", o)] + slot[("j", o)], "n_delta": count}
                            entry.accumulate(label, Matrix, base + lo * slot[("j", o)], strides, scale)
                        else:
                            entry.accumulate(label, Matrix, base, slot, scale)
                        timings.record("block evaluation")
    return _finish(dev, Matrix, device_result)
