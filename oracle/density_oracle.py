"""CPU checkers for the density-tensor build (the step before the H build).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use this module.

* ``reference_c``: the reference's own general-XRCC/density_tensors.c compiled where it lies (oracle/Makefile ->
  oracle/_ref/libdensity_tensors_ref.so), called with the argument lists general-XRCC/build_density_tensors.py:70-157 passes.
* ``tensor``: a plain-Python restatement of density_tensors.c:142-556 on occupation bit masks (small cases), in the same
  loop order (ket configuration outermost), so that even the floating-point summation order is the reference's.
* ``all_configs`` / ``make_states``: synthetic inputs -- every valence configuration over frozen cores, in the order
  find_config_index (density_tensors.c:29-64) ranks them, and seeded random orthonormal CI vectors.
"""
import ctypes
import itertools
import math
import os
import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = ("a", "c", "aa", "cc", "ca", "caa", "cca", "ccaa")


def comb_table(n_val_elec, n_val_spin_orbs):
    """build_density_tensors.py:34-40 (FCIcomboMat): C(j, i) for j >= i, zero elsewhere"""
    table = numpy.zeros((n_val_elec, n_val_spin_orbs), dtype=numpy.int64)
    for i in range(n_val_elec):
        for j in range(i, n_val_spin_orbs):
            table[i, j] = math.comb(j, i)
    return table


def config_index(config, n_elec, n_orbs, n_core, table):
    """density_tensors.c:29-64: lexicographic rank of the valence part of an ascending occupied-orbital list"""
    n_val_elec, S = n_elec - 2 * n_core, 2 * (n_orbs - n_core)
    val = []
    for idx in config:
        if n_core <= idx < n_orbs:
            val.append(idx - n_core)
        elif n_orbs + n_core <= idx < 2 * n_orbs:
            val.append(idx - 2 * n_core)
    flat = table.reshape(-1)
    index = 0
    if n_val_elec > 0:
        for n in range(1, val[0] + 1):
            index += int(flat[n_val_elec * S - n])
        for i in range(n_val_elec - 1):
            for n in range(val[i] + 2, val[i + 1] + 1):
                index += int(flat[(n_val_elec - i - 1) * S - n])
    return index


def all_configs(n_elec, n_orbs, n_core):
    """int64 [n_configs, n_elec]: every configuration with the cores doubly occupied, row P having config_index P"""
    cores = list(range(n_core)) + list(range(n_orbs, n_orbs + n_core))
    valence = [i for i in range(2 * n_orbs) if i not in cores]
    table = comb_table(n_elec - 2 * n_core, 2 * (n_orbs - n_core))
    rows = []
    for pick in itertools.combinations(valence, n_elec - 2 * n_core):
        rows.append(sorted(cores + list(pick)))
    rows.sort(key=lambda c: config_index(c, n_elec, n_orbs, n_core, table))
    out = numpy.array(rows, dtype=numpy.int64).reshape(len(rows), n_elec)
    for P, c in enumerate(out):
        assert config_index(list(c), n_elec, n_orbs, n_core, table) == P
    return out


class _states(object):
    pass


def make_states(n_orbs, n_core, n_elec_ref, n_states, seed):
    """z_lists as general-XRCC/Be631g.py:70-76 builds them: {charge: object with .configs [n_cfg, n_elec], .coeffs [n_st, n_cfg]}"""
    rng = numpy.random.default_rng(seed)
    z_lists = {}
    for chg, n_st in n_states.items():
        data = _states()
        data.configs = all_configs(n_elec_ref - chg, n_orbs, n_core)
        n_cfg = len(data.configs)
        q, _ = numpy.linalg.qr(rng.standard_normal((n_cfg, min(n_st, n_cfg))))
        data.coeffs = numpy.ascontiguousarray(q.T[:n_st])
        z_lists[chg] = data
    return z_lists


def _annihilate(a, mask):
    """density_tensors.c:80-95 on a bit mask: parity of moving orbital a to the end of the ascending list, or 0"""
    if not (mask >> a) & 1:
        return 0, mask
    x = bin(mask & ((1 << a) - 1)).count("1")
    n = bin(mask).count("1")
    return (-1) ** (n - 1 - x), mask ^ (1 << a)


def _create(c, mask):
    """density_tensors.c:97-113"""
    if (mask >> c) & 1:
        return 0, mask
    x = bin(mask & ((1 << c) - 1)).count("1")
    n = bin(mask).count("1")
    return (-1) ** (n - x), mask | (1 << c)


def tensor(op, z_lists, bra_chg, ket_chg, n_orbs, n_core):
    """rho[op][(bra_chg,ket_chg)] as one ndarray [N_bra, N_ket, dim, ...] (density_tensors.c:162-556); the operators act on
    the ket right to left, the tensor index is row-major over the operator indices in string order."""
    bra, ket = z_lists[bra_chg], z_lists[ket_chg]
    dim, k = 2 * n_orbs, len(op)
    n_elec_bra = bra.configs.shape[1]
    table = comb_table(n_elec_bra - 2 * n_core, 2 * (n_orbs - n_core))
    core_mask = sum(1 << i for i in list(range(n_core)) + list(range(n_orbs, n_orbs + n_core)))
    zb, zk = numpy.asarray(bra.coeffs, dtype=numpy.float64), numpy.asarray(ket.coeffs, dtype=numpy.float64)
    out = numpy.zeros((zb.shape[0], zk.shape[0], dim ** k))

    def walk(level, mask, parity, index, Q):
        if level < 0:
            if mask & core_mask != core_mask:
                return
            config = [i for i in range(dim) if (mask >> i) & 1]
            P = config_index(config, n_elec_bra, n_orbs, n_core, table)
            out[:, :, index] += numpy.outer(parity * zb[:, P], zk[:, Q])
            return
        for i in range(dim):
            sign, new = (_annihilate if op[level] == "a" else _create)(i, mask)
            if sign:
                walk(level - 1, new, parity * sign, index + i * dim ** (k - 1 - level), Q)

    for Q, config in enumerate(ket.configs):
        walk(k - 1, sum(1 << int(i) for i in config), 1, 0, Q)
    return out.reshape((zb.shape[0], zk.shape[0]) + (dim,) * k)


class reference_c(object):
    """the reference's compiled density_tensors.c with build_density_tensors.py:70-157's calling convention"""
    def __init__(self, path=None):
        self.lib = ctypes.CDLL(path or os.path.join(HERE, "_ref", "libdensity_tensors_ref.so"))

    def tensor(self, op, z_lists, bra_chg, ket_chg, n_orbs, n_core):
        charges = list(z_lists)
        idx = {chg: i for i, chg in enumerate(charges)}
        n_elec = numpy.array([z_lists[c].configs.shape[1] for c in charges], dtype=numpy.int64)
        n_configs = numpy.array([z_lists[c].configs.shape[0] for c in charges], dtype=numpy.int64)
        n_states = numpy.array([z_lists[c].coeffs.shape[0] for c in charges], dtype=numpy.int64)
        configs = [numpy.ascontiguousarray(z_lists[c].configs, dtype=numpy.int64) for c in charges]
        coeffs = [numpy.ascontiguousarray(z_lists[c].coeffs, dtype=numpy.float64) for c in charges]
        combos = [comb_table(int(n_elec[i]) - 2 * n_core, 2 * (n_orbs - n_core)) for i in range(len(charges))]
        ptrs = lambda arrays, ctype: (ctypes.POINTER(ctype) * len(arrays))(*[a.ctypes.data_as(ctypes.POINTER(ctype)) for a in arrays])
        dim = 2 * n_orbs
        storage = numpy.zeros(int(n_states[idx[bra_chg]] * n_states[idx[ket_chg]]) * dim ** len(op))
        fn = getattr(self.lib, op + "_tensor")
        fn.restype = None
        i64 = ctypes.c_int64
        fn(storage.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), i64(idx[bra_chg]), i64(idx[ket_chg]),
           n_elec.ctypes.data_as(ctypes.POINTER(i64)), n_states.ctypes.data_as(ctypes.POINTER(i64)), ptrs(coeffs, ctypes.c_double),
           n_configs.ctypes.data_as(ctypes.POINTER(i64)), ptrs(configs, i64), i64(n_orbs), i64(n_core), ptrs(combos, i64), i64(1))
        return storage.reshape((int(n_states[idx[bra_chg]]), int(n_states[idx[ket_chg]])) + (dim,) * len(op))


def op_dchg(op):
    return op.count("a") - op.count("c")
