"""B200 drop-in for general-XRCC/build_density_tensors.py (the step before the H build).

    densities, total_size = build_density_tensors(z_lists, n_orbs, Vints, n_core, n_threads=1)

Same arguments and return value as the reference (build_density_tensors.py:60-160): ``z_lists[chg]`` carries ``.configs``
[n_configs, n_elec] (ascending occupied spin-orbital indices, alpha 0..n_orbs-1, beta n_orbs..2 n_orbs-1) and ``.coeffs``
[n_states, n_configs]; ``densities[op][(bra_chg, ket_chg)][i][j]`` is the ndarray (2 n_orbs,)*k for op in a, c, aa, cc, ca,
caa, cca, and the scalar sum_pqrs V[p,q,r,s] ccaa[p,q,s,r] for "ccaa" (build_density_tensors.py:125-133); ``total_size``
counts the stored numbers the way the reference does.

Every tensor is produced by ONE launch of xr_density_tensor per (operator string, charge pair) with the CI vectors resident
in HBM.  With ``device_result=True`` the blocks stay on the GPU as CUDA tensors [N_bra, N_ket, n, ...] (and [N_bra, N_ket]
for ccaa) -- exactly what ``general.build_H.build_matrix_elements`` accepts, so the largest arrays of the H build are never
uploaded (or even seen by the host).
"""
import math
import numpy
import torch

_OPS_BY_DCHG = {+2: ("aa",), +1: ("a", "caa"), 0: ("ca", "ccaa"), -1: ("c", "cca"), -2: ("cc",)}     # build_density_tensors.py:78-157


def _masks(configs):
    configs = numpy.asarray(configs, dtype=numpy.int64)
    if configs.size and (configs.min() < 0 or configs.max() > 63):
        raise ValueError("orbital indices must lie in 0..63 (occupation masks are 64-bit)")
    bits = numpy.left_shift(numpy.uint64(1), configs.astype(numpy.uint64))
    return numpy.bitwise_or.reduce(bits, axis=1) if configs.shape[1] else numpy.zeros(len(configs), dtype=numpy.uint64)


def _check_enumeration(configs, n_orbs, n_core):
    """The C code indexes CI coefficients by the lexicographic rank of the valence configuration (find_config_index,
    density_tensors.c:29-64): row P of ``configs`` must BE the configuration of rank P, and all of them must be present."""
    configs = numpy.asarray(configs, dtype=numpy.int64)
    n_cfg, n_elec = configs.shape
    e, S = n_elec - 2 * n_core, 2 * (n_orbs - n_core)
    if e < 0 or math.comb(S, e) != n_cfg:
        raise ValueError("%d configurations of %d electrons: expected all C(%d,%d) valence configurations" % (n_cfg, n_elec, S, max(e, 0)))
    is_core = (configs < n_core) | ((configs >= n_orbs) & (configs < n_orbs + n_core))
    if (is_core.sum(axis=1) != 2 * n_core).any():
        raise ValueError("a configuration does not keep all core orbitals occupied")
    val = numpy.where(configs < n_orbs, configs - n_core, configs - 2 * n_core)
    val = numpy.sort(numpy.where(is_core, numpy.int64(1 << 40), val), axis=1)[:, :e]
    rank, prev = numpy.zeros(n_cfg, dtype=numpy.int64), numpy.full(n_cfg, -1, dtype=numpy.int64)
    G = numpy.zeros((max(e, 1), S + 1), dtype=numpy.int64)
    for i in range(e):
        for m in range(1, S + 1):
            G[i, m] = G[i, m - 1] + math.comb(S - m, e - i - 1)
    for i in range(e):
        rank += G[i, val[:, i]] - G[i, prev + 1]
        prev = val[:, i]
    if not numpy.array_equal(rank, numpy.arange(n_cfg)):
        raise ValueError("configurations are not listed in find_config_index order")


def build_density_tensors(z_lists, n_orbs, Vints, n_core, n_threads=1, device=None, device_result=False):
    from ..device import Device
    dev = device if device is not None else Device()
    ctx = dev.ctx
    n = 2 * n_orbs
    coeffs, masks, n_elec = {}, {}, {}
    for chg, states in z_lists.items():
        configs = numpy.asarray(states.configs, dtype=numpy.int64)
        _check_enumeration(configs, n_orbs, n_core)
        n_elec[chg] = configs.shape[1]
        coeffs[chg] = dev.upload(numpy.array(list(states.coeffs), dtype=numpy.float64))
        masks[chg] = dev.upload(_masks(configs).view(numpy.int64), numpy.int64)
    Vt = None
    densities = {op: {} for op in ("aa", "a", "caa", "ca", "ccaa", "c", "cca", "cc")}
    total_size = 0
    for bra in z_lists:
        for ket in z_lists:
            Nb, Nk = coeffs[bra].shape[0], coeffs[ket].shape[0]
            for op in _OPS_BY_DCHG.get(bra - ket, ()):
                T = n ** len(op)
                args = (Nb, Nk, coeffs[bra], coeffs[bra].shape[1], coeffs[ket], coeffs[ket].shape[1], masks[ket], n_elec[bra],
                        n_elec[ket], n_orbs, n_core)
                if op == "ccaa":        # build_density_tensors.py:125-133 keeps only sum V[p,q,r,s] ccaa[p,q,s,r]: fused, never stored
                    if Vt is None:
                        V = numpy.asarray(Vints, dtype=numpy.float64).reshape(n, n, n, n)
                        Vt = dev.upload(V.transpose(0, 1, 3, 2).reshape(-1))
                    scalars = dev.empty((Nb, Nk))
                    ctx.density_contracted(op, scalars, Vt, *args)
                    block, count = scalars, Nb * Nk
                    if not device_result:
                        host = dev.download(block)
                        block = [[float(host[i, j]) for j in range(Nk)] for i in range(Nb)]
                else:
                    rho = dev.empty((Nb * Nk, T))
                    ctx.density_tensor(op, rho, *args)
                    block, count = rho.reshape((Nb, Nk) + (n,) * len(op)), Nb * Nk * T
                    if not device_result:
                        host = dev.download(block)
                        block = [[numpy.array(host[i, j]) for j in range(Nk)] for i in range(Nb)]
                densities[op][bra, ket] = block
                total_size += count
    return densities, total_size
