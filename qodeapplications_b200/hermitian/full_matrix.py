"""The consumer of ``get_xr_H`` in hermitian-XRCC/mains/workflow.py:214-226 (and StateSpaceOptimizer/state_gradients.py:99-107):
the full two-fragment Hamiltonian  H2 + H1[0] (x) 1 + 1 (x) H1[1],  where -- exactly as the reference's slice loop does -- only
the CHARGE-DIAGONAL blocks of the monomer matrices are added, each to the dimer blocks with the same bra and ket charges.

    full = full_matrix(H1, H2, state_dict, monomer_charges)        # ndarray (dim0*dim1, dim0*dim1), states ordered by charge

On the GPU this is three ``xr_embed_add`` passes (general/hamiltonian.supersystem_matrix); H2 / H1 may be ndarrays or device
tensors (then the sum never leaves HBM: pass ``device_result=True``).
"""
import numpy
import torch

from ..general.hamiltonian import supersystem_matrix


def full_matrix(H1, H2, state_dict, monomer_charges, device=None, device_result=False):
    dims = [sum(state_dict[m][chg] for chg in monomer_charges[m]) for m in (0, 1)]
    big = supersystem_matrix(dims, device)
    dev = big.dev
    big.add((0, 1), H2)
    for m in (0, 1):
        # keep the charge-diagonal blocks only (workflow.py:221-225 adds H1[m][slice(chg), slice(chg)] per charge): each of them
        # is copied into a zeroed block on the device (xr_copy2d_scaled), so no element of H1 is touched by the host
        d = dims[m]
        block = H1[m]
        src = block if isinstance(block, torch.Tensor) else dev.upload(numpy.asarray(block, dtype=numpy.float64))
        masked = dev.zeros((d, d))
        at = 0
        for chg in monomer_charges[m]:
            n = state_dict[m][chg]
            if n:
                dev.ctx.copy2d_scaled(masked.data_ptr() + 8 * (at * d + at), d, src.data_ptr() + 8 * (at * d + at), d, n, n, 1.0)
            at += n
        big.add((m,), masked)
    return big.matrix if device_result else dev.download(big.matrix)
