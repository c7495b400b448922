"""Block-level replacement for the driver loops of general-XRCC/test_H.py:65-142.

The reference allocates ``H1[M]``, ``H2[M][N]`` (M<N, else None) and ``H3[M][N][O]`` as dense NumPy
arrays (:65-87), enumerates every (bra, ket) tuple of every fragment pair (:90-128) and maps
``compute.monomer/dimer/trimer`` over them with ``multiprocessing.Pool(30)`` (:131-142; the trimer part is
commented out there because it is intractable element by element).  ``build_H`` returns the same nested
lists, filled by whole-block GPU builds; nothing is forked (a CUDA context does not survive fork).
"""
import numpy


def build_H(compute, n_frag=None, trimers=False):
    """compute: qodeapplications_b200.general.build_H.build_matrix_elements.
    Returns (H1, H2, H3) laid out exactly as test_H.py:65-87 (H3 is None unless trimers=True)."""
    if n_frag is None:
        n_frag = len(compute.data[0])
    H1 = [compute.H1(M) for M in range(n_frag)]
    H2 = [[compute.H2(M, N) if M < N else None for N in range(n_frag)] for M in range(n_frag)]
    H3 = None
    if trimers:
        H3 = [[[compute.H3(M, N, O) if N < O else None for O in range(n_frag)] if M < N else None
               for N in range(n_frag)] for M in range(n_frag)]
    return H1, H2, H3
