"""Empty stand-in: hermitian-XRCC/get_xr_result.py:28 imports `excitonic` (the solver,
which needs qode.many_body) but the H build never calls it."""
