"""B200 drop-in for the reference's hermitian-XRCC Hamiltonian build (XR[o]):
XR_tensor, precontract, diagrams/*, diagrammatic_expansion, XR_term, diagram_lists, get_xr_result
keep the reference's module names and call signatures; every contraction runs in libxr_b200.so."""
