"""B200 drop-in for the reference's hermitian-XRCC Hamiltonian build (XR[o]):
XR_tensor, precontract, diagrams/*, diagrammatic_expansion, XR_term, diagram_lists, get_xr_result
keep the reference's module names and call signatures; every contraction runs in libxr_b200.so."""


def install_aliases():
    """Register this package's modules under the TOP-LEVEL names the reference's callers import
    (hermitian-XRCC/mains/xr_ccsd.py, StateSpaceOptimizer/*: ``import XR_term``, ``import
    diagrammatic_expansion``, ``from diagrams import SV_diagrams``, ``from get_xr_result import
    get_xr_H``, ``import XR_tensor`` ...), so those scripts pick up the B200 implementation unchanged.
    Call once before the reference's modules are imported."""
    import sys
    from . import XR_tensor, XR_term, diagrammatic_expansion, precontract, get_xr_result, diagram_lists, diagrams
    from .diagrams import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams, build_diagram
    top = {"XR_tensor": XR_tensor, "XR_term": XR_term, "diagrammatic_expansion": diagrammatic_expansion,
           "precontract": precontract, "get_xr_result": get_xr_result, "diagram_lists": diagram_lists, "diagrams": diagrams,
           "diagrams.S_diagrams": S_diagrams, "diagrams.ST_diagrams": ST_diagrams, "diagrams.SU_diagrams": SU_diagrams,
           "diagrams.SV_diagrams": SV_diagrams, "diagrams.build_diagram": build_diagram}
    sys.modules.update(top)
    return sorted(top)
