"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/xr_b200.h declares; the product refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import pytest
import torch

from qodeapplications_b200 import lib as xr


def test_library_exports_every_declared_symbol():
    lib = xr.load()
    names = xr.declared_symbols()
    assert len(names) >= 27
    for legacy in xr.LEGACY_SYMBOLS:
        assert legacy in names
    for name in names:
        assert hasattr(lib, name), "libxr_b200.so does not export %s" % name


def test_version_string():
    assert b"sm_100a" in xr.load().xr_version()


def test_prototypes_cover_header():
    declared = set(xr.declared_symbols())
    bound = set(xr._PROTOTYPES) | set(xr.LEGACY_SYMBOLS) | set(xr.LEGACY_DENSITY_SYMBOLS)
    assert declared == bound, declared ^ bound


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(xr.XRError):
        xr.Context(0)
    from qodeapplications_b200.device import Device
    with pytest.raises(xr.XRError):
        Device(0)
    # legacy scalar ABI: NaN + error message, never a CPU result
    import numpy
    from qodeapplications_b200.general.H_contractions import import_C
    contract = import_C("H_contractions")
    with pytest.raises(xr.XRError):
        contract.monomer_1e(2, numpy.ones((2, 2)), numpy.ones((2, 2)))
