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


def _build_c99_host(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    here = os.path.dirname(os.path.abspath(__file__))
    libdir = os.path.dirname(xr.LIB_PATH)
    exe = str(tmp_path / "abi_check")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(here, "..", "include"),
                           os.path.join(here, "c_host", "abi_check.c"), "-o", exe, "-L", libdir, "-lxr_b200", "-lm",
                           "-Wl,-rpath," + libdir])
    return subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_c99_host_links_and_loads(tmp_path):
    """include/xr_b200.h is plain C (gcc -std=c99 -pedantic -Werror) and a C host with no Python/torch in the process can
    link libxr_b200.so: tests/c_host/abi_check.c reports XR_ERR_NO_DEVICE + message here (its GPU branch runs a GEMM)"""
    run = _build_c99_host(tmp_path)
    assert run.returncode == 0, run.stdout
    assert "no CPU fallback" in run.stdout


@pytest.mark.gpu
def test_c99_host_runs_its_gpu_branch(tmp_path):
    """the same C99 host on a B200: context on its own stream, device info, xr_malloc/xr_upload, one xr_gemm_scatter and a
    legacy scalar through the C ABI with no Python or torch in the process"""
    run = _build_c99_host(tmp_path)
    assert run.returncode == 0, run.stdout
    assert "gemm ok" in run.stdout and "legacy scalar ok" in run.stdout, run.stdout
