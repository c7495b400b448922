"""GPU twins of host-logic tests: the SAME test bodies as tests/test_host_logic_cpu.py (same inputs, same golden vectors /
oracle, same assertions), with the test-only NumPy device stand-in replaced by the real Device -- i.e. every product in
them is computed by libxr_b200.so on the B200.  Covers what had CPU-only coverage: xr_ccsd.build_H at every S-order
("proper", 0-4; 8-operator densities at orders 3/4), the high-order assembly in the style of mains/xr_ccsd.py, the seeded
random general systems, H1 of held bra slabs and the det variants."""
import pytest

import test_host_logic_cpu as host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from qodeapplications_b200.device import Device
    return Device(0)


@pytest.fixture()
def on_device(dev, monkeypatch):
    monkeypatch.setattr(host, "FakeDevice", lambda *a, **k: dev)
    return dev


@pytest.mark.parametrize("order", ["proper", 0, 1, 2, 3, 4])
def test_xr_ccsd_build_H_on_device(on_device, order):
    host.test_hermitian_xr_ccsd_build_H_host_logic(order)


def test_high_order_assembly_on_device(on_device):
    host.test_hermitian_high_order_assembly_in_the_style_of_xr_ccsd(4)


@pytest.mark.parametrize("seed", [101, 202, 303])
def test_general_randomised_systems_on_device(on_device, seed):
    host.test_general_randomised_systems_host_logic(seed)


def test_general_H1_of_held_bra_slabs_on_device(on_device):
    host.test_general_H1_of_held_bra_slabs_host_logic()


@pytest.mark.parametrize("which", ["bra", "ket"])
def test_det_variants_on_device(on_device, which):
    host.test_hermitian_det_variants_host_logic(which)
