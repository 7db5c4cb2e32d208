"""CPU tier twin of tests/test_lanczos_diag_gpu.py over the device test double."""
import numpy as np
import pytest

import fake_device
import test_lanczos_diag_gpu as v


@pytest.fixture()
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


@pytest.mark.parametrize("with_prev", [False, True])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_lanczos_diag_contract(fake, dt, with_prev):
    v.check_lanczos_diag_kernel(fake, 301, dt, with_prev, 1)


@pytest.mark.parametrize("name", ["shifted_minres_ipB_f64", "shifted_minres_ipB"])
def test_fused_diagonal_ipB_lanczos(fake, monkeypatch, name):
    v.check_switch_parity(monkeypatch, "_LANCZOS_DIAGB", name)
    assert fake.calls.get("lanczos_diag", 0) == 30 and "axpy_dev" not in fake.calls
