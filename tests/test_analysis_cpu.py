"""CPU tier: utils analysis helpers, deflation.Arnoldifyer / bound_pseudo and the evaluator-driven
recycling factories over the device test double (tests/analysis_checks.py)."""
import pytest

import analysis_checks as ac
import fake_device


@pytest.fixture()
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


@pytest.mark.parametrize("cplx", [False, True])
def test_angles(fake, cplx):
    ac.check_angles(cplx)


@pytest.mark.parametrize("cplx", [False, True])
def test_hegedus_and_ritz(fake, cplx):
    ac.check_hegedus_and_ritz(cplx)


def test_spectral_helpers(fake):
    ac.check_spectral_helpers()


@pytest.mark.parametrize("with_M", [False, True])
@pytest.mark.parametrize("cplx", [False, True])
def test_arnoldifyer(fake, cplx, with_M):
    ac.check_arnoldifyer(cplx, with_M)


@pytest.mark.parametrize("solver,factory", [("RecyclingCg", "RitzAprioriCg"), ("RecyclingMinres", "RitzAprioriMinres"),
                                            ("RecyclingGmres", "RitzApproxKrylov"), ("RecyclingMinres", "RitzApproxKrylov")])
def test_evaluator_recycling(fake, solver, factory):
    ac.check_evaluator_recycling(solver, factory)


def test_ritz_factory_options(fake):
    ac.check_ritz_factory_options()


def test_device_linear_operator(fake):
    ac.check_device_linear_operator()


def test_timings_wall_clock_fallback(fake):
    ac.check_timings(False)
