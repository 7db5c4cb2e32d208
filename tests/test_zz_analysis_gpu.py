"""GPU tier twin of tests/test_analysis_cpu.py: the SURVEY 8f rank-4 additions (utils analysis
helpers, deflation.Arnoldifyer / bound_pseudo, evaluator-driven recycling) over the real kernels.
Like tests/test_zcomplex_gpu.py this file was written without GPU time left and sorts last."""
import pytest

import analysis_checks as ac

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cplx", [False, True])
def test_angles(cplx):
    ac.check_angles(cplx)


@pytest.mark.parametrize("cplx", [False, True])
def test_hegedus_and_ritz(cplx):
    ac.check_hegedus_and_ritz(cplx)


def test_spectral_helpers():
    ac.check_spectral_helpers()


@pytest.mark.parametrize("with_M", [False, True])
@pytest.mark.parametrize("cplx", [False, True])
def test_arnoldifyer(cplx, with_M):
    ac.check_arnoldifyer(cplx, with_M)


@pytest.mark.parametrize("solver,factory", [("RecyclingCg", "RitzAprioriCg"), ("RecyclingMinres", "RitzAprioriMinres"),
                                            ("RecyclingGmres", "RitzApproxKrylov")])
def test_evaluator_recycling(solver, factory):
    ac.check_evaluator_recycling(solver, factory)


def test_ritz_factory_options():
    ac.check_ritz_factory_options()


def test_device_linear_operator():
    ac.check_device_linear_operator()


def test_cuda_event_timings():
    ac.check_timings(True)

