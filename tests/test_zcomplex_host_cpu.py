"""CPU tier twin of tests/test_zcomplex_gpu.py: the same checks driven over the numpy test double of
the device layer (tests/fake_device.py).  Validates the host logic of the complex path (real
embedding, twin storage, interleaved coefficients) and that the GPU test file itself is sound."""
import pytest

import fake_device
import test_zcomplex_gpu as z


@pytest.fixture()
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


def test_rot90_and_twin_storage(fake):
    z.test_rot90_is_multiplication_by_i(fake, 7, z.np.float64)
    z.test_rot90_is_multiplication_by_i(fake, 1000, z.np.float32)
    z.test_rot90_on_complex_tensors_and_twin_storage(fake)


@pytest.mark.parametrize("real_valued", [False, True])
def test_givens_z_contract(fake, real_valued):
    z.test_givens_update_z_against_numpy(fake, real_valued)


def test_tri_solve_z_contract(fake):
    z.test_tri_solve_z_against_scipy(fake)


def test_embedded_operators(fake):
    z.test_embedded_operators_apply_complex_products(fake)


def test_complex_projection(fake):
    z.test_complex_projection_matches_dense_formula(fake)


@pytest.mark.parametrize("ortho", ["cgs", "cgs2"])
@pytest.mark.parametrize("name", ["z_gmres_helmholtz", "z_defl_gmres"])
def test_complex_block_gram_schmidt(fake, name, ortho):
    z.test_complex_block_gram_schmidt_matches_reference_mgs(name, ortho)


def test_complex_arnoldi(fake):
    z.test_complex_arnoldi_relation_and_orthonormality(fake)


def test_complex_convenience(fake):
    z.test_complex_convenience_wrappers(fake)


@pytest.mark.parametrize("name", z.cases.COMPLEX_CASES)
def test_complex_cases(fake, name):
    z.test_complex_cases_match_reference_fixture_and_oracle(name)


@pytest.mark.parametrize("cplx", [False, True])
def test_householder_arnoldi(fake, cplx):
    z.test_householder_arnoldi(fake, cplx)


def test_real_system_promotion(fake):
    z.test_real_system_becomes_complex_for_complex_x0_and_deflation_vectors(fake)
