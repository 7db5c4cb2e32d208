"""CPU tier: the serial cores of the complex device recurrences (krypy_b200/csrc/kry_small_core.h --
the code givens_z_kernel / tri_solve_z_kernel execute in their single-thread sections) compiled
for the host with g++ and checked against scipy's BLAS (zrotg/drotg, krypy/utils.py:419-427) and
numpy/scipy linear algebra (krypy/linsys.py:946, 982-993)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
import scipy.linalg.blas as blas

HERE = os.path.dirname(os.path.abspath(__file__))
D = ctypes.c_double
PD = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def core(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("core") / "libkry_small_core_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "csrc", "small_core_host.cpp")])
    lib = ctypes.CDLL(out)
    lib.host_zrotg.argtypes = [D, D, D, D, PD]
    lib.host_drotg.argtypes = [D, D, PD]
    lib.host_givens_step.argtypes = [ctypes.c_int, PD, PD, PD, PD]
    lib.host_givens_step.restype = D
    lib.host_tri_solve.argtypes = [ctypes.c_int, PD, ctypes.c_longlong, PD]
    return lib


def _p(a):
    return a.ctypes.data_as(PD)


def _zrotg(lib, a, b):
    o = np.zeros(3)
    lib.host_zrotg(a.real, a.imag, b.real, b.imag, _p(o))
    return o[0], complex(o[1], o[2])


def test_zrotg_matches_blas(core):
    rng = np.random.default_rng(0)
    pairs = [(complex(*rng.standard_normal(2)), complex(*rng.standard_normal(2))) for _ in range(500)]
    pairs += [(complex(*rng.standard_normal(2)) * 10.0 ** e, complex(*rng.standard_normal(2)) * 10.0 ** f)
              for e, f in [(200, 200), (-200, -200), (150, -150), (-150, 150), (0, -170), (-170, 0), (300, 10)]]
    pairs += [(0j, 2 + 0j), (0j, -2j), (0j, 1 - 1j), (2 + 0j, 0j), (0j, 0j), (3j, 4 + 0j), (-3 + 0j, 4j)]
    for a, b in pairs:
        c, s = _zrotg(core, a, b)
        mags = [abs(z) for z in (a, b) if z != 0]
        if all(1e-100 < v < 1e100 for v in mags):  # (OpenBLAS' zrotg itself is inaccurate at extreme scales)
            cr, sr = blas.zrotg(a, b)
            assert abs(c - cr.real) <= 4e-16 and abs(s - sr) <= 4e-16, (a, b, c, s, cr, sr)
        assert abs(c * c + abs(s) ** 2 - 1.0) <= 1e-15
        # the defining property: G [a, b]^T = [r, 0]^T with |r| = hypot(|a|, |b|)
        if a != 0 or b != 0:
            scale = max(abs(a), abs(b))
            z = -np.conj(s) * (a / scale) + c * (b / scale)
            assert abs(z) <= 4e-16, (a, b, z)
            if a != 0:                                 # r = c a + s b has the phase of a (LAPACK zrotg)
                r = c * (a / scale) + s * (b / scale)
                assert abs(r / abs(r) - a / abs(a)) <= 1e-15


def test_drotg_matches_blas_and_reference_table(core):
    tab = [(3, 4, .6, .8), (-3, 4, -.6, .8), (3, -4, -.6, .8), (-4, 3, .8, -.6), (0, 2, 0, 1), (0, -2, 0, 1),
           (-2, 0, 1, 0), (0, 0, 1, 0)]                                   # SURVEY a10 (measured convention)
    o = np.zeros(2)
    for a, b, c, s in tab:
        core.host_drotg(float(a), float(b), _p(o))
        assert np.allclose(o, [c, s], rtol=0, atol=1e-16), (a, b, o)
    rng = np.random.default_rng(1)
    for _ in range(300):
        a, b = rng.standard_normal(2) * 10.0 ** rng.integers(-200, 200)
        core.host_drotg(a, b, _p(o))
        assert np.allclose(o, blas.drotg(a, b), rtol=4e-16, atol=0)


@pytest.mark.parametrize("real_valued", [False, True])
def test_givens_steps_reproduce_hessenberg_qr(core, real_valued):
    """drive kryc_givens_step column by column like Gmres._solve (linsys.py:975-993) and compare R, y and
    the residual norms with a dense least-squares solve of the same Hessenberg system"""
    rng = np.random.default_rng(2)
    m = 12
    H = np.triu(rng.standard_normal((m + 1, m)) + (0 if real_valued else 1j) * rng.standard_normal((m + 1, m)), -1)
    H = H.astype(np.complex128)
    for k in range(m):
        H[k + 1, k] = abs(H[k + 1, k])                 # a norm: real, >= 0
    beta = 1.7
    rot = np.zeros(4 * m)
    y = np.zeros(2 * (m + 2))
    y[0] = beta
    R = np.zeros((m + 1, m), dtype=np.complex128)
    for k in range(m):
        r = np.zeros(2 * (k + 2))
        r[:] = H[: k + 2, k].copy().view(np.float64)
        rn = np.zeros(4)
        y2 = y[2 * k: 2 * k + 4].copy()
        res = core.host_givens_step(k, _p(r), _p(rot), _p(rn), _p(y2))
        rot[4 * k: 4 * k + 4] = rn
        y[2 * k: 2 * k + 4] = y2
        R[: k + 2, k] = r.view(np.complex128)
        assert rn[1] == (1.0 if real_valued else 0.0)              # drotg branch only for real-valued pairs
        e1 = np.zeros(k + 2, dtype=np.complex128)
        e1[0] = beta
        sol, resid, _, _ = np.linalg.lstsq(H[: k + 2, : k + 1], e1, rcond=None)
        want = np.linalg.norm(H[: k + 2, : k + 1] @ sol - e1)
        assert abs(res - want) <= 1e-13 * beta, (k, res, want)
        assert abs(R[k + 1, k]) <= 1e-14 * np.abs(H).max()
        # back substitution with the accumulated R and y gives the least-squares solution
        x = y[: 2 * (k + 1)].copy()
        Rk = np.ascontiguousarray(R[: k + 1, : k + 1])
        core.host_tri_solve(k + 1, _p(Rk.view(np.float64)), k + 1, _p(x))
        assert np.allclose(x.view(np.complex128), sol, rtol=1e-10, atol=1e-12)


def test_tri_solve_matches_scipy_with_leading_dimension(core):
    rng = np.random.default_rng(3)
    k, ld = 9, 14
    R = np.zeros((k, ld), dtype=np.complex128)
    R[:, :k] = np.triu(rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))) + 3 * np.eye(k)
    R[2, 2] = 1e-3j                                    # |im| > |re| branch of the complex division
    y = rng.standard_normal(k) + 1j * rng.standard_normal(k)
    x = y.copy().view(np.float64)
    core.host_tri_solve(k, _p(R.view(np.float64)), ld, _p(x))
    want = scipy.linalg.solve_triangular(R[:, :k], y)
    assert np.allclose(x.view(np.complex128), want, rtol=1e-13, atol=0)
