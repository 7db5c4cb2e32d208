"""Parity at (near-)BASELINE sizes through size-independent properties (the oracle needs
minutes there): explicit residuals recomputed independently on the host, GMRES monotonicity,
Arnoldi relation and orthonormality on the device, SpMV linearity and checksum, truncated
histories against the oracle where the CPU can still follow."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(fn):
    import krypy_b200 as kp
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return fn()
        except kp.utils.ConvergenceError as e:
            return e.solver


def test_c2_full_size_gmres30_properties():
    """BASELINE config 2 at full size: N = 9,998,244"""
    import krypy_b200 as kp
    from krypy_b200 import problems, utils as u
    n = 3162
    A = problems.laplace2d(n)
    b = problems.rhs_normal(n * n)
    ls = kp.linsys.LinearSystem(A, b)
    sol = _solve(lambda: kp.linsys.Gmres(ls, maxiter=30, tol=1e-12, ortho="cgs", store_arnoldi=True))
    rn = np.array(sol.resnorms)
    assert len(rn) == 31
    # reference history of this exact system (measured with the reference)
    np.testing.assert_allclose(rn[:4], [1.0, 0.4472050679842598, 0.2862284374501589, 0.21061040340762568], rtol=1e-10)
    assert abs(rn[-1] - 0.025770869513655478) <= 1e-10 * rn[-1] + 1e-13
    assert np.all(np.diff(rn[:-1]) <= 0)                       # minimal residual property
    # last entry is the explicit residual (linsys.py:453): recompute it on the host
    x = sol.xk.reshape(-1)
    ex = np.linalg.norm(b.reshape(-1) - A @ x) / np.linalg.norm(b)
    assert abs(ex - rn[-1]) <= 1e-12 * rn[-1] + 1e-14
    # orthonormality and Arnoldi relation of the stored basis, on the device
    ar = sol.arnoldi
    G = u._inner_dev(ar._Vd[:31], ar._Vd[:31]).cpu().numpy()
    assert np.abs(G - np.eye(31)).max() < 1e-11
    H = sol.H
    rows = np.random.default_rng(1).choice(n * n, size=2000, replace=False)
    V = ar._Vd[:31][:, rows].cpu().numpy().T                    # (2000, 31) sample of rows
    AV = np.stack([(A[rows] @ ar._Vd[j].cpu().numpy()) for j in range(0, 30, 7)], axis=1)
    VH = V @ H[:, list(range(0, 30, 7))]
    assert np.abs(AV - VH).max() < 1e-11 * np.abs(H).max()


def test_c3_like_cg_jacobi_against_oracle_and_explicit_residual():
    """config 3 shape at N = 8M: first iterations against the oracle, explicit residual at the end"""
    import krypy_b200 as kp
    from krypy_b200 import problems
    from oracle import krylov_oracle as ko
    n = 200
    A = problems.poisson3d(n)
    b = problems.rhs_normal(n ** 3)
    M = problems.jacobi_csr(A)
    ls = kp.linsys.LinearSystem(A, b, M=M, self_adjoint=True, positive_definite=True)
    sol = _solve(lambda: kp.linsys.Cg(ls, tol=1e-8, maxiter=12))
    try:
        ref = ko.cg(ko.System(A, b, M=M), tol=1e-8, maxiter=12)
    except ko.OracleConvergenceError as e:
        ref = e.result
    got, want = np.array(sol.resnorms), np.array(ref.resnorms)
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= 1e-10 * want + 1e-13)
    x = sol.xk.reshape(-1)
    r = b.reshape(-1) - A @ x
    ex = np.sqrt(r @ (M @ r)) / np.sqrt(b.reshape(-1) @ (M @ b.reshape(-1)))
    assert abs(ex - got[-1]) <= 1e-12 * got[-1]


def test_c4_like_deflated_gmres_properties():
    """config 4 shape at N = 1M, d = 20: E, C and the deflated residual"""
    import krypy_b200 as kp
    from krypy_b200 import problems, utils as u
    n, d = 1000, 20
    A = problems.convdiff2d(n, c=0.1)
    b = np.ones((n * n, 1))
    U = np.random.default_rng(2).standard_normal((n * n, d))
    ls = kp.linsys.LinearSystem(A, b)
    sol = _solve(lambda: kp.deflation.DeflatedGmres(ls, U=U, maxiter=25, tol=1e-10, ortho="cgs", store_arnoldi=True))
    pr = sol.projection
    G = u._inner_dev(pr._Ud, pr._Ud).cpu().numpy()
    assert np.abs(G - np.eye(d)).max() < 1e-12                          # deflation.py:40
    E = u._inner_dev(pr._Ud, pr._AUd).cpu().numpy()
    assert np.abs(sol.E - E).max() < 1e-12 * np.abs(E).max()            # test_deflation.py:52-58
    nH = sol.H.shape[1]
    Vd = sol.arnoldi._Vd[:nH]
    AV = ls.MlAMr._apply_dev(Vd)
    C = u._inner_dev(pr._Ud, AV).cpu().numpy()
    assert sol.C.shape == (d, nH)
    assert np.abs(sol.C - C).max() < 1e-10 * np.abs(C).max()            # test_deflation.py:60-63
    # projected Krylov vectors are orthogonal to U (P = I - AU <U,AU>^-1 <U,.>)
    assert np.abs(u._inner_dev(pr._Ud, sol.arnoldi._Vd[:nH + 1]).cpu().numpy()).max() < 1e-9
    x = sol.xk.reshape(-1)
    ex = np.linalg.norm(b.reshape(-1) - A @ x) / np.linalg.norm(b)
    assert abs(ex - sol.resnorms[-1]) <= 1e-9 * max(ex, 1e-30)


def test_c5_like_minres_ipB_fp32_vs_fp64():
    """config 5 shape at N = 4M: fp32 storage against the fp64 device run and the oracle's start"""
    import krypy_b200 as kp
    from krypy_b200 import problems
    from oracle import krylov_oracle as ko
    n = 2000
    A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32)
    b = problems.rhs_normal(n * n, dtype=np.float32)
    hist = {}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        ls = kp.linsys.LinearSystem(A, b, ip_B=B, self_adjoint=True, dtype=dt)
        sol = _solve(lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=40))
        hist[name] = np.array(sol.resnorms)
        assert sol.xk.dtype == dt
    assert hist["f32"].shape == hist["f64"].shape == (41,)
    np.testing.assert_allclose(hist["f32"], hist["f64"], rtol=1e-4)
    try:
        ref = ko.minres(ko.System(A.astype(np.float64), b.astype(np.float64), B=B.astype(np.float64)), tol=1e-5, maxiter=6)
    except ko.OracleConvergenceError as e:
        ref = e.result
    np.testing.assert_allclose(hist["f64"][:6], np.array(ref.resnorms)[:6], rtol=1e-10)


def test_spmv_linearity_and_checksum_at_scale():
    import torch
    from krypy_b200 import problems, _device
    ctx = _device.Context.get()
    A = problems.poisson3d(160)                                  # 4.1M rows
    N = A.shape[0]
    Ad = ctx.upload_csr(A, torch.float64)
    rng = np.random.default_rng(4)
    x, y = rng.standard_normal(N), rng.standard_normal(N)
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    ax, ay, az = (torch.empty(N, dtype=torch.float64, device="cuda") for _ in range(3))
    ctx.spmv(Ad, xd, ax)
    ctx.spmv(Ad, yd, ay)
    z = torch.from_numpy(2.0 * x - 0.5 * y).cuda()
    ctx.spmv(Ad, z, az)
    lin = (2.0 * ax - 0.5 * ay - az).abs().max().item()
    assert lin < 1e-12 * az.abs().max().item()
    colsum = np.asarray(A.sum(axis=0)).reshape(-1)              # 1^T A
    assert abs(ax.sum().item() - colsum @ x) < 1e-9 * np.abs(colsum) @ np.abs(x)
    assert np.array_equal(ax.cpu().numpy(), A @ x)               # and bit-exact against scipy
