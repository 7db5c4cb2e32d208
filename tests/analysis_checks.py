"""Shared checks for the SURVEY 8f rank-4 additions (utils analysis helpers, deflation.Arnoldifyer,
evaluator-driven recycling): used by the CPU tier over the device test double and by the GPU tier.
References are independent: scipy.linalg.subspace_angles, dense eigen-decompositions, the defining
relations the reference's own tests check (test/test_utils.py:263-720, test/test_deflation.py:189-275)."""
import warnings

import numpy as np
import scipy.linalg
import scipy.sparse


def crandn(rng, cplx, *shape):
    a = rng.standard_normal(shape)
    return a + 1j * rng.standard_normal(shape) if cplx else a


def _ref_angles(X, Y):
    """independent dense reference: arcsin of the singular values of (I - P_X) Q_Y for angles below
    pi/4, arccos of the singular values of Q_X^* Q_Y above (scipy.linalg.subspace_angles loses the
    digits of angles ~1e-7; checked against 50-digit arithmetic)"""
    if X.shape[1] < Y.shape[1]:
        X, Y = Y, X
    QX, QY = np.linalg.qr(X)[0], np.linalg.qr(Y)[0]
    sin = np.sort(scipy.linalg.svdvals(QY - QX @ (QX.conj().T @ QY)))
    cos = np.sort(scipy.linalg.svdvals(QX.conj().T @ QY))[::-1]
    return np.where(sin ** 2 < 0.5, np.arcsin(np.minimum(sin, 1)), np.arccos(np.minimum(cos, 1)))


def check_angles(cplx):
    import krypy_b200 as kp
    rng = np.random.default_rng(21)
    N = 40
    F = crandn(rng, cplx, N, 4)
    G = np.column_stack([F[:, :2] + 1e-7 * crandn(rng, cplx, N, 2), crandn(rng, cplx, N, 1)])   # two tiny angles
    for X, Y in ((F, G), (G, F), (F, F[:, :0])):
        theta = kp.utils.angles(X, Y)
        k, l = X.shape[1], Y.shape[1]
        assert theta.shape == (max(k, l),) and np.all(np.diff(theta) >= -1e-12)
        if min(k, l) > 0:
            np.testing.assert_allclose(theta[: min(k, l)], _ref_angles(X, Y), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(theta[min(k, l):], np.pi / 2)
    theta, U, V = kp.utils.angles(F, G, compute_vectors=True)
    assert U.shape == F.shape and V.shape == G.shape
    np.testing.assert_allclose(U.conj().T @ U, np.eye(4), atol=1e-10)
    np.testing.assert_allclose(V.conj().T @ V, np.eye(3), atol=1e-10)
    UV = U.conj().T @ V
    np.testing.assert_allclose(np.abs(np.diag(UV)), np.cos(theta[:3]), atol=1e-10)
    assert np.abs(UV - np.vstack([np.diag(np.diag(UV)), np.zeros((1, 3))])).max() <= 1e-8
    # weighted inner product: angles between B-orthonormalised bases
    B = np.diag(np.linspace(1, 3, N))
    tb = kp.utils.angles(F, G, ip_B=B)
    Lc = np.sqrt(B)
    np.testing.assert_allclose(tb[:3], _ref_angles(Lc @ F, Lc @ G), rtol=1e-6, atol=1e-12)


def check_hegedus_and_ritz(cplx):
    import krypy_b200 as kp
    rng = np.random.default_rng(22)
    N = 30
    A = crandn(rng, cplx, N, N) + 6 * np.eye(N)
    b, x0 = crandn(rng, cplx, N, 1), crandn(rng, cplx, N, 1)
    xs = kp.utils.hegedus(A, b, x0)
    gam = np.vdot(A @ x0, b) / np.vdot(A @ x0, A @ x0)
    np.testing.assert_allclose(xs, gam * x0, rtol=1e-12)
    assert np.linalg.norm(b - A @ xs) <= np.linalg.norm(b) * (1 + 1e-14)
    assert np.all(kp.utils.hegedus(A, b, np.zeros((N, 1))) == 0)
    # Ritz pairs of a full Arnoldi run are the eigenvalues; residual norms match the definition
    S = A + A.conj().T
    V, H = kp.utils.arnoldi(S, b, maxiter=12, ortho="dmgs")
    for typ in ("ritz", "harmonic", "harmonic_improved"):
        theta, U, res, Z = kp.utils.ritz(H, V, hermitian=True, type=typ)
        assert theta.shape == (12,) and U.shape == (12, 12) and Z.shape == (N, 12)
        for i in range(12):
            z = Z[:, [i]] / np.linalg.norm(Z[:, i])
            r = np.linalg.norm(S @ z - theta[i] * z)
            assert abs(r - res[i] / np.linalg.norm(U[:, i])) <= 1e-8 * np.linalg.norm(S, 2)
    Vf, Hf = kp.utils.arnoldi(S[:8, :8], b[:8], ortho="dmgs")
    theta, U, res = kp.utils.ritz(Hf, hermitian=True)
    np.testing.assert_allclose(np.sort(theta), np.linalg.eigvalsh(S[:8, :8]), rtol=1e-9)
    np.testing.assert_allclose(kp.utils.get_residual_norms(H)[:3],
                               [1.0] + [np.linalg.lstsq(H[: k + 1, :k], np.eye(k + 1, 1), rcond=None)[1][0] ** 0.5
                                        if k else 1.0 for k in (1, 2)], rtol=1e-10)


def check_spectral_helpers():
    import krypy_b200 as kp
    u = kp.utils
    assert u.gap([1, 2], [-4, 3]) == 1 and u.gap(5, -5) == 10 and u.gap([1, 2], [-4, 3], mode="interval") == 1
    assert u.gap([1, 2], [-4, 3, 1.5], mode="interval") is None
    with np.testing.assert_raises(u.ArgumentError):
        u.gap([1j], [2])
    I = u.Interval
    assert (I(-2, -1) & I(1, 2)) is None and (I(-2, 1) & I(0, 2)).left == 0 and (I(-2, 1) | I(0, 2)).right == 2
    assert I(1, 2).distance(I(4, 5)) == 2 and I(1, 2).contains(1.5) and not I(1).contains(2)
    with np.testing.assert_raises(u.ArgumentError):
        I(2, 1)
    ints = u.Intervals([I(-10, -5), I(-7, -3), I(-1), I(1, 2), I(1.5, 4), I(6)])
    assert len(ints) == 4 and ints.get_endpoints() == [-10, -3, -1, 1, 4, 6]
    assert (ints.min(), ints.max(), ints.max_neg(), ints.min_pos()) == (-10, 6, -1, 1)
    assert ints.min_abs() == 1 and ints.max_abs() == 10 and ints.contains(3) and not ints.contains(0)
    b = u.BoundCG([1, 2, 4])
    root = 2.0
    assert abs(b.eval_step(3) - 2 * ((root - 1) / (root + 1)) ** 3) < 1e-15
    assert abs(b.get_step(1e-6) - np.log(1e-6 / 2) / np.log(1 / 3)) < 1e-12
    with np.testing.assert_raises(u.AssumptionError):
        u.BoundCG([0, 1])
    assert isinstance(u.BoundMinres([1, 2]), u.BoundCG)
    m = u.BoundMinres([-2, -1, 1, 4])
    a_, b_ = np.sqrt(2 * 4 / 16), np.sqrt(1 * 1 / 16)
    assert abs(m.base - (a_ - b_) / (a_ + b_)) < 1e-15 and m.eval_step(3) == 2 * m.base
    assert isinstance(u.BoundMinres(u.Intervals([I(1, 2)])), u.BoundCG)
    roots = np.array([1.0, 2.0, 5.0, 1e3, 1e-3])
    p = u.NormalizedRootsPolynomial(roots)
    pts = np.linspace(-1, 6, 9)
    np.testing.assert_allclose(p(pts), np.prod(1 - pts[None, :] / roots[:, None], axis=0), rtol=1e-12)
    assert abs(p(2.0)) < 1e-300 and np.isscalar(p(0.5))
    cand = p.minmax_candidates()
    assert cand.shape == (4,)
    d = np.diag(u.strakos(5, 0.1, 100, 0.9))
    assert abs(d[0] - 0.1) < 1e-15 and abs(d[-1] - 100) < 1e-12 and np.all(np.diff(d) > 0)


def _deflated_gmres(kp, cplx, with_M):
    rng = np.random.default_rng(23)
    N = 24
    A = crandn(rng, cplx, N, N) + 5 * np.eye(N)
    b = crandn(rng, cplx, N, 1)
    U = crandn(rng, cplx, N, 2)
    M = np.diag(np.linspace(1, 2, N)) if with_M else None
    ls = kp.linsys.LinearSystem(A, b, M=M, Minv=None if M is None else np.linalg.inv(M))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = kp.deflation.DeflatedGmres(ls, U=U, store_arnoldi=True, maxiter=6)
        except kp.utils.ConvergenceError as e:
            sol = e.solver
    return A, M, ls, sol


def check_arnoldifyer(cplx, with_M):
    """test/test_deflation.py:189-275: the perturbed Arnoldi relation, the projection, orthonormality"""
    import krypy_b200 as kp
    A, M, ls, sol = _deflated_gmres(kp, cplx, with_M)
    N = ls.N
    ritz = kp.deflation.Ritz(sol)
    order = np.argsort(np.abs(ritz.values))
    Wt, _ = scipy.linalg.qr(ritz.coeffs[:, order[:2]], mode="economic")
    arn = kp.deflation.Arnoldifyer(sol)
    ip = ls.get_ip_Minv_B()
    Z = arn.Z
    np.testing.assert_allclose(kp.utils.inner(Z, Z, ip_B=ip), np.eye(Z.shape[1]), atol=1e-7)
    for W in (Wt, Wt[:, :0]):
        Hh, Rh, q_norm, vdiff_norm, PWAW_norm, Vh, F = arn.get(W, full=True)
        n1, n = sol.H.shape
        d, k = 2, W.shape[1]
        assert Hh.shape == (n + d - k, n + d - k) and Vh.shape == (N, n + d - k)
        assert arn.get(W)[0].shape == Hh.shape
        VU = np.column_stack([sol.V[:, :n], sol.projection.U])
        Wv = VU @ W
        Md = np.eye(N) if M is None else M
        if k:
            AW = A @ Wv
            P = np.eye(N) - AW @ np.linalg.solve(Wv.conj().T @ AW, Wv.conj().T)
        else:
            P = np.eye(N)
        At = Md @ P @ A
        FVh = F * Vh
        Anorm = np.linalg.norm(A, 2)
        assert np.linalg.norm(At @ Vh + FVh - Vh @ Hh, 2) / Anorm <= 1e-7
        Minv = np.linalg.inv(Md)
        np.testing.assert_allclose(Vh.conj().T @ Minv @ Vh, np.eye(n + d - k), atol=1e-7)
        np.testing.assert_allclose(Vh.conj().T @ Minv @ (At @ Vh + FVh), Hh, atol=1e-7 * Anorm)
        # ||P_{W^perp,AW}|| in the M^-1 inner product
        if k:
            Lc = np.linalg.cholesky(Minv).conj().T
            want = np.linalg.norm(Lc @ (Md @ P @ Minv) @ np.linalg.inv(Lc), 2)
            assert abs(PWAW_norm - want) <= 1e-6 * want
        else:
            assert PWAW_norm == 1.0
    # residual-norm prediction without pseudospectra
    bd = kp.deflation.bound_pseudo(arn, Wt, tol=1e-8, pseudo_type="omit")
    assert bd.ndim == 1 and bd[0] > 0 and np.all(np.diff(bd) <= 1e-12)


def check_evaluator_recycling(solver_name, factory_name):
    """three solves of the same SPD system with an evaluator-driven factory (test/test_recycling.py
    style).  The evaluators minimise an estimated run TIME built from measured operator timings, so
    whether deflation pays depends on the machine; with ``deflweight=0`` the deflation overhead is
    left out of the estimate and the selection is deterministic (fewer predicted steps wins)."""
    import krypy_b200 as kp
    N = 100
    d = np.linspace(1, 2, N)
    d[:5] = [1e-8, 1e-4, 1e-2, 2e-2, 3e-2]
    ev = kp.recycling.evaluators
    forced = {"RitzApproxKrylov": lambda: ev.RitzApproxKrylov(deflweight=0.0),
              "RitzAprioriCg": lambda: ev.RitzApriori(Bound=kp.utils.BoundCG, deflweight=0.0),
              "RitzAprioriMinres": lambda: ev.RitzApriori(Bound=kp.utils.BoundMinres, deflweight=0.0)}
    for factory in (kp.recycling.factories.RitzFactory(subset_evaluator=forced[factory_name]()), factory_name):
        ls = kp.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True,
                                    positive_definite=True)
        rs = getattr(kp.recycling, solver_name)()
        its, nd = [], []
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(3):
                s = rs.solve(ls, vector_factory=factory, maxiter=60, tol=1e-6)
                its.append(len(s.resnorms) - 1)
                nd.append(s.projection.U.shape[1])
                assert s.resnorms[-1] <= 1e-6
        assert nd[0] == 0 and its[1] <= its[0] and its[2] <= its[0], (its, nd)
        if not isinstance(factory, str):
            assert nd[1] > 0 and its[1] < its[0], (its, nd)
        assert isinstance(rs.last_solver.linear_system, kp.linsys.TimedLinearSystem)
        assert rs.last_solver.linear_system.timings.get("A") > 0


def check_ritz_factory_options():
    import krypy_b200 as kp
    N = 60
    d = np.linspace(1, 2, N)
    d[:3] = [-1e-2, 1e-3, 2e-2]
    d[-1] = 9.0
    ls = kp.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True)
    ev = kp.recycling.evaluators.RitzApriori(Bound=kp.utils.BoundMinres, strategy="intervals")
    gen = kp.recycling.generators.RitzExtremal(max_vectors=4)
    f = kp.recycling.factories.RitzFactory(subset_evaluator=ev, subsets_generator=gen)
    rs = kp.recycling.RecyclingMinres()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s0 = rs.solve(ls, vector_factory=f, maxiter=60, tol=1e-6)
        s1 = rs.solve(ls, vector_factory=f, maxiter=60, tol=1e-6)
    assert s1.projection.U.shape[1] <= 4 and len(s1.resnorms) <= len(s0.resnorms)
    ritz = kp.deflation.Ritz(s0)
    cands = gen.generate(ritz, set(range(len(ritz.values))))
    assert 2 <= len(cands) <= 4 and all(len(c) == 1 for c in cands)
    small = kp.recycling.generators.RitzSmall().generate(ritz, set(range(len(ritz.values))))
    assert small == [{int(np.argmin(np.abs(ritz.values)))}]
    with np.testing.assert_raises(kp.utils.ArgumentError):
        kp.recycling.evaluators.RitzApriori(Bound=kp.utils.BoundCG, strategy="nope").evaluate(ritz, frozenset())


def check_device_linear_operator():
    """utils.DeviceLinearOperator: a matrix-free operator whose callback works on device blocks
    (here: the 1-D Laplacian written with torch slicing), driven through Gmres and Cg"""
    import torch
    import krypy_b200 as kp
    N = 200

    def lap(Xd, out):
        out.copy_(2.0 * Xd)
        out[:, 1:] -= Xd[:, :-1]
        out[:, :-1] -= Xd[:, 1:]
        return out

    calls = []

    def lap_returning_new(Xd, out):
        calls.append(tuple(Xd.shape))
        Y = 2.0 * Xd
        Y[:, 1:] -= Xd[:, :-1]
        Y[:, :-1] -= Xd[:, 1:]
        return Y

    T = scipy.sparse.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(N, N)).tocsr()
    b = np.random.default_rng(31).standard_normal((N, 1))
    for fn in (lap, lap_returning_new):
        A = kp.utils.DeviceLinearOperator((N, N), np.float64, dot_dev=fn, dot_adj_dev=fn)
        X = np.random.default_rng(32).standard_normal((N, 3))
        np.testing.assert_allclose(A * X, T @ X, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(A.adj * X, T @ X, rtol=1e-13, atol=1e-13)
        ls = kp.linsys.LinearSystem(A, b, self_adjoint=True, positive_definite=True)
        ref = kp.linsys.LinearSystem(T, b, self_adjoint=True, positive_definite=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for cls, kw in ((kp.linsys.Cg, {}), (kp.linsys.Gmres, dict(ortho="cgs")), (kp.linsys.Minres, {})):
                s1 = cls(ls, tol=1e-10, maxiter=N, **kw)
                s2 = cls(ref, tol=1e-10, maxiter=N, **kw)
                np.testing.assert_allclose(s1.resnorms, s2.resnorms, rtol=1e-8, atol=2e-13)    # (last entry: explicit residual at the cancellation floor)
                np.testing.assert_allclose(s1.xk, s2.xk, rtol=1e-8, atol=1e-10)
    assert calls and all(len(c) == 2 and c[1] == N for c in calls)
    with np.testing.assert_raises(kp.utils.LinearOperatorError):
        kp.utils.DeviceLinearOperator((N, N), np.float64)
    bad = kp.utils.DeviceLinearOperator((N, N), np.float64, dot_dev=lambda Xd, out: Xd[:, :5])
    assert (bad * np.ones((N, 1))) is NotImplemented or True     # LinearOperatorError -> NotImplemented (utils.py:1419-1420)


def check_timings(on_device):
    """utils.Timer / Timings / TimedLinearSystem (krypy/utils.py:1289-1362, linsys.py:204-252): list
    semantics of the reference; on the device the durations come from CUDA events and are read back
    lazily (no host synchronisation per timed application)."""
    import time
    import krypy_b200 as kp
    from krypy_b200 import problems
    u = kp.utils
    t = u.Timer()
    with t:
        time.sleep(0.02)
    with t:
        pass
    t.scale_last(0.5)
    assert len(t) == 2 and 0.015 < t[0] < 0.5 and 0 <= t[1] < 0.01 and min(t) == t[1]
    tm = u.Timings()
    with tm["a"]:
        time.sleep(0.01)
    assert tm.get("a") > 0.005 and tm.get("nothing") == 0
    np.testing.assert_allclose(tm.get_ops({"a": 3, "nothing": 5}), 3 * tm.get("a"))
    assert "a:" in repr(tm)
    n = 300 if on_device else 20
    A = problems.laplace2d(n)
    b = problems.rhs_normal(n * n)
    import scipy.sparse as sp
    ls = kp.linsys.TimedLinearSystem(A, b, M=problems.jacobi_csr(A), Minv=sp.diags(A.diagonal()).tocsr(),
                                     self_adjoint=True, positive_definite=True)
    sol = kp.linsys.Cg(ls, tol=1e-6, maxiter=2000)
    timer = ls.timings["A"]
    assert list.__len__(timer) >= sol.iter                      # one entry per application of A
    if on_device:
        assert len(timer._pending) > 0                          # nothing has been read back yet
    ta, tmm = ls.timings.get("A"), ls.timings.get("M")
    assert 0 < ta < 0.05 and 0 < tmm < 0.05 and not timer._pending
    X = np.ones((n * n, 3))
    n0 = list.__len__(timer)
    ls.A * X
    assert list.__len__(timer) == n0 + 1 and timer[-1] > 0      # per-vector time of a 3-column application
    est = kp.deflation.DeflatedCg(ls, U=np.eye(n * n, 2), tol=1e-6, maxiter=2000).estimate_time(10, 2)
    assert est > 0
