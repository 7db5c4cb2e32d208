"""CPU tier: the product's HOST logic (krypy_b200/{utils,linsys,deflation,_convenience}.py) driven
end to end over a numpy test double of the device layer (tests/fake_device.py) and compared with the
reference fixtures.  This validates the control flow, bookkeeping and attribute semantics of the
Python layer without a GPU; the kernels themselves are covered by the -m gpu tests."""
import warnings

import numpy as np
import pytest

import cases
import fake_device
import runners


@pytest.fixture()
def fake(monkeypatch):
    return fake_device.install(monkeypatch)


def _check_history(got, ref, rtol=1e-10, atol=5e-13):   # explicit entries of the kappa=1e5 case: cancellation noise
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + atol
    assert np.all(err <= bound), (int(np.argmax(err / bound)), float((err / bound).max()))


@pytest.mark.parametrize("name", cases.ALL_CASES + cases.COMPLEX_CASES)
def test_host_logic_reproduces_reference_fixtures(fake, name):
    gold = runners.load_golden(name)
    got = runners.run_product(name)
    assert bool(got["converged"]) == bool(gold["converged"])
    rtol = 1e-5 if name == "shifted_minres_ipB" else 1e-10      # fp32 INPUTS: see test_solvers_gpu.py
    _check_history(got["resnorms"], gold["resnorms"], rtol=rtol)
    scale = np.abs(gold["xk"]).max() + 1e-300
    assert np.abs(got["xk"] - gold["xk"]).max() <= max(1e-8, rtol) * scale
    for k in ("iter", "V_shape"):
        if k in gold:
            assert np.array_equal(got[k], gold[k]), k
    rn = gold["resnorms"]
    ngood = int(np.argmax(rn < 1e-7)) if np.any(rn < 1e-7) else len(rn)
    for k in ("H", "C", "V_colsum_abs"):
        if k in gold:
            assert got[k].shape == gold[k].shape, k
            nc = min(max(ngood - 1, 0), gold[k].shape[-1])
            a, b = got[k][..., :nc], gold[k][..., :nc]
            if a.size:
                assert np.abs(a - b).max() <= 1e-6 * (np.abs(b).max() + 1e-300), k
    for k in ("E", "UMlr", "rhos"):
        if k in gold:
            np.testing.assert_allclose(got[k], gold[k], rtol=1e-8, atol=1e-13 * (np.abs(gold[k]).max() + 1e-300))
    # the host layer went through the device entry points, not around them
    assert fake.launch_count() > 0


def test_host_logic_block_variants_and_lookahead(fake):
    gold = runners.load_golden("lap2d_gmres30")
    for ortho in ("cgs", "cgs2", "dmgs"):
        got = runners.run_product("lap2d_gmres30", ortho=ortho)
        _check_history(got["resnorms"], gold["resnorms"])
    # explicit_residual=True disables the look-ahead: same history
    got = runners.run_product("lap2d_gmres30", explicit_residual=True)
    assert got["resnorms"].shape == gold["resnorms"].shape
    np.testing.assert_allclose(got["resnorms"], gold["resnorms"], rtol=1e-6)


def test_host_logic_known_answers_and_semantics(fake):
    """reference test/test_convenience_wrappers.py numbers + SURVEY 3.6 traps on the CPU tier"""
    import krypy_b200 as kp
    A = np.diag([1.0e-3] + list(range(2, 101))).astype(float)
    b = np.ones(100)
    ref = {kp.cg: [1004.1873775173957, 1000.0003174916551, 999.9999999997555],
           kp.gmres: [1004.1873724888546, 1000.0003124630923, 999.999994971191],
           kp.minres: [1004.187372488912, 1000.0003124632159, 999.9999949713145]}
    for fn, r in ref.items():
        sol, _ = fn(A, b, inner_product=np.dot)
        assert sol.shape == b.shape
        assert abs(np.sum(np.abs(sol)) - r[0]) < 1e-11 * r[0]
        assert abs(np.sqrt(sol @ sol) - r[1]) < 1e-11 * r[1]
        assert abs(np.max(np.abs(sol)) - r[2]) < 1e-11 * r[2]
    sol, _ = kp.cg(A, b, inner_product=lambda x, y: np.dot(x, y))     # callable ip_B path
    assert abs(np.sum(np.abs(sol)) - 1004.1873775173957) < 1e-11 * 1004.1873775173957
    ls = kp.linsys.LinearSystem(A, b, self_adjoint=True, positive_definite=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for cls, it in ((kp.linsys.Cg, 55), (kp.linsys.Minres, 54), (kp.linsys.Gmres, 54)):
            s = cls(ls, store_arnoldi=True)
            assert s.iter == it and s.V.shape == (100, 56) and s.H.shape == (56, 55)
        for cls, vs, hs in ((kp.linsys.Cg, (100, 10), (10, 9)), (kp.linsys.Minres, (100, 11), (11, 10)),
                            (kp.linsys.Gmres, (100, 11), (11, 10))):
            with pytest.raises(kp.utils.ConvergenceError) as ei:
                cls(ls, maxiter=10, store_arnoldi=True)
            s = ei.value.solver
            assert len(s.resnorms) - 1 == 10 and s.iter == 9 and s.V.shape == vs and s.H.shape == hs
    z = kp.linsys.Gmres(kp.linsys.LinearSystem(A, np.zeros((100, 1))))
    assert z.resnorms == [0.0] and np.all(z.xk == 0)
    d = kp.deflation.DeflatedGmres(kp.linsys.LinearSystem(A, b), U=np.eye(100, 2), store_arnoldi=True)
    n = d.H.shape[1]
    assert d.C.shape == (2, n) and d.E.shape == (2, 2) and d.B_.shape == (n + 1, 2)
    # restart loop hands the residual over and reuses the workspace
    r = kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=10)
    assert r.resnorms[-1] <= 1e-5
    with pytest.raises(kp.utils.ConvergenceError):
        kp.linsys.RestartedGmres(ls, maxiter=5, max_restarts=1)
    zs = kp.linsys.Gmres(kp.linsys.LinearSystem(A.astype(complex), (1 + 1j) * b), tol=1e-8)   # complex: real embedding
    assert zs.xk.dtype == np.complex128 and zs.resnorms[-1] <= 1e-8
    assert np.allclose(zs.xk.reshape(-1), (1 + 1j) * np.linalg.solve(A, b).reshape(-1), rtol=1e-6)
    hs = kp.linsys.Gmres(ls, ortho="house", maxiter=60)                 # Householder Arnoldi (utils.py:970-994)
    ms = kp.linsys.Gmres(ls, maxiter=60)
    assert len(hs.resnorms) == len(ms.resnorms) and np.allclose(hs.resnorms, ms.resnorms, rtol=1e-6)
    with pytest.raises(kp.utils.ArgumentError):
        kp.linsys.Gmres(kp.linsys.LinearSystem(A, b, ip_B=np.eye(100) * 2.0), ortho="house")
    with pytest.raises(kp.utils.ArgumentError):
        kp.linsys.Gmres(ls, ortho="nope")


@pytest.mark.parametrize("name", ["convdiff_defl_gmres", "lap2d_defl_minres", "lap2d_defl_cg", "c1_gmres_defl"])
def test_ritz_pairs_host_logic(fake, name):
    import ritz_checks
    ritz_checks.check_ritz_pairs(name)


@pytest.mark.parametrize("sname", ["cg", "minres", "gmres"])
@pytest.mark.parametrize("which", ["lm", "sm", "lr", "sr", "li", "si", "smallest_res"])
def test_recycling_host_logic(fake, sname, which):
    import ritz_checks
    ritz_checks.check_recycling(sname, which)


def test_gmres_step_limit_is_reported_up_front(fake):
    """the single-CTA Givens recurrence caps a cycle at 2000 steps (complex: 1000): a larger maxiter
    (the default is N) is refused before any work, with a pointer to RestartedGmres"""
    import krypy_b200 as kp
    import scipy.sparse as sp
    N = 2500
    ls = kp.linsys.LinearSystem(sp.identity(N, format="csr") * 2.0, np.ones(N))
    with pytest.raises(kp.utils.ArgumentError, match="RestartedGmres"):
        kp.linsys.Gmres(ls)                       # maxiter defaults to N = 2500 > 2000
    with pytest.raises(kp.utils.ArgumentError):
        kp.linsys.Gmres(kp.linsys.LinearSystem(sp.identity(N, format="csr") * (2.0 + 1j), np.ones(N)), maxiter=1001)
    assert kp.linsys.Gmres(ls, maxiter=2000).resnorms[-1] <= 1e-5


def test_no_reference_cycles_keep_device_memory_alive(fake):
    """operators and finished solvers are freed by reference counting alone: a cycle would keep the CSR
    matrix and the bases in HBM until Python's cyclic collector happens to run (found on the B200 as
    sporadic 100 ms steps of the end-to-end bench)"""
    import gc
    import weakref
    import krypy_b200 as kp
    from krypy_b200 import problems
    A, b = problems.laplace2d(12), problems.rhs_normal(144)
    gc.collect()
    gc.disable()
    try:
        refs = []
        for make in (lambda ls: kp.linsys.Gmres(ls, maxiter=100, tol=1e-8, store_arnoldi=True),
                     lambda ls: kp.linsys.Cg(ls, tol=1e-8, store_arnoldi=True),
                     lambda ls: kp.linsys.Minres(ls, tol=1e-8),
                     lambda ls: kp.deflation.DeflatedGmres(ls, U=np.eye(144, 2), tol=1e-8, store_arnoldi=True),
                     lambda ls: kp.deflation.DeflatedCg(ls, U=np.eye(144, 2), tol=1e-8, store_arnoldi=True)):
            ls = kp.linsys.LinearSystem(A, b, self_adjoint=True, positive_definite=True)
            sol = make(ls)
            getattr(sol, "V", None)
            refs += [weakref.ref(sol), weakref.ref(ls), weakref.ref(ls.A)]
            del sol, ls
        assert all(r() is None for r in refs)
    finally:
        gc.enable()


def check_cholqr2_against_mgs(kp, N=300, k=7, seed=4):
    """utils.qr by CholQR2 (block kernels) against the column-by-column MGS path: same factorisation up to
    round-off; rank-deficient and ill-conditioned blocks fall back to MGS (utils.py:705-706 rule)"""
    u = kp.utils
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, k)) @ np.triu(rng.standard_normal((k, k)) + 3 * np.eye(k))
    old = u._BLOCK_MIN_N
    try:
        u._BLOCK_MIN_N = 0
        Q, R = u.qr(X)
        u._BLOCK_MIN_N = 10 ** 9
        Q0, R0 = u.qr(X)
        np.testing.assert_allclose(Q.T @ Q, np.eye(k), atol=5e-15 * k)
        np.testing.assert_allclose(Q @ R, X, rtol=0, atol=1e-13 * np.abs(X).max() * k)
        assert np.allclose(np.tril(R, -1), 0) and np.all(np.diag(R) > 0)
        np.testing.assert_allclose(R, R0, rtol=0, atol=1e-12 * np.abs(R0).max())
        np.testing.assert_allclose(Q, Q0, rtol=0, atol=1e-11)
        # block inner products through the one-pass Gram kernel
        u._BLOCK_MIN_N = 0
        Y = rng.standard_normal((N, 5))
        np.testing.assert_allclose(u.inner(X, Y), X.T @ Y, rtol=0, atol=1e-13 * (np.abs(X).T @ np.abs(Y)).max())
        np.testing.assert_allclose(u.inner(X, X), X.T @ X, rtol=0, atol=1e-13 * (np.abs(X).T @ np.abs(X)).max())
        # rank deficient: the MGS fallback leaves the dependent column un-normalised like the reference
        Xd = X.copy()
        Xd[:, 3] = Xd[:, 1] * 2.0
        Qd, Rd = u.qr(Xd)
        u._BLOCK_MIN_N = 10 ** 9
        Qd0, Rd0 = u.qr(Xd)
        np.testing.assert_allclose(Rd, Rd0, rtol=0, atol=1e-10 * np.abs(Rd0).max())
    finally:
        u._BLOCK_MIN_N = old


def test_cholqr2_host_logic(fake):
    import krypy_b200 as kp
    check_cholqr2_against_mgs(kp)
    assert fake.calls.get("gram", 0) >= 4 and fake.calls.get("block_trsm", 0) >= 2


# ---------------------------------------------------------------- restart cycles enqueued ahead of the host
def _restarted_vs_oracle(kp, A, b, m, restarts, tol, M=None):
    from oracle import krylov_oracle as ko
    ls = kp.linsys.LinearSystem(A, b, M=M)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = kp.linsys.RestartedGmres(ls, maxiter=m, max_restarts=restarts, tol=tol, ortho="cgs")
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        try:
            ref = ko.restarted_gmres(ko.System(A, b, M=M), maxiter=m, max_restarts=restarts, tol=tol)
        except ko.OracleConvergenceError as e:
            ref = e.result
    return sol, ref


@pytest.mark.parametrize("withM", [False, True])
@pytest.mark.parametrize("tol,restarts", [(1e-14, 4), (1e-6, 60)])
def test_cycle_ahead_mode_over_the_double(fake, monkeypatch, tol, restarts, withM):
    """linsys.Gmres from the third restart cycle on: all steps of a cycle enqueued at once, records booked in
    bulk, the cycle's end enqueued before the bookkeeping, the next cycle launched speculatively on the
    device-side norm (a real-device feature, switched on for the double here).  Same histories, solution, H
    and R as the step-by-step pace and as the oracle; the 1e-6 case converges in the middle of such a cycle."""
    import krypy_b200 as kp
    from krypy_b200 import problems
    import scipy.sparse as sp
    n = 24
    A = problems.laplace2d(n)
    rng = np.random.default_rng(3)
    dM = sp.diags(1.0 / (4.0 + rng.random(n * n))).tocsr() if withM else None
    b = rng.standard_normal((n * n, 1))
    fake.cycle_ahead = False
    base, ref = _restarted_vs_oracle(kp, A, b, 8, restarts, tol, dM)
    calls_base = dict(fake.calls)
    fake.reset_launch_count()
    fake.cycle_ahead = True
    sol, _ = _restarted_vs_oracle(kp, A, b, 8, restarts, tol, dM)
    assert fake.calls.get("tri_solve_t", 0) > 0 and calls_base.get("tri_solve_t", 0) == 0      # the mode was on
    hits = getattr(sol._workspace, "prelaunch_hits", 0)
    ncycles = (len(sol.resnorms) - 1 + 7) // 8
    # cycles 4.. of a solve without M start from a cycle their predecessor launched speculatively
    # (fewer when the last cycles before convergence ended within 2 tol: no speculation that close to it)
    assert (hits == 0) if withM else (max(ncycles - 7, min(ncycles - 3, 1), 0) <= hits <= max(ncycles - 3, 0)), \
        (hits, ncycles)
    assert np.array_equal(np.array(sol.resnorms), np.array(base.resnorms))                      # bit for bit
    assert np.array_equal(sol.xk, base.xk)
    last, last0 = sol._last, base._last
    assert np.array_equal(last.R, last0.R) and np.array_equal(last.arnoldi.H, last0.arnoldi.H)
    assert last.iter == last0.iter and last.arnoldi.iter == last0.arnoldi.iter
    _check_history(np.array(sol.resnorms), np.array(ref.resnorms))
    if tol > 1e-10:
        assert sol.resnorms[-1] <= tol and (len(sol.resnorms) - 1) % 8 != 0
    # speculation is announced by the restart driver only: a cycle launched for a solve that then converges
    # would be garbage nobody reads, but none is launched once the updated residual is within 2 tol
    monkeypatch.setenv("KRY_PRELAUNCH", "0")
    fake.reset_launch_count()
    sol2, _ = _restarted_vs_oracle(kp, A, b, 8, restarts, tol, dM)
    assert np.array_equal(np.array(sol2.resnorms), np.array(base.resnorms))
    assert getattr(sol2._workspace, "prelaunch_hits", 0) == 0


def test_cycle_records_bulk_booking_stops_at_the_first_step_that_needs_a_decision(fake):
    """Gmres._book_cycle_records / _no_invariance_in_sight on a synthetic mailbox: convergence in the middle,
    an invariant-looking step, NaN -- the bulk path books exactly the steps before, the loop takes the rest"""
    import krypy_b200 as kp
    m = 6
    offs = np.concatenate([[0], np.cumsum([2 * (j + 2) + 1 for j in range(m)])]).astype(np.int64)
    rng = np.random.default_rng(5)

    def mailbox(resid, sub):
        mb = np.zeros(int(offs[-1]))
        for j in range(m):
            o = offs[j]
            mb[o] = resid[j]
            mb[o + 1:o + 1 + j + 2] = rng.standard_normal(j + 2)
            mb[o + 1 + j + 1] = sub[j]                       # H[j+1, j]
            mb[o + 1 + j + 2:o + 1 + 2 * (j + 2)] = rng.standard_normal(j + 2)
        return mb

    class _Ls(object):
        MMlb_norm = 2.0

    def solver():
        s = object.__new__(kp.linsys.Gmres)
        s.linear_system = _Ls()
        s.tol = 1e-3
        s.resnorms = [1.0]
        s.R = np.zeros((m + 1, m))
        s._ws = kp.utils.SolverWorkspace(graphs="off")
        ar = type("A", (), {})()
        ar.H = np.zeros((m + 1, m))
        ar._hfro2 = 0.0
        ar.iter = 0
        return s, ar

    # (1) nothing special: all but the last step are booked, the fills are deferred
    s, ar = solver()
    mb = mailbox([1.0, 0.8, 0.6, 0.4, 0.2, 0.1], [1.0] * m)
    assert s._no_invariance_in_sight(ar, mb, offs, m)
    fill = s._book_cycle_records(ar, mb, offs, m)
    assert fill is not None and ar.iter == m - 1 and s.iter == m - 2
    assert np.all(s.R == 0.0) and len(s.resnorms) == 2
    fill()
    assert s.resnorms == [1.0] + [v / 2.0 for v in (1.0, 0.8, 0.6, 0.4, 0.2)]
    for j in range(m - 1):
        assert np.array_equal(ar.H[: j + 2, j], mb[offs[j] + 1:offs[j] + 1 + j + 2])
        assert np.array_equal(s.R[: j + 2, j], mb[offs[j] + 1 + j + 2:offs[j] + 1 + 2 * (j + 2)])
    assert np.all(ar.H[:, m - 1] == 0.0) and np.all(s.R[:, m - 1] == 0.0)
    assert np.isclose(ar._hfro2, sum(float(np.sum(ar.H[:, j] ** 2)) for j in range(m - 1)))
    # (2) the updated residual meets the tolerance at step 3: steps 0..2 in bulk, at once
    s, ar = solver()
    mb = mailbox([1.0, 0.8, 0.6, 1e-3, 0.2, 0.1], [1.0] * m)
    assert s._book_cycle_records(ar, mb, offs, m) is None and ar.iter == 3 and len(s.resnorms) == 4
    assert np.any(s.R[:, 2] != 0.0) and np.all(s.R[:, 3] == 0.0)
    # (3) an invariant-looking subspace at step 2 (tiny H[3, 2]) and (4) NaN: the bulk path stops before
    for bad in (1e-16, np.nan):
        s, ar = solver()
        sub = [1.0] * m
        sub[2] = bad
        mb = mailbox([1.0, 0.8, 0.6, 0.4, 0.2, 0.1], sub)
        assert not s._no_invariance_in_sight(ar, mb, offs, m)
        assert s._book_cycle_records(ar, mb, offs, m) is None and ar.iter == 2 and len(s.resnorms) == 3
    # (5) convergence at step 0: nothing is booked
    s, ar = solver()
    mb = mailbox([1e-4, 0.8, 0.6, 0.4, 0.2, 0.1], [1.0] * m)
    assert s._book_cycle_records(ar, mb, offs, m) is None and ar.iter == 0 and s.resnorms == [1.0]
