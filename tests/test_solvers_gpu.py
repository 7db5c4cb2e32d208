"""Solver-level parity on the GPU: the CUDA path (through the KryPy-compatible
classes, which call the C ABI) against (i) the oracle on the same seeded inputs,
(ii) the reference fixtures in tests/golden/ and (iii) the literal known answers
of the reference's own test-suite.

Tolerances (fp64): iteration-for-iteration relative residual norms within 1e-10
(BASELINE.json north_star); entries that are EXPLICIT residuals
(krypy/linsys.py:450-463: the last entry of a converged/ended solve) are subject
to cancellation in b - A x_k and get 1e-10*res + 1e-13."""
import warnings

import numpy as np
import pytest

import cases
import runners

pytestmark = pytest.mark.gpu

KNOWN = {  # reference test/test_convenience_wrappers.py:10-12 and :37-39
    "c1_cg": [1004.1873775173957, 1000.0003174916551, 999.9999999997555],
    "c1_gmres": [1004.1873724888546, 1000.0003124630923, 999.999994971191],
    "c1_minres": [1004.187372488912, 1000.0003124632159, 999.9999949713145],
    "c1_cg_defl": [1004.1873775173271, 1000.0003174918709, 1000.0],
    "c1_minres_defl": [1004.1873774950692, 1000.0003174918709, 1000.0],
    "c1_gmres_defl": [1004.1873774950692, 1000.0003174918709, 1000.0],
}

F32_INPUT_CASES = {"shifted_minres_ipB"}


def _check_history(got, ref, rtol=1e-10, atol=1e-13):
    """|d| <= 1e-10 * res_k + 1e-13: the absolute term (450 eps, in units of ||b||) is the
    attainable accuracy of a residual norm -- the reference's own histories move by that
    much under a 1e-16 relative perturbation of b once res_k approaches 1e-8 (measured
    with the oracle, e.g. case dense_gmres_MlMr)."""
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + atol
    assert np.all(err <= bound), (np.argmax(err / bound), (err / bound).max())


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_known_answers_through_convenience_api(name):
    """the reference's golden numbers, through krypy_b200.cg/minres/gmres"""
    import krypy_b200 as kp
    c = cases.case_inputs(name)
    fn = {"cg": kp.cg, "gmres": kp.gmres, "minres": kp.minres}[c["solver"]]
    kw = {}
    if "U" in c["kw"]:
        kw["U"] = c["kw"]["U"].reshape(-1)
    else:
        kw["inner_product"] = np.dot
    sol, solver = fn(c["A"], c["b"], **kw)
    assert sol.shape == c["b"].shape
    ref = KNOWN[name]
    tol = 1.0e-11
    assert abs(np.sum(np.abs(sol)) - ref[0]) < tol * ref[0]
    assert abs(np.sqrt(np.dot(sol, sol)) - ref[1]) < tol * ref[1]
    assert abs(np.max(np.abs(sol)) - ref[2]) < tol * ref[2]


@pytest.mark.parametrize("name", cases.ALL_CASES)
def test_matches_reference_fixture_and_oracle(name):
    gold = runners.load_golden(name)
    orac = runners.run_oracle(name)
    got = runners.run_product(name)    # default dtype rule = the reference's promotion (>= fp64)
    assert bool(got["converged"]) == bool(gold["converged"])
    # fp32 INPUTS: the reference forms ||b|| and v_1 = b/||b|| in fp32 arithmetic before it
    # promotes (linsys.py:120-122 on a float32 b), so its own history carries 1e-7 noise
    rtol = 1e-5 if name in F32_INPUT_CASES else 1e-10
    for ref in (gold, orac):
        _check_history(got["resnorms"], ref["resnorms"], rtol=rtol)
        scale = np.abs(ref["xk"]).max() + 1e-300
        assert np.abs(got["xk"] - ref["xk"]).max() <= max(1e-8, rtol) * scale
        for k in ("iter", "V_shape"):
            if k in ref:
                assert np.array_equal(got[k], ref[k]), k
        # Derived matrices: compare the columns built while the residual is still >= 1e-7.
        # Later Arnoldi vectors are q/||q|| with ||q|| at the cancellation floor, so H, C and
        # V there are rounding noise in the reference itself.
        rn = ref["resnorms"]
        ngood = int(np.argmax(rn < 1e-7)) if np.any(rn < 1e-7) else len(rn)
        for k in ("H", "C", "V_colsum_abs"):
            if k in ref:
                assert got[k].shape == ref[k].shape, k
                nc = min(max(ngood - 1, 0), ref[k].shape[-1])
                a, b = got[k][..., :nc], ref[k][..., :nc]
                if a.size:
                    sc = np.abs(b).max() + 1e-300
                    assert np.abs(a - b).max() <= 1e-6 * sc, k
        for k in ("E", "UMlr", "rhos"):
            if k in ref:
                assert got[k].shape == ref[k].shape, k
                np.testing.assert_allclose(got[k], ref[k], rtol=1e-8, atol=1e-13 * (np.abs(ref[k]).max() + 1e-300))


def test_fp32_storage_minres_ipB_against_fp64_reference():
    """BASELINE config 5 shape at test size: fp32 storage vs the fp64 reference, 1e-4 relative"""
    import krypy_b200 as kp
    gold = runners.load_golden("shifted_minres_ipB")
    c = cases.case_inputs("shifted_minres_ipB")
    ls = kp.linsys.LinearSystem(c["A"], c["b"], dtype=np.float32, **c["ls"])
    assert ls.dtype == np.float32
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = kp.linsys.Minres(ls, **c["kw"])
        except kp.utils.ConvergenceError as e:
            sol = e.solver
    got = np.array(sol.resnorms)
    ref = gold["resnorms"]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-4)
    assert sol.xk.dtype == np.float32


@pytest.mark.parametrize("ortho", ["cgs", "cgs2", "dmgs"])
def test_block_gram_schmidt_variants_match_reference_mgs(ortho):
    """fused block CGS / CGS2 histories agree with the reference's MGS on the sparse
    configs to the 1e-10 contract"""
    gold = runners.load_golden("lap2d_gmres30")
    got = runners.run_product("lap2d_gmres30", ortho=ortho)
    _check_history(got["resnorms"], gold["resnorms"], rtol=1e-10)
    gold = runners.load_golden("convdiff_defl_gmres")
    got = runners.run_product("convdiff_defl_gmres", ortho=ortho)
    _check_history(got["resnorms"], gold["resnorms"], rtol=1e-10)


def test_post_solve_attribute_semantics():
    """SURVEY 3.6 drop-in traps, checked against the reference's measured behaviour"""
    import krypy_b200 as kp
    A = np.diag([1e-3] + list(range(2, 101))).astype(float)
    b = np.ones(100)
    ls = kp.linsys.LinearSystem(A, b, self_adjoint=True, positive_definite=True)
    for cls, it in ((kp.linsys.Cg, 55), (kp.linsys.Minres, 54), (kp.linsys.Gmres, 54)):
        sol = cls(ls, store_arnoldi=True)
        assert sol.iter == it
        assert sol.V.shape == (100, 56) and sol.H.shape == (56, 55) and sol.xk.shape == (100, 1)
        assert isinstance(sol.resnorms, list) and isinstance(sol.resnorms[1], np.floating)
    for cls, vs, hs in ((kp.linsys.Cg, (100, 10), (10, 9)), (kp.linsys.Minres, (100, 11), (11, 10)),
                        (kp.linsys.Gmres, (100, 11), (11, 10))):
        with pytest.raises(kp.utils.ConvergenceError) as ei:
            cls(ls, maxiter=10, store_arnoldi=True)
        s = ei.value.solver
        assert s.xk is not None and len(s.resnorms) - 1 == 10 and s.iter == 9
        assert s.V.shape == vs and s.H.shape == hs
    # zero right hand side
    z = kp.linsys.Gmres(kp.linsys.LinearSystem(A, np.zeros((100, 1))))
    assert z.resnorms == [0.0] and np.all(z.xk == 0) and z.xk.shape == (100, 1)
    # exact initial guess
    x = np.linalg.solve(A, b).reshape(-1, 1)
    e = kp.linsys.Gmres(kp.linsys.LinearSystem(A, b.reshape(-1, 1)), x0=x, tol=1e-10)
    assert len(e.resnorms) == 1 and e.iter == 0
    np.testing.assert_allclose(e.xk, x)
    # RestartedGmres surface
    r = kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=10)
    assert r.resnorms[-1] <= 1e-5 and r.xk.shape == (100, 1)
    with pytest.raises(kp.utils.ConvergenceError):
        kp.linsys.RestartedGmres(ls, maxiter=5, max_restarts=1)
    # deflated attributes
    d = kp.deflation.DeflatedGmres(kp.linsys.LinearSystem(A, b), U=np.eye(100, 2), store_arnoldi=True)
    n = d.H.shape[1]
    assert d.C.shape == (2, n) and d.E.shape == (2, 2) and d.B_.shape == (n + 1, 2) and d.UMlr.shape == (2, 1)
    # final residual is what it claims to be (reference test_linsys.py:189-195)
    for sol in (kp.linsys.Gmres(ls, tol=1e-9), kp.linsys.Cg(ls, tol=1e-9), kp.linsys.Minres(ls, tol=1e-9)):
        _, _, rn = ls.get_residual(sol.xk, compute_norm=True)
        np.testing.assert_almost_equal(sol.resnorms[-1], rn / ls.MMlb_norm, decimal=13)


def test_operator_algebra_and_errors():
    import krypy_b200 as kp
    u = kp.utils
    rng = np.random.default_rng(0)
    A = rng.standard_normal((12, 12)); B = rng.standard_normal((12, 12)); X = rng.standard_normal((12, 3))
    a, bop = u.MatrixLinearOperator(A), u.MatrixLinearOperator(B)
    np.testing.assert_allclose((a * bop) * X, A @ B @ X, rtol=1e-12)
    np.testing.assert_allclose((a + bop) * X, (A + B) @ X, rtol=1e-12)
    np.testing.assert_allclose((a - bop) * X, (A - B) @ X, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose((2.5 * a) * X, 2.5 * A @ X, rtol=1e-12)
    np.testing.assert_allclose((a ** 2) * X, A @ A @ X, rtol=1e-12)
    np.testing.assert_allclose(a.adj * X, A.T @ X, rtol=1e-12)
    I = u.IdentityLinearOperator((12, 12))
    assert (I * a) is a and (a * I) is a
    assert (a * X[:, :0]).shape == (12, 0)
    with pytest.raises(u.LinearOperatorError):
        u.get_linearoperator((3, 3), A)
    with pytest.raises(TypeError):
        u.get_linearoperator((3, 3), "nope")
    assert kp.linsys.LinearSystem(A.astype(complex), np.ones(12)).dtype == np.complex128   # complex: see test_zcomplex_gpu.py
    # inner / norm / qr / Projection properties (reference test_utils.py:157-246)
    Bip = np.diag(np.linspace(1, 2, 12))
    np.testing.assert_allclose(u.inner(X, X, ip_B=Bip), X.T @ Bip @ X, rtol=1e-12)
    np.testing.assert_allclose(u.norm(X[:, [0]]), np.linalg.norm(X[:, 0]), rtol=1e-13)
    Q, R = u.qr(X, ip_B=Bip)
    np.testing.assert_allclose(Q.T @ Bip @ Q, np.eye(3), atol=1e-13)
    np.testing.assert_allclose(Q @ R, X, atol=1e-13)
    Y = rng.standard_normal((12, 3))
    for its in (1, 2, 3):
        P = u.Projection(X, Y, ip_B=Bip, iterations=its)
        z = rng.standard_normal((12, 2))
        Pz = P.apply(z)
        np.testing.assert_allclose(P.apply(Pz), Pz, atol=1e-12)              # P^2 = P
        np.testing.assert_allclose(Pz + P.apply_complement(z), z, atol=1e-12)
        np.testing.assert_allclose(u.inner(Y, P.apply_complement(z), ip_B=Bip), 0, atol=1e-12)
        _, Ya = P.apply(z, return_Ya=True)
        np.testing.assert_allclose(Ya, u.inner(Y, z, ip_B=Bip), atol=1e-12)
    Pe = u.Projection(X)
    np.testing.assert_allclose(X.T @ Pe.apply_complement(z), 0, atol=1e-12)


@pytest.mark.parametrize("ortho", ["mgs", "dmgs", "lanczos", "cgs2"])
@pytest.mark.parametrize("with_M", [False, True])
@pytest.mark.parametrize("ipB", ["none", "matrix", "callable"])
def test_arnoldi_relation(ortho, with_M, ipB):
    """reference test_utils.py:440-542 (assert_arnoldi): first vector, Hessenberg
    structure, Arnoldi residual and orthogonality"""
    import krypy_b200 as kp
    u = kp.utils
    rng = np.random.default_rng(5)
    N = 10
    Bm = np.diag(np.linspace(1, 5, N))
    sym = ortho == "lanczos"
    K = rng.standard_normal((N, N))
    K = K + K.T if sym else K
    ip_B = {"none": None, "matrix": Bm, "callable": (lambda X, Y: X.T.conj() @ Bm @ Y)}[ipB]
    Bmat = np.eye(N) if ipB == "none" else Bm
    M = 2.0 * np.eye(N) if with_M else None
    A = np.linalg.solve(Bmat, K) if sym else K          # self-adjoint in <.,.>_B when needed
    v = rng.standard_normal((N, 1))
    res = u.arnoldi(A, v, maxiter=6, ortho=ortho, M=M, ip_B=ip_B)
    V, H = res[0], res[1]
    n = H.shape[1]
    assert V.shape == (N, n + 1) and H.shape == (n + 1, n)
    assert np.allclose(np.tril(H, -2), 0)
    assert np.all(np.diag(H, -1) >= 0)
    MA = (M @ A) if with_M else A
    ipM = Bmat if not with_M else np.linalg.inv(M) @ Bmat
    np.testing.assert_allclose(MA @ V[:, :n], V @ H, atol=1e-12 * np.linalg.norm(MA))
    np.testing.assert_allclose(V.T @ ipM @ V, np.eye(n + 1), atol=1e-11)
    if with_M:
        np.testing.assert_allclose(M @ res[2], V, atol=1e-13)


@pytest.mark.parametrize("name", ["convdiff_defl_gmres", "lap2d_defl_minres", "lap2d_defl_cg", "c1_gmres_defl"])
def test_ritz_pairs_match_reference(name):
    """deflation.Ritz against the reference's values / residual norms"""
    import ritz_checks
    ritz_checks.check_ritz_pairs(name)


@pytest.mark.parametrize("sname", ["cg", "minres", "gmres"])
@pytest.mark.parametrize("which", ["sm", "lm", "smallest_res"])
def test_recycling_ritz_factory_simple(sname, which):
    """reference test/test_recycling.py: three recycled solves, iteration counts as the reference"""
    import ritz_checks
    ritz_checks.check_recycling(sname, which)


@pytest.mark.parametrize("withM", [True, False])
@pytest.mark.parametrize("tol,restarts", [(1e-14, 4), (1e-6, 60)])
@pytest.mark.parametrize("graphs", ["on", "off"])
def test_restarted_gmres_preconditioned_graph_replay(graphs, tol, restarts, withM):
    """CUDA-graph replay of Arnoldi steps across >= 3 restart cycles with a preconditioner M (second
    basis P, scratch vector): the recorded graphs must find the same buffers in every cycle
    (round-1 advisor finding: P and the scratch were re-allocated per cycle).  From the third cycle on the
    whole cycle is ONE graph and its records are booked in bulk; the 1e-6 case converges in the middle of
    such a cycle (the steps behind the converged one are speculative and must leave no trace).  Without M
    every such cycle also launches its successor speculatively on the device-side residual norm (the last
    one for nothing: the solve converges)."""
    import krypy_b200 as kp
    from krypy_b200 import problems
    from oracle import krylov_oracle as ko
    import scipy.sparse as sp
    n = 24
    A = problems.laplace2d(n)
    rng = np.random.default_rng(3)
    dM = sp.diags(1.0 / (4.0 + rng.random(n * n))).tocsr()
    b = rng.standard_normal((n * n, 1))
    if not withM:
        dM = None
    ls = kp.linsys.LinearSystem(A, b, M=dM)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = kp.linsys.RestartedGmres(ls, maxiter=8, max_restarts=restarts, tol=tol, ortho="cgs",
                                           _workspace=kp.utils.SolverWorkspace(graphs=graphs))
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        try:
            ref = ko.restarted_gmres(ko.System(A, b, M=dM), maxiter=8, max_restarts=restarts, tol=tol)
        except ko.OracleConvergenceError as e:
            ref = e.result
    _check_history(np.array(sol.resnorms), np.array(ref.resnorms))
    hits = getattr(sol._workspace, "prelaunch_hits", 0)
    ncycles = (len(sol.resnorms) - 1 + 7) // 8
    # (fewer when the last cycles before convergence ended within 2 tol: no speculation that close to it)
    assert (hits == 0) if withM else (max(ncycles - 7, min(ncycles - 3, 1), 0) <= hits <= max(ncycles - 3, 0)), \
        (hits, ncycles)     # speculative starts that were used
    if tol > 1e-10:
        assert sol.resnorms[-1] <= tol and len(sol.resnorms) > 3 * 8
        np.testing.assert_allclose(sol.xk, ref.xk, rtol=1e-8, atol=1e-10)


def test_cholqr2_projector_setup_matches_mgs():
    """CholQR2 set-up of the deflation projector (kry_gram + kry_block_trsm, sizes >= 4096) against the
    column-by-column MGS path: factorisation, and the deflated history at the 1e-10 contract"""
    import krypy_b200 as kp
    from krypy_b200 import problems
    import test_host_logic_cpu
    test_host_logic_cpu.check_cholqr2_against_mgs(kp, N=5000, k=20)
    n = 96
    A = problems.convdiff2d(n, c=0.1)
    N = n * n
    b = np.ones((N, 1))
    xs = np.arange(1, n + 1) / (n + 1.0)
    U = np.stack([np.outer(np.sin(p * np.pi * xs), np.sin(q * np.pi * xs)).reshape(-1)
                  for p in range(1, 4) for q in range(1, 4)], axis=1)
    hist = {}
    old = kp.utils._BLOCK_MIN_N
    try:
        for name, thr in (("cholqr2", 0), ("mgs", 10 ** 9)):
            kp.utils._BLOCK_MIN_N = thr
            ls = kp.linsys.LinearSystem(A, b)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    s = kp.deflation.DeflatedGmres(ls, U=U, maxiter=40, tol=1e-12)
                except kp.utils.ConvergenceError as e:
                    s = e.solver
            hist[name] = (np.array(s.resnorms), s.E, s.UMlr)
    finally:
        kp.utils._BLOCK_MIN_N = old
    _check_history(hist["cholqr2"][0], hist["mgs"][0])
    np.testing.assert_allclose(hist["cholqr2"][1], hist["mgs"][1], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(hist["cholqr2"][2], hist["mgs"][2], rtol=1e-9, atol=1e-12)
