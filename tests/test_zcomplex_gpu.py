"""Complex128 systems on the GPU (real embedding + twin storage, krypy_b200/_cplx.py): the new
entry points (kry_rot90, kry_givens_update_z, kry_tri_solve_z) against numpy, the embedded
operators against numpy's complex products, and the solver classes against the reference's
complex fixtures (tests/golden/z_*.npz, oracle/make_golden.py) and the oracle.

Tolerances as in test_solvers_gpu.py: residual histories |d| <= 1e-10*res_k + 1e-13.

(File name sorts last on purpose: these tests were written in a session that had no GPU time
left -- the host logic is covered on the CPU tier over the test double, the serial cores of the
two new recurrences by tests/test_small_core_cpu.py -- so a surprise here cannot hide the
established parity tests behind pytest -x.)"""
import warnings

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp

import cases
import runners

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from krypy_b200 import _device
    assert torch.cuda.is_available()
    return _device.Context.get()


def T(ctx, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# ------------------------------------------------------------------ new entry points
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 100003])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_rot90_is_multiplication_by_i(ctx, n, dt):
    rng = np.random.default_rng(n)
    z = crandn(rng, n).astype(np.complex128 if dt == np.float64 else np.complex64)
    x = T(ctx, z.view(dt))
    y = ctx.empty((2 * n,), x.dtype)
    ctx.rot90(x, y)
    got = y.cpu().numpy().view(z.dtype)
    assert np.array_equal(got, 1j * z)                          # exact: a swap and a sign


def test_rot90_on_complex_tensors_and_twin_storage(ctx):
    import torch
    from krypy_b200 import utils
    rng = np.random.default_rng(0)
    k, N = 3, 101
    Z = crandn(rng, k, N)
    Zd = T(ctx, Z)
    assert Zd.dtype == torch.complex128
    tw = utils._Twin.of(ctx, Zd)
    assert tw.T.shape == (2 * k, 2 * N) and tw.C.shape == (k, N)
    Tn = tw.T.cpu().numpy()
    assert np.array_equal(Tn[0::2].copy().view(np.complex128), Z)
    assert np.array_equal(Tn[1::2].copy().view(np.complex128), 1j * Z)
    assert np.array_equal(tw.C.cpu().numpy(), Z)
    # real dots against the twin rows are the interleaved complex coefficients <z_j, q>
    q = crandn(rng, N)
    out = ctx.scalars(2 * k)
    ctx.block_dot(tw.T, 2 * k, T(ctx, q), out)
    got = out.cpu().numpy().view(np.complex128)
    np.testing.assert_allclose(got, Z.conj() @ q, rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize("real_valued", [False, True])
def test_givens_update_z_against_numpy(ctx, real_valued):
    """drive the device recurrence column by column like Gmres._solve and compare the mailbox, R, y
    with numpy (same drotg / zrotg switch as krypy/utils.py:419-427)"""
    import fake_device                     # numpy restatement of the kernel contract (test infra)
    rng = np.random.default_rng(5)
    m = 9
    H = np.triu(rng.standard_normal((m + 1, m)) + (0 if real_valued else 1j) * rng.standard_normal((m + 1, m)), -1)
    H = H.astype(np.complex128)
    for k in range(m):
        H[k + 1, k] = abs(H[k + 1, k])
    beta = 2.5
    import torch
    fake = fake_device.FakeContext()
    f = dict(h=torch.zeros(2 * (m + 2), dtype=torch.float64), r=torch.zeros(2 * (m + 2), dtype=torch.float64),
             cs=torch.zeros(4 * (m + 1), dtype=torch.float64), y=torch.zeros(2 * (m + 2), dtype=torch.float64))
    d = {k_: ctx.scalars(v.numel()) for k_, v in f.items()}
    f["y"][0] = beta
    d["y"][0:1].fill_(beta)
    for k in range(m):
        col = H[: k + 2, k].copy().view(np.float64)
        f["h"][: col.size] = torch.from_numpy(col)
        d["h"][: col.size].copy_(T(ctx, col))
        off = (k & 1) * 4096
        fake.givens_update_z(k, f["h"], f["r"], f["cs"], f["y"], off)
        ctx.givens_update_z(k, d["h"], d["r"], d["cs"], d["y"], off)
        ctx.sync()
        n = 4 * (k + 2) + 1
        np.testing.assert_allclose(ctx.mailbox[off:off + n], fake.mailbox[off:off + n], rtol=1e-13, atol=1e-15)
        for key in ("r", "cs", "y"):
            np.testing.assert_allclose(d[key].cpu().numpy(), f[key].numpy(), rtol=1e-13, atol=1e-15, err_msg=key)
        assert np.all(d["h"].cpu().numpy() == 0.0)                       # accumulator left zeroed
        # residual norm of the least-squares problem
        e1 = np.zeros(k + 2, dtype=np.complex128)
        e1[0] = beta
        sol = np.linalg.lstsq(H[: k + 2, : k + 1], e1, rcond=None)[0]
        want = np.linalg.norm(H[: k + 2, : k + 1] @ sol - e1)
        assert abs(ctx.mailbox[off] - want) <= 1e-13 * beta


def test_tri_solve_z_against_scipy(ctx):
    rng = np.random.default_rng(6)
    for k in (1, 2, 9, 40):
        R = np.triu(crandn(rng, k, k)) + 3 * np.eye(k)
        y = crandn(rng, k)
        Rd = T(ctx, R.view(np.float64))                                   # (k, 2k) interleaved
        out = ctx.scalars(2 * k)
        ctx.tri_solve_z(k, Rd, T(ctx, y.view(np.float64)), out)
        got = out.cpu().numpy().view(np.complex128)
        np.testing.assert_allclose(got, scipy.linalg.solve_triangular(R, y), rtol=1e-12)


# ------------------------------------------------------------------ embedded operators
def test_embedded_operators_apply_complex_products(ctx):
    from krypy_b200 import utils
    rng = np.random.default_rng(7)
    N, k = 57, 3
    X = crandn(rng, N, k)
    As = sp.random(N, N, density=0.1, random_state=3, format="csr") * (1 + 0j)
    As.data = As.data + 1j * rng.standard_normal(As.nnz)
    Ad = crandn(rng, N, N)
    d = rng.standard_normal(N)
    dz = crandn(rng, N)
    for op, ref in ((utils.MatrixLinearOperator(As), As @ X), (utils.MatrixLinearOperator(Ad), Ad @ X),
                    (utils.MatrixLinearOperator(As.real.tocsr()), As.real @ X),
                    (utils.DiagonalLinearOperator(d), d[:, None] * X),
                    (utils.DiagonalLinearOperator(dz), dz[:, None] * X),
                    ((2 - 3j) * utils.MatrixLinearOperator(Ad), (2 - 3j) * (Ad @ X)),
                    (utils.MatrixLinearOperator(As).adj, As.conj().T @ X),
                    (((2 - 3j) * utils.MatrixLinearOperator(Ad)).adj, np.conj(2 - 3j) * (Ad.conj().T @ X))):
        got = op * X
        assert got.dtype == np.complex128 and got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-13)
    Y = crandn(rng, N, 2)
    B = np.diag(np.linspace(1, 2, N))
    np.testing.assert_allclose(utils.inner(X, Y), X.conj().T @ Y, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(utils.inner(X, Y, ip_B=B), X.conj().T @ B @ Y, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(utils.norm(X[:, [0]]), np.linalg.norm(X[:, 0]), rtol=1e-13)
    Q, R = utils.qr(X)
    np.testing.assert_allclose(Q @ R, X, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Q.conj().T @ Q, np.eye(k), atol=1e-13)
    Q, R = utils.qr(X, ip_B=B)
    np.testing.assert_allclose(Q @ R, X, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Q.conj().T @ B @ Q, np.eye(k), atol=1e-13)


def test_complex_projection_matches_dense_formula(ctx):
    from krypy_b200 import utils
    rng = np.random.default_rng(8)
    N, k = 64, 4
    X, Y, a = crandn(rng, N, k), crandn(rng, N, k), crandn(rng, N, 2)
    P = utils.Projection(X, Y)
    Pm = X @ np.linalg.solve(Y.conj().T @ X, Y.conj().T)
    np.testing.assert_allclose(P.apply(a), Pm @ a, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(P.apply_complement(a), a - Pm @ a, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(P.apply_adj(a), Pm.conj().T @ a, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(P.apply_complement_adj(a), a - Pm.conj().T @ a, rtol=1e-11, atol=1e-12)


# ------------------------------------------------------------------ solvers vs the reference
def _check_history(got, ref, rtol=1e-10, atol=1e-13):
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + atol
    assert np.all(err <= bound), (np.argmax(err / bound), (err / bound).max())


@pytest.mark.parametrize("name", cases.COMPLEX_CASES)
def test_complex_cases_match_reference_fixture_and_oracle(name):
    gold = runners.load_golden(name)
    orac = runners.run_oracle(name)
    got = runners.run_product(name)
    assert got["xk"].dtype == np.complex128
    assert bool(got["converged"]) == bool(gold["converged"])
    for ref in (gold, orac):
        _check_history(got["resnorms"], ref["resnorms"])
        scale = np.abs(ref["xk"]).max() + 1e-300
        assert np.abs(got["xk"] - ref["xk"]).max() <= 1e-8 * scale
        for k in ("iter", "V_shape"):
            if k in ref:
                assert np.array_equal(got[k], ref[k]), k
        rn = ref["resnorms"]
        ngood = int(np.argmax(rn < 1e-7)) if np.any(rn < 1e-7) else len(rn)
        for k in ("H", "C", "V_colsum_abs"):
            if k in ref:
                assert got[k].shape == ref[k].shape, k
                nc = min(max(ngood - 1, 0), ref[k].shape[-1])
                a, b = got[k][..., :nc], ref[k][..., :nc]
                if a.size:
                    assert np.abs(a - b).max() <= 1e-6 * (np.abs(b).max() + 1e-300), k
        for k in ("E", "UMlr", "rhos"):
            if k in ref:
                assert got[k].shape == ref[k].shape, k
                np.testing.assert_allclose(got[k], ref[k], rtol=1e-8, atol=1e-13 * (np.abs(ref[k]).max() + 1e-300))


@pytest.mark.parametrize("ortho", ["cgs", "cgs2"])
@pytest.mark.parametrize("name", ["z_gmres_helmholtz", "z_defl_gmres"])
def test_complex_block_gram_schmidt_matches_reference_mgs(name, ortho):
    gold = runners.load_golden(name)
    got = runners.run_product(name, ortho=ortho)
    _check_history(got["resnorms"], gold["resnorms"])


def test_complex_arnoldi_relation_and_orthonormality(ctx):
    """test/test_utils.py:440-542 style: A V_k = V_{k+1} H, V^H V = I, for complex data through the
    twin-storage Gram-Schmidt kernel (all ortho variants)"""
    from krypy_b200 import utils
    rng = np.random.default_rng(9)
    N = 300
    A = sp.random(N, N, density=0.03, random_state=1, format="csr") * (1 + 0j)
    A.data = A.data + 1j * rng.standard_normal(A.nnz)
    A = sp.csr_matrix(A + 4 * sp.identity(N))
    v = crandn(rng, N, 1)
    for ortho in ("mgs", "dmgs", "cgs", "cgs2"):
        V, H = utils.arnoldi(A, v, maxiter=20, ortho=ortho)
        assert V.dtype == np.complex128 and H.dtype == np.complex128 and V.shape == (N, 21)
        assert np.linalg.norm(A @ V[:, :-1] - V @ H) <= 1e-12 * np.linalg.norm(H)
        assert np.linalg.norm(V.conj().T @ V - np.eye(21)) <= (1e-9 if ortho in ("mgs", "cgs") else 1e-13)
        assert np.abs(np.tril(H, -2)).max() == 0.0 and np.abs(np.diag(H, -1).imag).max() == 0.0


def test_complex_convenience_wrappers(ctx):
    import krypy_b200 as kp
    rng = np.random.default_rng(10)
    c = cases.case_inputs("z_cg_hpd")
    A, b = c["A"], c["b"].reshape(-1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for fn in (kp.cg, kp.minres, kp.gmres):
            x, sol = fn(A, b, tol=1e-9, maxiter=200)
            assert x is not None and x.shape == b.shape and x.dtype == np.complex128
            assert np.linalg.norm(A @ x - b) <= 1e-8 * np.linalg.norm(b)


# ------------------------------------------------------------------ Householder Arnoldi (ortho='house')
@pytest.mark.parametrize("cplx", [False, True])
def test_householder_arnoldi(ctx, cplx):
    """krypy/utils.py:970-994 on the device (reflectors in HBM, applied with kry_block_dot/axpy):
    Arnoldi relation, orthonormality to machine precision, and the same H as modified Gram-Schmidt"""
    from krypy_b200 import utils
    rng = np.random.default_rng(11)
    N = 120
    A = rng.standard_normal((N, N)) + 6 * np.eye(N)
    v = rng.standard_normal((N, 1))
    if cplx:
        A = A + 1j * rng.standard_normal((N, N))
        v = v + 1j * rng.standard_normal((N, 1))
    V, H = utils.arnoldi(A, v, maxiter=15, ortho="house")
    Vm, Hm = utils.arnoldi(A, v, maxiter=15, ortho="dmgs")
    assert V.shape == (N, 16) and H.shape == (16, 15)
    assert np.linalg.norm(A @ V[:, :-1] - V @ H) <= 1e-12 * np.linalg.norm(H)
    assert np.linalg.norm(V.conj().T @ V - np.eye(16)) <= 1e-13
    np.testing.assert_allclose(H, Hm, rtol=0, atol=1e-10 * np.abs(Hm).max())
    np.testing.assert_allclose(V, Vm, rtol=0, atol=1e-10)
    # small host helper, utils.py:332-402
    x = v[:7]
    h = utils.House(x)
    y = h.apply(x)
    assert abs(y[0, 0] - h.alpha * np.linalg.norm(x)) <= 1e-13 and np.abs(y[1:]).max() <= 1e-13
    # full-dimensional run ends with an invariant subspace (k+1 == N branch)
    Vf, Hf = utils.arnoldi(A[:9, :9], v[:9], ortho="house")
    assert Vf.shape == (9, 9) and Hf.shape == (9, 9)


# ------------------------------------------------------------------ real system, complex x0 / complex U
def test_real_system_becomes_complex_for_complex_x0_and_deflation_vectors(ctx):
    """SURVEY 3.6 / F10 (krypy/linsys.py:370-372, deflation.py:123-125): a complex x0 or complex
    deflation vectors make the whole solve complex; the result must equal the solve of the explicitly
    complexified system"""
    import krypy_b200 as kp
    from krypy_b200 import problems
    rng = np.random.default_rng(12)
    n = 12
    N = n * n
    A = problems.convdiff2d(n, c=0.4)
    b = rng.standard_normal((N, 1))
    x0 = crandn(rng, N, 1)
    U = crandn(rng, N, 3)

    def run(cls, ls, **kw):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                return cls(ls, **kw)
            except kp.utils.ConvergenceError as e:
                return e.solver

    Ac, bc = sp.csr_matrix(A.astype(np.complex128)), b.astype(np.complex128)
    for cls, kw in ((kp.linsys.Gmres, dict(x0=x0)), (kp.deflation.DeflatedGmres, dict(U=U)),
                    (kp.deflation.DeflatedGmres, dict(U=U, x0=x0))):
        ls = kp.linsys.LinearSystem(A, b)
        assert ls.dtype == np.float64
        got = run(cls, ls, tol=1e-9, maxiter=40, **kw)
        ref = run(cls, kp.linsys.LinearSystem(Ac, bc), tol=1e-9, maxiter=40, **kw)
        assert got.dtype == np.complex128 and got.xk.dtype == np.complex128
        assert ls.dtype == np.float64                       # the user's system object is left as it was
        _check_history(np.array(got.resnorms), np.array(ref.resnorms))
        np.testing.assert_allclose(got.xk, ref.xk, rtol=1e-9, atol=1e-12)
    # float64 x0 with an fp32-storage system promotes the solve to fp64
    L = problems.laplace2d(n)
    ls32 = kp.linsys.LinearSystem(L.astype(np.float32), b.astype(np.float32), dtype=np.float32,
                                  self_adjoint=True, positive_definite=True)
    s = run(kp.linsys.Cg, ls32, x0=np.zeros((N, 1)), tol=1e-8)
    assert s.dtype == np.float64 and s.resnorms[-1] <= 1e-8
