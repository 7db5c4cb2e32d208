"""Named parity cases shared by oracle/make_golden.py (reference run ->
fixtures), tests/test_oracle_golden.py (oracle vs fixtures) and the GPU parity
tests (CUDA path vs oracle and vs fixtures).  Inputs are regenerated from seeds,
so the fixtures only hold the reference's OUTPUTS.
"""
import numpy as np
import scipy.sparse as sp

from krypy_b200 import problems


def _diag100():
    return np.diag([1.0e-3] + list(range(2, 101))).astype(np.float64)


def case_inputs(name):
    """Returns dict(A, b, solver, ls_kwargs, solver_kwargs)."""
    rng = np.random.default_rng(1234)
    c = dict(ls={}, kw={})
    if name.startswith("c1_"):
        # BASELINE config 1: README example, test_convenience_wrappers.py:10-31
        c["A"] = _diag100()
        c["b"] = np.ones(100)
        c["solver"] = name.split("_")[1]           # gmres | cg | minres
        if c["solver"] in ("cg",):
            c["ls"] = dict(self_adjoint=True, positive_definite=True)
        if c["solver"] == "minres":
            c["ls"] = dict(self_adjoint=True)
            c["kw"] = dict(ortho="mgs")            # _convenience.py:91
        if name.endswith("_defl"):
            U = np.zeros((100, 1)); U[0] = 1.0     # test_convenience_wrappers.py:34-55
            c["kw"]["U"] = U
        if name.endswith("_store"):
            c["kw"]["store_arnoldi"] = True
    elif name in ("lap2d_gmres30", "lap2d_gmres_mgs", "lap2d_gmres_dmgs"):
        n = 24
        c["A"] = problems.laplace2d(n)
        c["b"] = problems.rhs_normal(n * n)
        if name == "lap2d_gmres30":
            c["solver"] = "restarted_gmres"
            c["kw"] = dict(maxiter=30, max_restarts=2, tol=1e-12)
        else:
            c["solver"] = "gmres"
            c["kw"] = dict(maxiter=40, tol=1e-12, store_arnoldi=True,
                           ortho=name.rsplit("_", 1)[1])
    elif name == "poisson3d_cg_jacobi":
        n = 8
        A = problems.poisson3d(n)
        c["A"] = A
        c["b"] = problems.rhs_normal(n ** 3)
        M = problems.jacobi_csr(A)
        c["ls"] = dict(M=M, Minv=sp.diags(A.diagonal()).tocsr(), self_adjoint=True,
                       positive_definite=True)
        c["solver"] = "cg"
        c["kw"] = dict(tol=1e-8, maxiter=200, store_arnoldi=True)
    elif name == "convdiff_defl_gmres":
        n = 20
        c["A"] = problems.convdiff2d(n, c=0.1)
        c["b"] = np.ones((n * n, 1))
        c["solver"] = "gmres"
        c["kw"] = dict(U=rng.standard_normal((n * n, 5)), maxiter=30, tol=1e-10,
                       store_arnoldi=True)
    elif name == "shifted_minres_ipB":
        n = 48      # 30 steps at N=2304: no Ritz value converges, the history is well conditioned
        A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32)
        c["A"] = A
        c["b"] = problems.rhs_normal(n * n, dtype=np.float32)
        c["ls"] = dict(ip_B=B, self_adjoint=True)
        c["solver"] = "minres"
        c["kw"] = dict(tol=1e-5, maxiter=30)
    elif name == "shifted_minres_ipB_f64":
        n = 48
        A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float64)
        c["A"] = A
        c["b"] = problems.rhs_normal(n * n)
        c["ls"] = dict(ip_B=B, self_adjoint=True)
        c["solver"] = "minres"
        c["kw"] = dict(tol=1e-9, maxiter=30, store_arnoldi=True)
    elif name in ("dense_gmres_M_ipB", "dense_minres_M_ipB", "dense_cg_M_ipB"):
        # preconditioned + non-Euclidean inner product (test_utils.py:355-383 style)
        N = 30
        d = np.linspace(1, 4, N)
        Bm = np.diag(np.linspace(1, 2, N))
        S = rng.standard_normal((N, N)); S = S + S.T
        K = np.diag(d) + 0.05 * S + 2 * np.eye(N)          # SPD-ish, symmetric
        c["A"] = np.linalg.solve(Bm, K)                    # self-adjoint in <.,.>_B
        c["b"] = rng.standard_normal((N, 1))
        # M must be self-adjoint pos.def. w.r.t. <.,.>_B and keep M*A self-adjoint
        # in <.,.>_{M^{-1}B}: a scalar multiple of the identity is.
        Mm = 0.5 * np.eye(N)
        c["ls"] = dict(M=Mm, Minv=2.0 * np.eye(N), ip_B=Bm, self_adjoint=True,
                       positive_definite=True)
        c["solver"] = name.split("_")[1]
        c["kw"] = dict(tol=1e-11, maxiter=40, store_arnoldi=True)
    elif name == "dense_gmres_MlMr":
        N = 40
        A = np.diag(np.linspace(1, 10, N)) + 0.3 * rng.standard_normal((N, N))
        c["A"] = A
        c["b"] = rng.standard_normal((N,))
        c["ls"] = dict(Ml=np.diag(1.0 / np.linspace(1, 10, N)),
                       Mr=np.diag(np.linspace(0.5, 1.5, N)))
        c["solver"] = "gmres"
        c["kw"] = dict(tol=1e-12, maxiter=40, x0=rng.standard_normal((N, 1)),
                       store_arnoldi=True)
    elif name == "lucky_breakdown":
        # 3 distinct eigenvalues, tol 1e-14
        N = 10
        c["A"] = np.diag([1.0] * 4 + [2.0] * 3 + [5.0] * 3)
        c["b"] = np.ones((N, 1))
        c["solver"] = "gmres"
        c["kw"] = dict(tol=1e-14, store_arnoldi=True)
    elif name == "zero_rhs":
        c["A"] = _diag100()
        c["b"] = np.zeros((100, 1))
        c["solver"] = "gmres"
    elif name == "maxiter_noconv":
        c["A"] = _diag100()
        c["b"] = np.ones((100, 1))
        c["solver"] = "gmres"
        c["kw"] = dict(maxiter=10, store_arnoldi=True)
    elif name == "cg_explicit_residual":
        n = 12
        c["A"] = problems.laplace2d(n)
        c["b"] = problems.rhs_normal(n * n)
        c["ls"] = dict(self_adjoint=True, positive_definite=True)
        c["solver"] = "cg"
        c["kw"] = dict(tol=1e-9, maxiter=80, explicit_residual=True)
    elif name in ("lap2d_defl_cg", "lap2d_defl_minres"):
        n = 14
        c["A"] = problems.laplace2d(n)
        c["b"] = problems.rhs_normal(n * n)
        c["ls"] = dict(self_adjoint=True, positive_definite=True)
        c["solver"] = name.rsplit("_", 1)[1]
        c["kw"] = dict(U=rng.standard_normal((n * n, 3)), tol=1e-9, maxiter=60,
                       store_arnoldi=True)
    elif name.startswith("z_"):
        _complex_case(name, c, rng)
    else:
        raise KeyError(name)
    return c


def _herm_sparse(n, theta=0.3):
    """Hermitian positive definite sparse test matrix: the 2-D 5-point Laplacian with complex hopping
    phases exp(+-i*theta) on the x-bonds (a "magnetic" Laplacian)."""
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    Tp = sp.diags([-np.exp(-1j * theta), 2.0, -np.exp(1j * theta)], [-1, 0, 1], shape=(n, n))
    I = sp.identity(n)
    return sp.csr_matrix(sp.kron(I, Tp) + sp.kron(T, I))


def _complex_case(name, c, rng):
    """complex128 cases (half of the reference's own test matrix is complex, test/test_linsys.py:118-141)"""
    def crandn(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    if name in ("z_gmres_helmholtz", "z_gmres_helmholtz_dmgs", "z_restarted_gmres"):
        # damped Helmholtz: (L - sigma I) + i*delta I, complex symmetric, non-Hermitian
        n = 20
        A = problems.laplace2d(n).astype(np.complex128) + (-0.4 + 0.35j) * sp.identity(n * n)
        c["A"] = sp.csr_matrix(A)
        c["b"] = crandn(n * n, 1)
        if name == "z_restarted_gmres":
            c["solver"] = "restarted_gmres"
            c["kw"] = dict(maxiter=15, max_restarts=3, tol=1e-12)
        else:
            c["solver"] = "gmres"
            c["kw"] = dict(maxiter=40, tol=1e-12, store_arnoldi=True,
                           ortho="dmgs" if name.endswith("dmgs") else "mgs")
    elif name == "z_gmres_realA_complexb":
        n = 16
        c["A"] = problems.convdiff2d(n, c=0.1)
        c["b"] = crandn(n * n)
        c["solver"] = "gmres"
        c["kw"] = dict(maxiter=40, tol=1e-11, x0=crandn(n * n, 1))
    elif name in ("z_cg_hpd", "z_minres_herm", "z_minres_herm_mgs", "z_defl_cg", "z_defl_minres"):
        n = 14
        A = _herm_sparse(n)
        c["b"] = crandn(n * n, 1)
        if "minres" in name:
            A = sp.csr_matrix(A - 1.1 * sp.identity(n * n))         # Hermitian indefinite
            c["ls"] = dict(self_adjoint=True)
            c["solver"] = "minres"
            c["kw"] = dict(tol=1e-9, maxiter=30, store_arnoldi=True)
            if name.endswith("mgs"):
                c["kw"]["ortho"] = "mgs"
        else:
            c["ls"] = dict(self_adjoint=True, positive_definite=True)
            c["solver"] = "cg"
            c["kw"] = dict(tol=1e-9, maxiter=60, store_arnoldi=True)
        c["A"] = A
        if "defl" in name:
            c["kw"]["U"] = crandn(n * n, 3)
    elif name in ("z_defl_gmres", "z_defl_gmres_ipB"):
        n = 16
        A = problems.convdiff2d(n, c=0.1).astype(np.complex128) + 0.25j * sp.identity(n * n)
        c["A"] = sp.csr_matrix(A)
        c["b"] = crandn(n * n, 1)
        c["solver"] = "gmres"
        c["kw"] = dict(U=crandn(n * n, 4), maxiter=30, tol=1e-10, store_arnoldi=True)
        if name.endswith("ipB"):
            c["ls"] = dict(ip_B=sp.diags(np.linspace(1.0, 2.0, n * n)).tocsr())
    elif name in ("z_dense_gmres_M_ipB", "z_dense_minres_M_ipB", "z_dense_cg_M_ipB"):
        N = 30
        S = crandn(N, N)
        S = S + S.conj().T
        Bs = crandn(N, N)
        Bm = np.eye(N) * 1.5 + 0.02 * (Bs + Bs.conj().T)          # Hermitian positive definite
        K = np.diag(np.linspace(1, 4, N)) + 0.05 * S + 2 * np.eye(N)
        c["A"] = np.linalg.solve(Bm, K)                            # self-adjoint in <.,.>_B
        c["b"] = crandn(N, 1)
        c["ls"] = dict(M=0.5 * np.eye(N), Minv=2.0 * np.eye(N), ip_B=Bm, self_adjoint=True,
                       positive_definite=True)
        c["solver"] = name.split("_")[2]
        c["kw"] = dict(tol=1e-11, maxiter=40, store_arnoldi=True)
    elif name == "z_dense_gmres_MlMr":
        N = 40
        c["A"] = np.diag(np.linspace(1, 10, N)) + 0.3 * crandn(N, N)
        c["b"] = crandn(N)
        c["ls"] = dict(Ml=np.diag(1.0 / np.linspace(1, 10, N)) * (1 + 0.2j),
                       Mr=np.diag(np.linspace(0.5, 1.5, N)))
        c["solver"] = "gmres"
        c["kw"] = dict(tol=1e-12, maxiter=40, x0=crandn(N, 1), store_arnoldi=True)
    else:
        raise KeyError(name)


ALL_CASES = [
    "c1_gmres", "c1_cg", "c1_minres", "c1_gmres_defl", "c1_cg_defl", "c1_minres_defl",
    "c1_gmres_store", "c1_cg_store", "c1_minres_store",
    "lap2d_gmres30", "lap2d_gmres_mgs", "lap2d_gmres_dmgs", "poisson3d_cg_jacobi",
    "convdiff_defl_gmres", "shifted_minres_ipB", "shifted_minres_ipB_f64",
    "dense_gmres_M_ipB", "dense_minres_M_ipB", "dense_cg_M_ipB", "dense_gmres_MlMr",
    "lucky_breakdown", "zero_rhs", "maxiter_noconv", "cg_explicit_residual",
    "lap2d_defl_cg", "lap2d_defl_minres",
]

# complex128 systems (device path: real embedding + twin storage, krypy_b200/_cplx.py)
COMPLEX_CASES = [
    "z_gmres_helmholtz", "z_gmres_helmholtz_dmgs", "z_restarted_gmres", "z_gmres_realA_complexb",
    "z_cg_hpd", "z_minres_herm", "z_minres_herm_mgs", "z_defl_cg", "z_defl_minres",
    "z_defl_gmres", "z_defl_gmres_ipB", "z_dense_gmres_M_ipB", "z_dense_minres_M_ipB",
    "z_dense_cg_M_ipB", "z_dense_gmres_MlMr",
]


class NearestRitzValuesFactory(object):
    """Deflation-vector factory (duck-typed like krypy.recycling.factories._DeflationVectorFactory)
    that selects the Ritz vectors whose Ritz values are nearest to the given targets.  Works with
    the reference package and with krypy_b200 (``kp`` = the package)."""

    def __init__(self, kp, targets):
        self.kp = kp
        self.targets = targets

    def get(self, solver):
        import numpy
        ritz = self.kp.deflation.Ritz(solver, mode="ritz")
        vals = numpy.asarray(ritz.values)
        idx = [int(numpy.argmin(numpy.abs(vals - t))) for t in self.targets]
        return ritz.get_vectors(idx)
