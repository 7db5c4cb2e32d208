"""CPU oracle for the KryPy Krylov hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the algorithm of the reference
(andrenarchy/krypy v2.2.0).  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
``krypy_b200`` never imports it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this file
against (i) the literal known-answer numbers of the reference's
``test/test_convenience_wrappers.py:10-12, 37-39`` and (ii) fixtures under
``tests/golden/`` produced by importing the real reference in the build
container (``oracle/make_golden.py``).

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The arithmetic is kept in the reference's order and the
reference's data layout (C-ordered ``(N, maxiter+1)`` basis addressed with
``V[:, [j]]`` column gathers, one numpy call per vector operation) so that
(a) histories agree with the reference to round-off and (b) its run time is a
fair stand-in for the reference's CPU path when it is timed as ``cpu_baseline``.
"""
import numpy as np
import scipy.linalg
import scipy.linalg.blas as _blas
from scipy.sparse import issparse

__all__ = [
    "OracleConvergenceError", "inner", "norm", "givens", "qr_ip", "apply_op",
    "System", "ArnoldiState", "Projector", "gmres", "restarted_gmres", "cg",
    "minres", "csr_matvec",
]


class OracleConvergenceError(Exception):
    """krypy/utils.py:81-91 -- carries the populated result."""

    def __init__(self, msg, result):
        super().__init__(msg)
        self.result = result


# --------------------------------------------------------------------------
# operators, inner products, norms
# --------------------------------------------------------------------------
def apply_op(op, X):
    """krypy/utils.py:1381-1390, 1565-1566, 1593-1594: ``op * X``.

    ``None`` is the identity (returns X itself, as IdentityLinearOperator._dot
    does), matrices use ``.dot``, anything else is called.
    """
    if op is None:
        return X
    if isinstance(op, np.ndarray) or issparse(op):
        return op.dot(X)
    return op(X)


def compose(*ops):
    """krypy/utils.py:1407-1414, 1497-1498: product with identities elided."""
    ops = [o for o in ops if o is not None]
    if not ops:
        return None
    if len(ops) == 1:
        return ops[0]

    def _prod(X):
        for o in reversed(ops):
            X = apply_op(o, X)
        return X

    return _prod


def inner(X, Y, B=None):
    """krypy/utils.py:160-193.  B: None | matrix/operator | 2-arg callable
    tagged with ``B.is_ip_callable = True``."""
    if B is None:
        return np.dot(X.T.conj(), Y)
    if getattr(B, "is_ip_callable", False):
        return B(X, Y)
    m = X.shape[1]
    n = Y.shape[1]
    if m > n:
        return np.dot(apply_op(B, X).T.conj(), Y)
    return np.dot(X.T.conj(), apply_op(B, Y))


class InnerProductError(Exception):
    pass


def norm(x, y=None, B=None):
    """krypy/utils.py:214-238."""
    if y is None and B is None:
        return np.linalg.norm(x, 2)
    if y is None:
        y = x
    ip = inner(x, y, B)
    nrm_diag = np.linalg.norm(np.diag(ip), 2)
    nrm_diag_imag = np.linalg.norm(np.imag(np.diag(ip)), 2)
    if nrm_diag_imag > nrm_diag * 1e-10:
        raise InnerProductError("inner product not positive definite?")
    return np.sqrt(np.linalg.norm(ip, 2))


def givens(x):
    """krypy/utils.py:405-436.  x has shape (2,1).  Returns (c, s, r, G)."""
    a = x[0].item()
    b = x[1].item()
    if np.isreal(x).all():
        a = np.real(a)
        b = np.real(b)
        c, s = _blas.drotg(a, b)
    else:
        c, s = _blas.zrotg(a, b)
    r = c * a + s * b
    G = np.array([[c, s], [-np.conj(s), c]])
    return c, s, r, G


def qr_ip(X, B=None, reorthos=1, lapack_shortcut=False):
    """krypy/utils.py:680-707.

    ``lapack_shortcut`` is the reference's ``ip_B is None`` branch (:692-693).
    Inside LinearSystem-driven code the inner product is an
    IdentityLinearOperator, never ``None`` (linsys.py:86-89), so the MGS loop
    (:695-707) runs even for the Euclidean inner product; callers here say
    which of the two the reference would take.
    """
    if lapack_shortcut and B is None and X.shape[1] > 0:
        return scipy.linalg.qr(X, mode="economic")
    (N, k) = X.shape
    Q = X.copy()
    R = np.zeros((k, k), dtype=X.dtype)
    for i in range(k):
        for _ in range(reorthos + 1):
            for j in range(i):
                alpha = inner(Q[:, [j]], Q[:, [i]], B)[0, 0]
                R[j, i] += alpha
                Q[:, [i]] -= alpha * Q[:, [j]]
        R[i, i] = norm(Q[:, [i]], B=B)
        if R[i, i] >= 1e-15:
            Q[:, [i]] /= R[i, i]
    return Q, R


def csr_matvec(n_row, Ap, Aj, Ax, x):
    """scipy/sparse/sparsetools/csr.h ``csr_matvec`` (scipy >= 0.17, the
    third-party kernel behind krypy/utils.py:1593-1594): per row, a sequential
    left-to-right ``sum += Ax[jj] * x[Aj[jj]]`` starting from 0, products and
    sums rounded separately.  Pure-numpy restatement using a segmented
    sequential accumulation (exact same order of additions)."""
    y = np.zeros(n_row, dtype=np.result_type(Ax, x))
    lens = np.diff(Ap)
    maxlen = int(lens.max()) if n_row else 0
    prod = Ax * x[Aj]
    start = Ap[:-1]
    for t in range(maxlen):
        sel = lens > t
        y[sel] = y[sel] + prod[start[sel] + t]
    return y


# --------------------------------------------------------------------------
# LinearSystem
# --------------------------------------------------------------------------
class System(object):
    """krypy/linsys.py:11-176 (LinearSystem) reduced to what the path needs."""

    def __init__(self, A, b, M=None, Minv=None, Ml=None, Mr=None, B=None,
                 exact_solution=None, dtype=None):
        self.N = len(b)
        self.A, self.M, self.Minv, self.Ml, self.Mr, self.B = A, M, Minv, Ml, Mr, B
        self.MlAMr = compose(Ml, A, Mr)                      # linsys.py:85
        self.b = b.reshape(self.N, 1) if b.ndim == 1 else b  # linsys.py:92-94
        self.exact_solution = exact_solution
        if exact_solution is not None and exact_solution.ndim == 1:
            self.exact_solution = exact_solution.reshape(self.N, 1)
        # dtype promotion, linsys.py:115-117 with utils.py:106-122; identity
        # operators contribute float64 (utils.py:1559-1563)
        dts = [np.dtype(np.float64), self.b.dtype]
        for o in (A, M, Ml, Mr, B):
            if hasattr(o, "dtype"):
                dts.append(o.dtype)
        if dtype is not None:
            dts.append(np.dtype(dtype))
        self.dtype = np.result_type(*dts)
        self.Mlb = apply_op(Ml, self.b)                      # linsys.py:120
        self.MMlb = apply_op(M, self.Mlb)                    # linsys.py:121
        self.MMlb_norm = norm(self.Mlb, self.MMlb, B)        # linsys.py:122

    def get_residual(self, z, compute_norm=False):
        """krypy/linsys.py:130-161."""
        if z is None:
            if compute_norm:
                return self.MMlb, self.Mlb, self.MMlb_norm
            return self.MMlb, self.Mlb
        r = self.b - apply_op(self.A, z)
        Mlr = apply_op(self.Ml, r)
        MMlr = apply_op(self.M, Mlr)
        if compute_norm:
            return MMlr, Mlr, norm(Mlr, MMlr, self.B)
        return MMlr, Mlr

    def ip_Minv_B(self):
        """krypy/linsys.py:163-176."""
        if self.M is not None:
            if self.Minv is None:
                raise ValueError("Minv has to be provided")
            if getattr(self.B, "is_ip_callable", False):
                f = lambda x, y: self.B(x, apply_op(self.Minv, y))
                f.is_ip_callable = True
                return f
            return compose(self.Minv, self.B)
        return self.B


# --------------------------------------------------------------------------
# Arnoldi / Lanczos
# --------------------------------------------------------------------------
class ArnoldiState(object):
    """krypy/utils.py:854-1074 for ortho in {mgs, dmgs, lanczos}."""

    def __init__(self, A, v, maxiter, ortho="mgs", M=None, Mv=None,
                 Mv_norm=None, B=None, dtype=None):
        N = v.shape[0]
        self.A, self.M, self.B, self.ortho = A, M, B, ortho
        self.maxiter = N if maxiter is None else maxiter
        dts = [np.dtype(np.float64), v.dtype]
        for o in (A, M):
            if hasattr(o, "dtype"):
                dts.append(o.dtype)
        if dtype is not None:
            dts.append(np.dtype(dtype))
        self.dtype = np.result_type(*dts)                    # utils.py:898
        self.iter = 0
        self.V = np.zeros((N, self.maxiter + 1), dtype=self.dtype)   # :902
        if M is not None:
            self.P = np.zeros((N, self.maxiter + 1), dtype=self.dtype)
        self.H = np.zeros((self.maxiter + 1, self.maxiter), dtype=self.dtype)
        self.invariant = False
        if ortho not in ("mgs", "dmgs", "lanczos"):
            raise ValueError("oracle covers mgs, dmgs, lanczos")
        self.reorthos = 1 if ortho == "dmgs" else 0          # :924-926
        if M is not None:                                    # :927-938
            p = v
            v = apply_op(M, p) if Mv is None else Mv
            self.vnorm = norm(p, v, B) if Mv_norm is None else Mv_norm
            if self.vnorm > 0:
                self.P[:, [0]] = p / self.vnorm
        else:                                                # :939-943
            self.vnorm = norm(v, B=B) if Mv_norm is None else Mv_norm
        if self.vnorm > 0:                                   # :949-952
            self.V[:, [0]] = v / self.vnorm
        else:
            self.invariant = True

    def advance(self):
        """krypy/utils.py:954-1048 (non-Householder branch)."""
        k = self.iter
        Av = apply_op(self.A, self.V[:, [k]])                # :968
        start = 0
        if self.ortho == "lanczos":                          # :1000-1009
            start = k
            if k > 0:
                self.H[k - 1, k] = self.H[k, k - 1]
                if self.M is not None:
                    Av -= self.H[k, k - 1] * self.P[:, [k - 1]]
                else:
                    Av -= self.H[k, k - 1] * self.V[:, [k - 1]]
        for _ in range(self.reorthos + 1):                   # :1012-1029
            for j in range(start, k + 1):
                alpha = inner(self.V[:, [j]], Av, self.B)[0, 0]
                if self.ortho == "lanczos":
                    alpha = alpha.real
                self.H[j, k] += alpha
                if self.M is not None:
                    Av -= alpha * self.P[:, [j]]
                else:
                    Av -= alpha * self.V[:, [j]]
        if self.M is not None:                               # :1030-1034
            MAv = apply_op(self.M, Av)
            self.H[k + 1, k] = norm(Av, MAv, self.B)
        else:
            self.H[k + 1, k] = norm(Av, B=self.B)
        if self.H[k + 1, k] / np.linalg.norm(self.H[: k + 2, : k + 1], 2) <= 1e-14:
            self.invariant = True                            # :1035-1039
        else:                                                # :1041-1045
            if self.M is not None:
                self.P[:, [k + 1]] = Av / self.H[k + 1, k]
                self.V[:, [k + 1]] = MAv / self.H[k + 1, k]
            else:
                self.V[:, [k + 1]] = Av / self.H[k + 1, k]
        self.iter += 1

    def get(self):
        """krypy/utils.py:1050-1061."""
        k = self.iter
        if self.invariant:
            out = (self.V[:, :k], self.H[:k, :k])
            return out + ((self.P[:, :k],) if self.M is not None else ())
        out = (self.V[:, : k + 1], self.H[: k + 1, :k])
        return out + ((self.P[:, : k + 1],) if self.M is not None else ())


# --------------------------------------------------------------------------
# Projection (deflation)
# --------------------------------------------------------------------------
class Projector(object):
    """krypy/utils.py:439-627 (XQRY oblique projection) and
    krypy/deflation.py:32-76 (ObliqueProjection)."""

    def __init__(self, system, U, qr_reorthos=0, iterations=2):
        self.sys = system
        self.B = B = system.B
        self.iterations = iterations
        U, _ = qr_ip(U, system.ip_Minv_B(), reorthos=qr_reorthos)   # deflation.py:40
        self.U = U
        self.AU = apply_op(system.MlAMr, U) if U.shape[1] > 0 else np.zeros(U.shape)  # :47
        X, Y = self.AU, self.U
        if X.shape[1] == 0:                                  # utils.py:498-501
            self.V = self.W = np.zeros(X.shape)
            self.VR = self.WR = self.Q = self.R = None
            return
        self.V, self.VR = qr_ip(X, B)                        # utils.py:505
        self.W, self.WR = qr_ip(Y, B)                        # utils.py:515
        Mx = inner(self.W, self.V, B)                        # utils.py:519
        self.Q, self.R = scipy.linalg.qr(Mx)                 # utils.py:520

    def _apply(self, a, return_Ya=False):
        """krypy/utils.py:522-552."""
        c = inner(self.W, a, self.B)
        if return_Ya:
            Ya = c.copy()
            if self.WR is not None:
                Ya = self.WR.T.conj().dot(Ya)
        if self.Q is not None and self.R is not None:
            c = scipy.linalg.solve_triangular(self.R, self.Q.T.conj().dot(c))
        Pa = self.V.dot(c)
        if return_Ya:
            return Pa, Ya
        return Pa

    def apply_complement(self, a, return_Ya=False):
        """krypy/utils.py:604-627."""
        if self.V.shape[1] == 0:
            if return_Ya:
                return a.copy(), np.zeros((0, a.shape[1]))
            return a.copy()
        if return_Ya:
            x, Ya = self._apply(a, True)
        else:
            x = self._apply(a)
        z = a - x
        for _ in range(self.iterations - 1):
            w = self._apply(z)
            z = z - w
        if return_Ya:
            return z, Ya
        return z

    def correct(self, z):
        """krypy/deflation.py:58-68."""
        if self.V.shape[1] == 0:
            # W has zero columns: c is (0,1); z + W.dot(c) == z + 0
            return z + np.zeros(z.shape)
        s = self.sys
        c = apply_op(s.Ml, s.b - apply_op(s.A, z))
        c = inner(self.W, c, self.B)
        if self.Q is not None and self.R is not None:
            c = scipy.linalg.solve_triangular(self.R, self.Q.T.conj().dot(c))
        if self.WR is not self.VR:
            c = self.WR.dot(scipy.linalg.solve_triangular(self.VR, c))
        return z + self.W.dot(c)

    def E(self):
        """krypy/deflation.py:105-111."""
        d = self.U.shape[1]
        if self.Q is None and self.R is None:
            E = np.eye(d)
        else:
            E = self.Q.dot(self.R)
        if self.VR is not None and self.WR is not None:
            E = self.WR.T.conj().dot(E.dot(self.VR))
        return E


# --------------------------------------------------------------------------
# solver skeleton
# --------------------------------------------------------------------------
class _Run(object):
    """State shared by the three solvers: krypy/linsys.py:277-493 plus the
    deflation hooks of krypy/deflation.py:93-163."""

    def __init__(self, system, x0, tol, maxiter, explicit_residual, U, dtype,
                 projection_kwargs=None):
        s = self.sys = system
        N = s.N
        self.maxiter = N if maxiter is None else maxiter
        if x0 is not None and x0.ndim == 1:
            x0 = x0.reshape(N, 1)
        self.explicit_residual = explicit_residual
        self.tol = tol
        self.proj = None
        udtype = None
        if U is not None:                                    # deflation.py:93-125
            self.proj = Projector(s, U, **(projection_kwargs or {}))
            self.E = self.proj.E()
            self.C = np.zeros((U.shape[1], 0))
            udtype = U.dtype
        # initial residual: linsys.py:359 / deflation.py:145-159
        if self.proj is None:
            self.MMlr0, self.Mlr0, self.MMlr0_norm = s.get_residual(x0, True)
        else:
            if x0 is None:
                Mlr = s.Mlb
            else:
                Mlr = apply_op(s.Ml, s.b - apply_op(s.A, x0))
            PMlr, self.UMlr = self.proj.apply_complement(Mlr, return_Ya=True)
            MPMlr = apply_op(s.M, PMlr)
            self.MMlr0, self.Mlr0 = MPMlr, PMlr
            self.MMlr0_norm = norm(PMlr, MPMlr, s.B)
        self.x0 = np.zeros((N, 1)) if x0 is None else x0     # linsys.py:362-363
        dts = [s.dtype, self.x0.dtype]
        for d in (dtype, udtype):
            if d is not None:
                dts.append(np.dtype(d))
        self.dtype = np.result_type(*dts)                    # linsys.py:370-372
        self.xk = None
        self.iter = 0
        self.resnorms = []
        if s.MMlb_norm == 0:                                 # linsys.py:385-387
            self.xk = self.x0 = np.zeros((N, 1))
            self.resnorms.append(0.0)
        else:
            self.resnorms.append(self.MMlr0_norm / s.MMlb_norm)
        self.errnorms = None
        if s.exact_solution is not None:                     # linsys.py:393-402
            self.errnorms = [norm(s.exact_solution - self.get_xk(None), B=s.B)]
        # operator with deflation hook: deflation.py:127-143
        if self.proj is None:
            self.op = s.MlAMr
        else:
            self.op = lambda X: self.apply_projection(apply_op(s.MlAMr, X))

    def apply_projection(self, Av):
        """krypy/deflation.py:135-143."""
        PAv, UAv = self.proj.apply_complement(Av, return_Ya=True)
        self.C = np.column_stack([self.C, UAv])
        return PAv

    def base_xk(self, yk):
        """krypy/linsys.py:423-428 (overridden by GMRES)."""
        if yk is not None:
            return self.x0 + apply_op(self.sys.Mr, yk)
        return self.x0

    def get_xk(self, yk):
        """krypy/deflation.py:161-163 on top of the solver's own _get_xk."""
        xk = self.base_xk(yk)
        if self.proj is not None:
            return self.proj.correct(xk)
        return xk

    def finalize_iteration(self, yk, resnorm):
        """krypy/linsys.py:430-493."""
        s = self.sys
        self.xk = None
        if s.exact_solution is not None:
            self.xk = self.get_xk(yk)
            self.errnorms.append(norm(s.exact_solution - self.xk, B=s.B))
        rkn = None
        if (self.explicit_residual or resnorm / s.MMlb_norm <= self.tol
                or self.iter + 1 == self.maxiter):
            if self.xk is None:
                self.xk = self.get_xk(yk)
            _, _, rkn = s.get_residual(self.xk, compute_norm=True)
            self.resnorms.append(rkn / s.MMlb_norm)
            if self.resnorms[-1] > self.tol and self.iter + 1 == self.maxiter:
                self.finalize()
                raise OracleConvergenceError(
                    "No convergence in last iteration (maxiter: %d, residual: %s)."
                    % (self.maxiter, self.resnorms[-1]), self)
        else:
            self.resnorms.append(resnorm / s.MMlb_norm)
        return rkn

    def finalize(self):
        pass


class _GmresRun(_Run):
    """krypy/linsys.py:877-1006."""

    def __init__(self, system, x0=None, tol=1e-5, maxiter=None,
                 explicit_residual=False, ortho="mgs", U=None, dtype=None,
                 projection_kwargs=None):
        self.ortho = ortho
        self.arnoldi = None
        super().__init__(system, x0, tol, maxiter, explicit_residual, U, dtype,
                         projection_kwargs)
        self.solve()
        self.finalize()

    def base_xk(self, y):
        """krypy/linsys.py:941-949."""
        if y is None:
            return self.x0
        k = self.arnoldi.iter
        if k > 0:
            yy = scipy.linalg.solve_triangular(self.R[:k, :k], y)
            yk = self.V[:, :k].dot(yy)
            return self.x0 + apply_op(self.sys.Mr, yk)
        return self.x0

    def solve(self):
        """krypy/linsys.py:951-997."""
        s = self.sys
        self.arnoldi = ar = ArnoldiState(
            self.op, self.Mlr0, maxiter=self.maxiter, ortho=self.ortho, M=s.M,
            Mv=self.MMlr0, Mv_norm=self.MMlr0_norm, B=s.B, dtype=self.dtype)
        G = []
        self.R = np.zeros([self.maxiter + 1, self.maxiter], dtype=self.dtype)
        y = np.zeros((self.maxiter + 1, 1), dtype=self.dtype)
        y[0] = self.MMlr0_norm
        self.V = ar.V
        while (self.resnorms[-1] > self.tol and ar.iter < ar.maxiter
               and not ar.invariant):
            k = self.iter = ar.iter
            ar.advance()
            self.V = ar.V
            self.R[: k + 2, k] = ar.H[: k + 2, k]
            for i in range(k):
                self.R[i: i + 2, k] = G[i].dot(self.R[i: i + 2, k])
            G.append(givens(self.R[k: k + 2, [k]])[3])
            self.R[k: k + 2, k] = G[k].dot(self.R[k: k + 2, k])
            y[k: k + 2] = G[k].dot(y[k: k + 2])
            self.finalize_iteration(y[: k + 1], abs(y[k + 1, 0]))
        if self.xk is None:
            self.xk = self.get_xk(y[: ar.iter])
        self.y = y

    def finalize(self):
        got = self.arnoldi.get()
        self.Vk, self.H = got[0], got[1]
        self.Pk = got[2] if len(got) > 2 else None


class _MinresRun(_Run):
    """krypy/linsys.py:711-862."""

    def __init__(self, system, x0=None, tol=1e-5, maxiter=None,
                 explicit_residual=False, ortho="lanczos", U=None, dtype=None,
                 projection_kwargs=None):
        self.ortho = ortho
        super().__init__(system, x0, tol, maxiter, explicit_residual, U, dtype,
                         projection_kwargs)
        self.solve()
        self.finalize()

    def solve(self):
        """krypy/linsys.py:791-853."""
        s = self.sys
        N = s.N
        self.lanczos = lz = ArnoldiState(
            self.op, self.Mlr0, maxiter=self.maxiter, ortho=self.ortho, M=s.M,
            Mv=self.MMlr0, Mv_norm=self.MMlr0_norm, B=s.B, dtype=self.dtype)
        W = np.column_stack([np.zeros(N, dtype=self.dtype), np.zeros(N)])
        y = [self.MMlr0_norm, 0]
        G2 = None
        G1 = None
        yk = np.zeros((N, 1), dtype=self.dtype)
        while (self.resnorms[-1] > self.tol and lz.iter < lz.maxiter
               and not lz.invariant):
            k = self.iter = lz.iter
            lz.advance()
            V, H = lz.V, lz.H
            R = np.zeros((4, 1))
            R[1] = H[k - 1, k].real
            if G1 is not None:
                R[:2] = G1.dot(R[:2])
            R[2:4, 0] = [H[k, k].real, H[k + 1, k].real]
            if G2 is not None:
                R[1:3] = G2.dot(R[1:3])
            G1 = G2
            c_, s_, r_, G2 = givens(R[2:4])
            R[2] = r_
            R[3] = 0.0
            y = G2.dot(y)
            z = (V[:, [k]] - R[0, 0] * W[:, [0]] - R[1, 0] * W[:, [1]]) / R[2, 0]
            W = np.column_stack([W[:, [1]], z])
            yk = yk + y[0] * z
            y = [y[1], 0]
            self.finalize_iteration(yk, np.abs(y[0]))
        if self.xk is None:
            self.xk = self.get_xk(yk)

    def finalize(self):
        got = self.lanczos.get()
        self.Vk, self.H = got[0], got[1]
        self.Pk = got[2] if len(got) > 2 else None


class _CgRun(_Run):
    """krypy/linsys.py:520-696 and krypy/deflation.py:236-263."""

    def __init__(self, system, x0=None, tol=1e-5, maxiter=None,
                 explicit_residual=False, U=None, dtype=None,
                 store_arnoldi=False, projection_kwargs=None):
        self._UAps = []
        self.store_arnoldi = store_arnoldi
        super().__init__(system, x0, tol, maxiter, explicit_residual, U, dtype,
                         projection_kwargs)
        self.solve()
        self.finalize()

    def apply_projection(self, Av):
        """krypy/deflation.py:247-263."""
        PAv, UAp = self.proj.apply_complement(Av, return_Ya=True)
        self._UAps.append(UAp)
        c = UAp.copy()
        rhos = self.rhos
        if self.iter > 0:
            c -= (1 + rhos[-1] / rhos[-2]) * self._UAps[-2]
        if self.iter > 1:
            c += rhos[-2] / rhos[-3] * self._UAps[-3]
        c *= ((-1) ** self.iter) / np.sqrt(rhos[-1])
        if self.iter > 0:
            c -= np.sqrt(rhos[-2] / rhos[-1]) * self.C[:, [-1]]
        self.C = np.column_stack([self.C, c])
        return PAv

    def solve(self):
        """krypy/linsys.py:593-689."""
        s = self.sys
        N = s.N
        yk = np.zeros((N, 1), dtype=self.dtype)
        self.rhos = rhos = [self.MMlr0_norm ** 2]
        self.Mlrk = self.Mlr0.copy()
        self.MMlrk = self.MMlr0.copy()
        p = self.MMlrk.copy()
        self.iter = 0
        if self.store_arnoldi:
            self.V = np.zeros((N, self.maxiter + 1), dtype=self.dtype)
            if self.MMlr0_norm > 0:
                self.V[:, [0]] = self.MMlr0 / self.MMlr0_norm
            if s.M is not None:
                self.P = np.zeros((N, self.maxiter + 1), dtype=self.dtype)
                if self.MMlr0_norm > 0:
                    self.P[:, [0]] = self.Mlr0 / self.MMlr0_norm
            self.H = np.zeros((self.maxiter + 1, self.maxiter))
            alpha_old = 0
        while self.resnorms[-1] > self.tol and self.iter < self.maxiter:
            k = self.iter
            if k > 0:
                p = self.MMlrk + rhos[-1] / rhos[-2] * p
                if self.store_arnoldi:
                    omega = rhos[-1] / rhos[-2]
            Ap = apply_op(self.op, p)
            alpha = rhos[-1] / inner(p, Ap, s.B)[0, 0]
            alpha = alpha.real
            if self.store_arnoldi:
                if k > 0:
                    self.H[k - 1, k] = self.H[k, k - 1]
                    self.H[k, k] = (1.0 + alpha * omega / alpha_old) / alpha
                else:
                    self.H[k, k] = 1.0 / alpha
            yk += alpha * p
            self.Mlrk -= alpha * Ap
            self.MMlrk = apply_op(s.M, self.Mlrk)
            MMlrk_norm = norm(self.Mlrk, self.MMlrk, s.B)
            rhos.append(MMlrk_norm ** 2)
            if self.store_arnoldi:
                self.V[:, [k + 1]] = (-1) ** (k + 1) * self.MMlrk / MMlrk_norm
                if s.M is not None:
                    self.P[:, [k + 1]] = (-1) ** (k + 1) * self.Mlrk / MMlrk_norm
                self.H[k + 1, k] = np.sqrt(rhos[-1] / rhos[-2]) / alpha
                alpha_old = alpha
            rkn = self.finalize_iteration(yk, MMlrk_norm)
            if rkn is not None:
                rhos[-1] = rkn ** 2
            self.iter += 1
        if self.xk is None:
            self.xk = self.get_xk(yk)

    def finalize(self):
        if self.store_arnoldi:
            self.Vk = self.V[:, : self.iter + 1]
            self.H = self.H[: self.iter + 1, : self.iter]


def gmres(system, **kw):
    """krypy.linsys.Gmres / krypy.deflation.DeflatedGmres (pass ``U=``)."""
    return _GmresRun(system, **kw)


def minres(system, **kw):
    """krypy.linsys.Minres / krypy.deflation.DeflatedMinres."""
    return _MinresRun(system, **kw)


def cg(system, **kw):
    """krypy.linsys.Cg / krypy.deflation.DeflatedCg."""
    return _CgRun(system, **kw)


class _Restarted(object):
    pass


def restarted_gmres(system, max_restarts=0, **kw):
    """krypy/linsys.py:1021-1081."""
    out = _Restarted()
    out.xk = None
    out.resnorms = [np.inf]
    if system.exact_solution is not None:
        out.errnorms = [np.inf]
    kw = dict(kw)
    tol = None
    restart = 0
    while restart == 0 or (out.resnorms[-1] > tol and restart <= max_restarts):
        try:
            if out.xk is not None:
                kw.update({"x0": out.xk})
            sol = _GmresRun(system, **kw)
        except OracleConvergenceError as e:
            sol = e.result
        out.xk = sol.xk
        tol = sol.tol
        del out.resnorms[-1]
        out.resnorms += sol.resnorms
        if system.exact_solution is not None:
            del out.errnorms[-1]
            out.errnorms += sol.errnorms
        restart += 1
    out.tol = tol
    if out.resnorms[-1] > tol:
        raise OracleConvergenceError(
            "No convergence after %d restarts." % max_restarts, out)
    return out
