"""Host-side analysis helpers of ``krypy.utils``: principal angles,
the Hegedues rescaling, Ritz pairs of an Arnoldi relation, the projected Arnoldi relation, spectral
gaps / interval sets, the CG and MINRES convergence bounds, root-normalised polynomials.

Everything here is small-matrix algebra (sizes n = number of iterations) except ``angles`` and
``hegedus``, whose N-sized products go through the device functions of ``krypy_b200.utils``
(``qr``, ``inner``, operator applies); there is no N-sized numpy arithmetic on vectors that live
in HBM.  Re-exported from ``krypy_b200.utils`` under the reference's names.
"""
import warnings

import numpy
import scipy.linalg


def _u():
    from . import utils
    return utils


# --------------------------------------------------------------------------------------------
# subspace angles, Hegedues trick
# --------------------------------------------------------------------------------------------
def angles(F, G, ip_B=None, compute_vectors=False):
    """Principal angles (ascending, ``max(k,l)`` of them) between span(F) and span(G) in the
    ``ip_B`` inner product, optionally with principal vectors (krypy/utils.py:710-809).

    Knyazev/Argentati 2002, Alg. 6.2: cosines from the SVD of ``<Q_F, Q_G>`` for the large angles
    (cos^2 < 1/2), sines from the part of ``Q_G`` outside span(F) for the small ones."""
    u = _u()
    swapped = F.shape[1] < G.shape[1]
    if swapped:
        F, G = G, F
    k, l = F.shape[1], G.shape[1]
    QF, _ = u.qr(F, ip_B=ip_B)
    QG, _ = u.qr(G, ip_B=ip_B)
    right = numpy.full(k - l, numpy.pi / 2)
    if l == 0:
        theta, PF, PG = numpy.full(k, numpy.pi / 2), QF, QG
    else:
        Y, cosines, Zh = scipy.linalg.svd(u.inner(QF, QG, ip_B=ip_B))
        G_cos = QG.dot(Zh.T.conj())
        n_small = int(numpy.count_nonzero(cosines ** 2 >= 0.5))      # leading singular values: small angles
        theta = numpy.concatenate([numpy.arccos(cosines[n_small:]), right])
        PF = PG = None
        if compute_vectors:
            F_cos = QF.dot(Y)
            PF, PG = F_cos[:, n_small:], G_cos[:, n_small:]
        if n_small > 0:
            Gs = G_cos[:, :n_small]
            outside = Gs - QF.dot(u.inner(QF, Gs, ip_B=ip_B))        # (I - P_F) of the near-parallel part
            _, R = u.qr(outside, ip_B=ip_B)
            _, sines, Zh2 = scipy.linalg.svd(R)
            theta = numpy.concatenate([numpy.arcsin(sines[::-1][:n_small]), theta])
            if compute_vectors:
                c = cosines[:n_small]
                rot = (Zh2.T.conj() * c[None, :]) / c[:, None]     # diag(1/c) Z diag(c), utils.py:791-799
                PF = numpy.column_stack([F_cos[:, :n_small].dot(rot), PF])
                PG = numpy.column_stack([Gs.dot(Zh2.T.conj()), PG])
    if not compute_vectors:
        return theta
    return (theta, PG, PF) if swapped else (theta, PF, PG)


def hegedus(A, b, x0, M=None, Ml=None, ip_B=None):
    """Scale the initial guess by the gamma that minimises ``||M Ml (b - A gamma x0)||_{M^-1}``
    (krypy/utils.py:812-851): gamma = <z, Ml b> / <z, Ml A x0> with z = M Ml A x0."""
    u = _u()
    N = len(b)
    A = u.get_linearoperator((N, N), A)
    M = u.get_linearoperator((N, N), M)
    Ml = u.get_linearoperator((N, N), Ml)
    w = Ml * (A * x0)
    z = M * w
    denom = u.inner(z, w, ip_B=ip_B)
    if denom <= 1e-15:
        return numpy.zeros((N, 1))
    return (u.inner(z, Ml * b, ip_B=ip_B) / denom) * x0


def norm_MMlr(M, Ml, A, Mr, b, x0, yk, inner_product=None):
    """x_k = x0 + Mr yk and its preconditioned residual, normalising before M is applied
    (krypy/utils.py:276-294)."""
    u = _u()
    xk = x0 + Mr * yk
    Mlr = Ml * (b - A * xk)
    nrm = u.norm(Mlr)
    if nrm == 0:
        return xk, Mlr, numpy.zeros(Mlr.shape), 0
    MMlr = (M * (Mlr / nrm)) * nrm
    ip = u.ip_euclid if inner_product is None else inner_product
    return xk, Mlr, MMlr, numpy.sqrt(numpy.linalg.norm(ip(Mlr, MMlr), 2))


# --------------------------------------------------------------------------------------------
# Ritz pairs / projected Arnoldi relation / residual norms from H
# --------------------------------------------------------------------------------------------
def ritz(H, V=None, hermitian=False, type="ritz"):
    """(Harmonic, improved harmonic) Ritz pairs of an Arnoldi relation (krypy/utils.py:1171-1286).
    Returns ``theta, U, resnorm[, Z = V[:, :n] U]``."""
    u = _u()
    H = numpy.asarray(H)
    n = H.shape[1]
    if V is not None and V.shape[1] != H.shape[0]:
        raise u.ArgumentError("shape mismatch with V and H")
    if H.shape[0] not in (n, n + 1):
        raise u.ArgumentError("H not of shape (n+1,n) or (n,n)")
    Hn = H[:n, :]
    if hermitian:
        asym = numpy.linalg.norm(Hn - Hn.T.conj())
        if asym >= 5e-14:
            warnings.warn("Hessenberg matrix is not symmetric: |H-H^*|=%s" % asym)
    eig = scipy.linalg.eigh if hermitian else scipy.linalg.eig

    def residuals(theta, U):
        R = H.dot(U).astype(numpy.result_type(H.dtype, U.dtype, theta.dtype))
        R[:n, :] -= U * theta[None, :]
        return numpy.linalg.norm(R, 2, axis=0)

    if type == "ritz":
        theta, U = eig(Hn)
        last = 0 if H.shape[0] == n else H[-1, -1]
        resnorm = numpy.abs(last * U[-1, :])
    elif type in ("harmonic", "harmonic_improved"):
        mu, U = eig(Hn.T.conj(), H.T.conj().dot(H))
        U = U / numpy.linalg.norm(U, 2, axis=0, keepdims=True)
        if type == "harmonic":
            theta = 1 / mu
        else:
            theta = numpy.array([U[:, i].conj().dot(Hn.dot(U[:, i])) for i in range(n)])
        resnorm = residuals(theta, U)
    else:
        raise u.ArgumentError("unknown Ritz type %s" % type)
    if V is not None:
        return theta, U, resnorm, numpy.dot(V[:, :n], U)
    return theta, U, resnorm


def arnoldi_projected(H, P, k, ortho="mgs"):
    """Arnoldi relation of the projected operator from the data of an existing one, without further
    applications of A (krypy/utils.py:1084-1168).  Returns ``U, G, F``."""
    u = _u()
    H = numpy.asarray(H)
    n = H.shape[1]
    dtype = u.find_common_dtype(H, P)
    square = H.shape[0] == n
    hlast = 0 if square else H[-1, -1]
    Hop = u.get_linearoperator((n, n), H if square else H[:-1, :])
    Pop = u.get_linearoperator((n, n), P)
    steps = n - k + 1
    F = numpy.zeros((1, steps), dtype=dtype)
    PH = u.LinearOperator((n, n), dtype, lambda x: Pop * (Hop * x))
    ar = u.Arnoldi(PH, Pop * numpy.eye(n, 1), maxiter=steps, ortho=ortho)
    while ar.iter < ar.maxiter and not ar.invariant:
        last, _ = ar.get_last()
        F[0, ar.iter] = hlast * last[-1, 0]
        ar.advance()
    U, G = ar.get()
    return U, G, F[[0], : ar.iter]


def get_residual_norms(H, self_adjoint=False):
    """Relative GMRES/MINRES residual norms that belong to a Hessenberg matrix, zero initial guess
    (krypy/utils.py:2103-2125)."""
    u = _u()
    R = numpy.array(H)
    rows, n = R.shape
    y = numpy.eye(rows, 1, dtype=R.dtype)
    out = [1.0]
    for i in range(rows - 1):
        rot = u.Givens(R[i:i + 2, [i]])
        hi = i + 3 if self_adjoint else n
        R[i:i + 2, i:hi] = rot.apply(R[i:i + 2, i:hi])
        y[i:i + 2] = rot.apply(y[i:i + 2])
        out.append(numpy.abs(y[i + 1, 0]))
    if rows == n:
        out.append(0.0)
    return numpy.array(out)


# --------------------------------------------------------------------------------------------
# spectra: gaps, interval sets, bounds
# --------------------------------------------------------------------------------------------
def strakos(n, l_min=0.1, l_max=100, rho=0.9):
    """Strakos' diagonal test matrix (krypy/utils.py:1639-1648)."""
    i = numpy.arange(1, n + 1)
    return numpy.diag(l_min + (i - 1) / (n - 1) * (l_max - l_min) * rho ** (n - i))


def gap(lamda, sigma, mode="individual"):
    """Gap between two sets of reals (krypy/utils.py:1651-1708): ``'individual'`` the smallest
    pairwise distance; ``'interval'`` the distance of sigma from the hull of lamda (``None`` when a
    sigma lies strictly inside)."""
    u = _u()
    lam = numpy.atleast_1d(numpy.array(lamda))
    sig = numpy.atleast_1d(numpy.array(sigma))
    if not (numpy.isreal(lam).all() and numpy.isreal(sig).all()):
        raise u.ArgumentError("complex spectra not yet implemented")
    if mode == "individual":
        return numpy.min(numpy.abs(lam[:, None] - sig[None, :]))
    if mode == "interval":
        lo, hi = lam.min(), lam.max()
        below, above = sig <= lo, sig >= hi
        if not numpy.all(below | above):
            return None
        delta = numpy.inf
        if below.any():
            delta = lo - sig[below].max()
        if above.any():
            delta = min(delta, sig[above].min() - hi)
        return delta


class Interval(object):
    """Closed real interval, possibly a point (krypy/utils.py:1711-1749)."""

    def __init__(self, left, right=None):
        right = left if right is None else right
        if left > right:
            raise _u().ArgumentError("left > right not allowed.")
        self.left, self.right = left, right

    def __and__(self, other):
        lo, hi = max(self.left, other.left), min(self.right, other.right)
        return Interval(lo, hi) if lo <= hi else None

    def __or__(self, other):
        if self & other:
            return Interval(min(self.left, other.left), max(self.right, other.right))
        return None

    def __repr__(self):
        return "[%s,%s]" % (self.left, self.right)

    def contains(self, alpha):
        return self.left <= alpha <= self.right

    def distance(self, other):
        if self & other:
            return 0
        return max(other.left - self.right, self.left - other.right)


class Intervals(object):
    """Set of pairwise disjoint intervals; adding merges (krypy/utils.py:1752-1844)."""

    def __init__(self, intervals=None):
        self.intervals = set()
        for iv in intervals or ():
            self.add(iv)

    def add(self, new):
        touching = {iv for iv in self.intervals if iv & new}
        for iv in touching:
            new = new | iv
        self.intervals -= touching
        self.intervals.add(new)

    def contains(self, alpha):
        return any(iv.contains(alpha) for iv in self.intervals)

    def get_endpoints(self):
        pts = []
        for iv in self.intervals:
            pts += [iv.left] if iv.left == iv.right else [iv.left, iv.right]
        return sorted(pts)

    def __len__(self):
        return len(self.intervals)

    def __iter__(self):
        return iter(self.intervals)

    def __repr__(self):
        return ", ".join(repr(iv) for iv in self.intervals)

    def _empty(self, what):
        # (the reference RETURNS the exception object for an empty set, utils.py:1803-1804)
        return _u().ArgumentError("empty set has no %s." % what)

    def min(self):
        if not self.intervals:
            return self._empty("minimum")
        return numpy.min([iv.left for iv in self.intervals])

    def max(self):
        if not self.intervals:
            return self._empty("maximum")
        return numpy.max([iv.right for iv in self.intervals])

    def min_pos(self):
        if not self.intervals:
            return self._empty("minimum positive value")
        cand = [iv.left for iv in self.intervals if iv.left > 0]
        if self.contains(0) or not cand:
            return None
        return numpy.min(cand)

    def max_neg(self):
        if not self.intervals:
            return self._empty("maximum negative value")
        cand = [iv.right for iv in self.intervals if iv.right < 0]
        if self.contains(0) or not cand:
            return None
        return numpy.max(cand)

    def min_abs(self):
        if not self.intervals:
            return self._empty("minimum absolute value")
        if self.contains(0):
            return 0
        return numpy.min([abs(v) for v in (self.max_neg(), self.min_pos()) if v is not None])

    def max_abs(self):
        if not self.intervals:
            return self._empty("maximum absolute value")
        return numpy.max(numpy.abs([self.max(), self.min()]))


def _real_sorted(evals, what):
    u = _u()
    if len(evals) == 0:
        raise u.AssumptionError("empty spectrum not allowed")
    if not numpy.isreal(evals).all():
        raise u.AssumptionError("non-real eigenvalues not allowed")
    return numpy.sort(numpy.array(evals, dtype=float))


class BoundCG(object):
    """kappa-bound ``2 ((sqrt(k)-1)/(sqrt(k)+1))^n`` of the CG error in the A-norm from the
    (effective) condition number of the given spectrum (krypy/utils.py:1847-1916)."""

    def __init__(self, evals, exclude_zeros=False):
        u = _u()
        if isinstance(evals, Intervals):
            evals = [evals.min(), evals.max()]
            if evals[0] <= 0:
                raise u.AssumptionError("non-positive eigenvalues not allowed with intervals")
        ev = _real_sorted(evals, "cg")
        ev = ev / ev[-1]
        if exclude_zeros is False and not (ev > 1e-15).all():
            raise u.AssumptionError("non-positive eigenvalues not allowed (use exclude_zeros?)")
        assert ev[0] > -1e-15
        root = numpy.sqrt(1 / numpy.min(ev[ev > 1e-15]))
        self.base = (root - 1) / (root + 1)

    def eval_step(self, step):
        return 2 * self.base ** step

    def get_step(self, tol):
        return numpy.log(tol / 2.0) / numpy.log(self.base)


class BoundMinres(object):
    """MINRES residual bound for a spectrum in ``[l_1, l_s] u [l_t, l_N]``, ``l_s < 0 < l_t``
    (krypy/utils.py:1919-2003); a non-negative spectrum yields a :class:`BoundCG` instead."""

    def __new__(cls, evals):
        if isinstance(evals, Intervals):
            nonneg = evals.min() > 0
        else:
            nonneg = bool((numpy.array(evals) > -1e-15).all())
        return BoundCG(evals) if nonneg else super(BoundMinres, cls).__new__(cls)

    def __init__(self, evals):
        u = _u()
        if isinstance(evals, Intervals):
            if evals.contains(0):
                raise u.AssumptionError("zero eigenvalues not allowed with intervals")
            ends = (evals.min(), evals.max_neg(), evals.min_pos(), evals.max())
            evals = [v for v in ends if v is not None]
        ev = _real_sorted(evals, "minres")
        ev = ev / numpy.max(numpy.abs(ev))
        neg, pos = ev[ev < -1e-15], ev[ev > 1e-15]
        outer = numpy.sqrt(numpy.abs(neg.min() * pos.max()))
        inner = numpy.sqrt(numpy.abs(neg.max() * pos.min()))
        self.base = (outer - inner) / (outer + inner)

    def eval_step(self, step):
        return 2 * self.base ** numpy.floor(step / 2.0)

    def get_step(self, tol):
        return 2 * numpy.log(tol / 2.0) / numpy.log(self.base)


def bound_perturbed_gmres(pseudo, p, epsilon, deltas):
    """GMRES bound for a perturbed operator from pseudospectral contours (krypy/utils.py:2006-2033;
    ``pseudo`` is a pseudopy object)."""
    if not numpy.all(numpy.array(deltas) > epsilon):
        raise _u().ArgumentError("all deltas have to be greater than epsilon")
    out = []
    for delta in deltas:
        paths = pseudo.contour_paths(delta)
        sup = numpy.max(numpy.abs(p(paths.vertices())))
        out.append(epsilon / (delta - epsilon) * paths.length() / (2 * numpy.pi * delta) * sup)
    return out


class NormalizedRootsPolynomial(object):
    """``p(x) = prod_i (1 - x/theta_i)`` (krypy/utils.py:2036-2100)."""

    def __init__(self, roots):
        roots = numpy.asarray(roots)
        if roots.ndim != 1:
            raise _u().ArgumentError("one-dimensional array of roots expected.")
        self.roots = roots

    def minmax_candidates(self):
        """roots of p' (the interior extrema candidates)"""
        from numpy.polynomial import Polynomial
        return Polynomial.fromroots(self.roots).deriv(1).roots()

    def __call__(self, points):
        pts = numpy.asarray(points)
        if pts.ndim > 1:
            raise _u().ArgumentError("scalar or one-dimensional array of points expected.")
        n = self.roots.shape[0]
        fac = 1 - pts / self.roots.reshape(n, 1)
        # multiply small and large factors alternately so partial products neither over- nor underflow
        order = numpy.argsort(numpy.abs(fac), axis=0)
        half = (n + 1) // 2
        mix = numpy.empty_like(order)
        mix[0::2] = order[:half]
        mix[1::2] = order[half:][::-1]
        vals = numpy.prod(numpy.take_along_axis(fac, mix, axis=0), axis=0)
        return vals.item() if numpy.isscalar(points) else vals
