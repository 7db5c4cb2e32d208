"""``Arnoldifyer`` and ``bound_pseudo`` of ``krypy.deflation`` (krypy/deflation.py:286-734; SURVEY.md
section 8f rank 4): Arnoldi relations of approximate deflated Krylov subspaces from the data a
deflated solve left behind, and the residual bounds built on them (used by the recycling subset
evaluators).

Almost everything is (n+d)-sized host algebra.  The N-sized pieces stay in HBM: the part of
``M A U`` outside ``span[V, U]`` (block combinations of device bases), its rank-revealing
orthonormalisation (a pivoted modified Gram-Schmidt on the device, carried out directly in the
``<.,.>_{M^-1 B}`` inner product), the Arnoldi basis ``Vh`` and the low-rank perturbation ``F``.
"""
import numpy
import scipy.linalg

from . import _device, linsys, utils
from .utils import _ctx


def _pivoted_mgs_dev(ctx, Xd, ip_B, thresh):
    """Rank-revealing QR ``X = Q1 R12`` of a device block (d, N) in the ``ip_B`` inner product:
    column-pivoted modified Gram-Schmidt (two projection sweeps per column); columns whose remaining
    norm is <= thresh end the factorisation.  Returns (Q1 device (l, N), R12 host (l, d) with the
    columns in the ORIGINAL order)."""
    d = Xd.shape[0]
    W = Xd.clone()
    R = numpy.zeros((d, d), dtype=numpy.complex128 if utils._is_cplx(Xd) else numpy.float64)
    perm = list(range(d))
    nrm2 = ctx.scalars(max(d, 1))
    tmp = ctx.scalars(1)
    rank = 0
    for i in range(d):
        for j in range(i, d):
            utils._ip_coef(W[j:j + 1], W[j:j + 1], ip_B, nrm2[j:])
        rest = nrm2[i:d].cpu().numpy()
        p = i + int(numpy.argmax(rest))
        nrm = float(numpy.sqrt(abs(rest[p - i])))
        if not nrm > thresh:
            break
        if p != i:
            keep = W[i].clone()
            W[i].copy_(W[p])
            W[p].copy_(keep)
            perm[i], perm[p] = perm[p], perm[i]
            R[:i, [i, p]] = R[:i, [p, i]]
        R[i, i] = nrm
        tmp.fill_(nrm)
        ctx.scale_dev(tmp, 1, 1.0, W[i], W[i])
        if i + 1 < d:
            for _ in range(2):
                c = utils._inner_dev(W[i:i + 1], W[i + 1:], ip_B).cpu().numpy().reshape(-1)
                for j in range(i + 1, d):
                    utils._caxpby(ctx, -c[j - i - 1], W[i], 1.0, W[j], W[j])
                R[i, i + 1:] += c
        rank += 1
    return W[:rank], R[:rank][:, numpy.argsort(perm)]


class _LowRankPerturbation(utils._DeviceOperator):
    """``F x = -(Z Rh <Vh, x> + Vh Rh^* <Z, x>)`` on device blocks (krypy/deflation.py:455-466)."""

    def __init__(self, Zd, Vhd, Rh, ip_B, dtype):
        N = Vhd.shape[1]
        super(_LowRankPerturbation, self).__init__((N, N), dtype)
        self._Zd, self._Vhd, self._Rh, self._ip = Zd, Vhd, Rh, ip_B

    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        a = self._Rh.dot(utils._inner_dev(self._Vhd, Xd, self._ip).cpu().numpy())
        b = self._Rh.T.conj().dot(utils._inner_dev(self._Zd, Xd, self._ip).cpu().numpy())
        if out is None:
            out = ctx.empty(Xd.shape, Xd.dtype)
        tmp = ctx.empty((1, Xd.shape[1]), Xd.dtype)
        for j in range(Xd.shape[0]):
            utils._combine(ctx, self._Zd, self._Zd.shape[0], -a[:, j], None, tmp[0])
            utils._combine(ctx, self._Vhd, self._Vhd.shape[0], -b[:, j], tmp[0], out[j])
        return out


class Arnoldifyer(object):
    """krypy/deflation.py:286-470."""

    def __init__(self, deflated_solver):
        self._deflated_solver = sv = deflated_solver
        self._cplx_copies = None
        ctx = _ctx()
        t = _device.torch()
        ls, pr = sv.linear_system, sv.projection
        H, Bx, C, E = numpy.asarray(sv.H), sv.B_, sv.C, sv.E
        n1, n = self.n_, self.n = H.shape
        d = self.d = pr._k
        ext = n1 - n                                   # 1, or 0 for an invariant Krylov subspace
        eye, zer = numpy.eye, numpy.zeros
        EinvC = numpy.linalg.solve(E, C) if d > 0 else zer((0, n))
        Bn = Bx[:n, :]
        self.L = numpy.block([[H, zer((n1, d))], [EinvC, eye(d)]])
        self.J = numpy.block([[eye(n, n1), Bn], [zer((d, n1)), E]])
        self.M = numpy.block([[H[:n, :n] + Bn.dot(EinvC), Bn], [C, E]])
        self.A_norm = numpy.linalg.norm(self.M, 2)

        Vd = sv._basis_dev()
        self._Vd, self._Ud = Vd, pr._Ud
        N = ls.N
        if d > 0:
            # the part of M A U outside span[V_{n+1}, U] (device), orthonormalised with pivoting
            MAUd = ls.M._apply_dev(pr._AUd)
            out = ctx.empty((d, N), sv._td)
            tmp = ctx.empty((1, N), sv._td)
            for j in range(d):
                utils._combine(ctx, pr._Ud, d, -E[:, j], MAUd[j], tmp[0])
                utils._combine(ctx, Vd, n1, -Bx[:, j], tmp[0], out[j])
            self._Q1d, self.R12 = _pivoted_mgs_dev(ctx, out, ls.get_ip_Minv_B(), 1e-14 * self.A_norm)
            l = self._Q1d.shape[0]
            tail = numpy.vstack([Bx[n:, :], self.R12])                       # (ext + l, d)
            self.N = numpy.column_stack([eye(l + ext, ext), tail]).dot(
                numpy.block([[zer((d + ext, n)), eye(d + ext)]]))
        else:
            self._Q1d = ctx.empty((0, N), sv._td)
            self.R12 = zer((0, 0))
            self.N = numpy.block([[zer((ext, n)), eye(ext, ext)]])
        # residual basis Z = [v_{n+1}, Q1] (device)
        self._Zd = t.cat([Vd[n:n1].contiguous(), self._Q1d], dim=0) if (ext or self._Q1d.shape[0]) \
            else ctx.empty((0, N), sv._td)

    @property
    def Z(self):
        if self._Zd.shape[0] == 0:
            return numpy.zeros((self._deflated_solver.linear_system.N, 0))
        return _ctx().to_numpy(self._Zd)

    def get(self, Wt, full=False):
        """Arnoldi relation for the deflation space ``W = [V_n, U] Wt`` (krypy/deflation.py:353-470):
        returns ``Hh, Rh, q_norm, vdiff_norm, PWAW_norm[, Vh, F]``."""
        sv = self._deflated_solver
        n, n1, d = self.n, self.n_, self.d
        Wt = numpy.asarray(Wt)
        k = Wt.shape[1]
        if k > 0:
            Qfull, _ = scipy.linalg.qr(Wt)
            Wt, Wperp = Qfull[:, :k], Qfull[:, k:]
        else:
            Wperp = numpy.eye(Wt.shape[0])

        # oblique projector onto range(L Wt)^c along range(J^* Wt)^perp, small dense algebra
        X, Y = self.L.dot(Wt), self.J.T.conj().dot(Wt)
        if k > 0:
            YX = Y.T.conj().dot(X)

            def Pt(a):
                for _ in range(2):                      # one refinement sweep, like utils.Projection
                    a = a - X.dot(numpy.linalg.solve(YX, Y.T.conj().dot(a)))
                return a
        else:
            def Pt(a):
                return a

        rhs = numpy.zeros((n1 + d, 1), dtype=numpy.result_type(self.L.dtype, numpy.float64))
        rhs[0] = sv.MMlr0_norm
        if d > 0:
            rhs = rhs.astype(numpy.result_type(rhs.dtype, sv.UMlr.dtype))
            rhs[n1:] = numpy.linalg.solve(sv.E, sv.UMlr)
        qt = Pt(rhs)
        q = Wperp.T.conj().dot(self.J.dot(qt))

        # rotate the closest vector of [V_n, U] to the first column, then reduce to Hessenberg form
        Q = utils.House(q)
        q_norm = Q.xnorm
        WperpQ = Q.apply(Wperp.T.conj()).T.conj()
        Hh, T = scipy.linalg.hessenberg(
            Q.apply(Wperp.T.conj().dot(self.J).dot(Pt(self.L.dot(WperpQ)))), calc_q=True)
        QT = Q.apply(T)
        Rh = self.N.dot(Pt(self.L.dot(Wperp.dot(QT))))
        vdiff = self.N.dot(qt)
        vdiff_norm = 0 if vdiff.size == 0 else numpy.linalg.norm(vdiff, 2)

        # norm of the projection P_{W^perp, AW}: coefficients of an orthonormal basis of A W in [V, Z]
        if k > 0:
            Ymat = numpy.block([[numpy.eye(n1), sv.B_],
                                [numpy.zeros((d, n1)), sv.E],
                                [numpy.zeros((self.R12.shape[0], n1)), self.R12]])
            AWq, _ = scipy.linalg.qr(Ymat.dot(self.L.dot(Wt)), mode="economic")
            WX = Wt.T.conj().dot(numpy.vstack([AWq[:n, :], AWq[n1:n1 + d, :]]))
            PWAW_norm = 1.0 / numpy.min(scipy.linalg.svdvals(WX))
        else:
            PWAW_norm = 1.0
        if not full:
            return Hh, Rh, q_norm, vdiff_norm, PWAW_norm

        ctx = _ctx()
        t = _device.torch()
        coef = Wperp.dot(QT)                                    # (n+d, n+d-k)
        N = sv.linear_system.N
        td, dtype = sv._td, sv.dtype
        Vb, Ub, Zb = self._Vd[:n1], self._Ud, self._Zd
        if numpy.iscomplexobj(coef) and not utils._is_cplx(Vb):
            if numpy.abs(coef.imag).max() <= 1e-14 * max(numpy.abs(coef).max(), 1e-300):
                coef = coef.real
            else:
                # complex Ritz coefficients of a real (nonsymmetric) problem: complex copies of the bases
                if self._cplx_copies is None:
                    self._cplx_copies = tuple(b.to(t.complex128) for b in (Vb, Ub, Zb))
                Vb, Ub, Zb = self._cplx_copies
                td, dtype = t.complex128, numpy.complex128
        Vhd = ctx.empty((coef.shape[1], N), td)
        tmp = ctx.empty((1, N), td)
        for c in range(coef.shape[1]):
            utils._combine(ctx, Vb, n, coef[:n, c], None, tmp[0])
            if d > 0:
                utils._combine(ctx, Ub, d, coef[n:, c], tmp[0], Vhd[c])
            else:
                Vhd[c].copy_(tmp[0])
        Vh = ctx.to_numpy(Vhd).astype(dtype, copy=False)
        F = _LowRankPerturbation(Zb, Vhd, Rh, sv.linear_system.get_ip_Minv_B(), dtype)
        return Hh, Rh, q_norm, vdiff_norm, PWAW_norm, Vh, F


def bound_pseudo(arnoldifyer, Wt, g_norm=0.0, G_norm=0.0, GW_norm=0.0, WGW_norm=0.0, tol=1e-6,
                 pseudo_type="auto", pseudo_kwargs=None, delta_n=20, terminate_factor=1.0):
    """Bound the residual norms of the next deflated solve (krypy/deflation.py:473-734).  The
    pseudospectral variants for non-self-adjoint problems need the optional ``pseudopy`` package."""
    from scipy.optimize import minimize_scalar
    pseudo_kwargs = pseudo_kwargs or {}
    Hh, Rh, q_norm, vdiff_norm, PWAW_norm = arnoldifyer.get(Wt)
    sv = arnoldifyer._deflated_solver
    ls_orig = sv.linear_system
    if Wt.shape[1] > 0:
        WAW = Wt.T.conj().dot(arnoldifyer.J.dot(arnoldifyer.L.dot(Wt)))
        sigma_min = numpy.min(scipy.linalg.svdvals(WAW))
        if sigma_min <= WGW_norm:
            raise utils.AssumptionError("sigma_min(W^*AW) > ||W^*GW|| not satisfied.")
        eta = GW_norm / (sigma_min - WGW_norm)
    else:
        eta = 0.0
    b_norm = ls_orig.MMlb_norm
    beta = PWAW_norm * (eta * (b_norm + g_norm) + g_norm) + vdiff_norm
    if g_norm >= b_norm:
        raise utils.AssumptionError("||g_norm|| < ||b_norm|| not satisfied")

    # residual norms of the small system Hh z = e_1 q_norm
    Solver = type(sv)
    self_adjoint, normal = ls_orig.self_adjoint, ls_orig.normal
    if issubclass(Solver, (linsys.Minres, linsys.Gmres)):
        aresnorms = utils.get_residual_norms(Hh, self_adjoint=self_adjoint)
    else:
        small = linsys.LinearSystem(Hh, numpy.eye(Hh.shape[0], 1) * q_norm, normal=normal,
                                    self_adjoint=self_adjoint, positive_definite=ls_orig.positive_definite)
        try:
            s = Solver(small, tol=tol, maxiter=Hh.shape[0])
        except utils.ConvergenceError as e:
            s = e.solver
        aresnorms = numpy.array(s.resnorms)
    aresnorms = aresnorms * q_norm
    if pseudo_type == "omit":
        return aresnorms / (b_norm - g_norm)

    evals, evecs = scipy.linalg.eig(Hh)
    if self_adjoint:
        evals = numpy.real(evals)
    Hh_norm = numpy.linalg.norm(Hh, 2)
    if pseudo_type == "auto":
        if numpy.linalg.norm(Hh - Hh.T.conj(), 2) < 1e-14 * Hh_norm:
            pseudo_type = "hermitian"
        elif numpy.linalg.cond(evecs, 2) < 1 + 1e-14:
            pseudo_type = "normal"
        else:
            pseudo_type = "nonnormal"
    if pseudo_type == "contain":
        raise NotImplementedError("contain not yet implemented")
    # perturbations larger than the spectrum put zero into the pseudospectrum; still useful early on
    delta_max = 1e2 * numpy.max(numpy.abs(evals))
    pert0 = PWAW_norm * (eta * (Hh_norm + G_norm) + G_norm)
    delta_min = pert0 + numpy.max(scipy.linalg.svdvals(Rh[:, :1])) if Rh[:, :1].size else pert0
    if delta_min == 0:
        delta_min = 1e-16
    pseudo = None
    if not normal:
        import pseudopy
        pseudo = pseudopy.NonnormalAuto(Hh, delta_min * 0.99, delta_max * 1.01, **pseudo_kwargs)
    elif not self_adjoint:
        import pseudopy
        pseudo = pseudopy.NormalEvals(evals)

    bounds = [aresnorms[0]]
    for i in range(1, len(aresnorms)):
        # roots of the residual polynomial of step i
        if issubclass(Solver, linsys.Cg):
            roots = scipy.linalg.eigvalsh(Hh[:i, :i])
        else:
            Qi, Ri = scipy.linalg.qr(Hh[: i + 1, :i], mode="economic")
            inv = scipy.linalg.eigvals(Qi[:i, :].T.conj(), Ri)
            roots = 1.0 / inv[numpy.abs(inv) > 1e-14]
        if self_adjoint:
            roots = numpy.real(roots)
        p = utils.NormalizedRootsPolynomial(roots)
        extrema = p.minmax_candidates() if self_adjoint else None
        Rh_i = Rh[:, :i]
        epsilon = pert0 + (numpy.max(scipy.linalg.svdvals(Rh_i)) if Rh_i.size else 0.0)
        if epsilon == 0:
            epsilon = 1e-16
        if epsilon >= delta_max:
            break
        lo, hi = numpy.log10(1.01 * epsilon), numpy.log10(delta_max)
        grid = numpy.linspace(lo, hi, delta_n + 2)[:-1]

        def value(delta_log):
            delta = 10 ** delta_log
            if self_adjoint:
                ivs = utils.Intervals([utils.Interval(ev - delta, ev + delta) for ev in evals])
                inside = [c for c in extrema if ivs.contains(c)]
                pts = numpy.hstack([ivs.get_endpoints(), numpy.array(inside)])
                polymax = numpy.max(numpy.abs(p(pts)))
                length = 2 * delta
            else:
                path = pseudo.contour_paths(delta)
                length = path.length()
                polymax = numpy.max(numpy.abs(p(path.vertices()))) if length > 0 else numpy.inf
            return (length / (2 * numpy.pi * delta)
                    * (epsilon / (delta - epsilon) * (q_norm + beta) + beta) * polymax)

        opt = minimize_scalar(value, bounds=(grid[0], grid[-1]), method="bounded",
                              options={"maxiter": delta_n})
        boundval = aresnorms[i] + opt.fun
        if i > 1 and boundval / bounds[-1] > terminate_factor:
            break
        bounds.append(numpy.min([boundval, bounds[-1]]))
    return numpy.array(bounds) / (b_norm - g_norm)
