"""Host-side mirror of the deflation projector and mixin of ``krypy.deflation``
(krypy/deflation.py:19-283).

The projector ``P = I - AU <U,AU>^{-1} <U, .>`` is applied right after the
operator inside the Arnoldi/Lanczos step by ONE fused cooperative kernel
(kry_project: block dots, small QR solve, block update, twice); the columns of
``C = <U, MlAMr V_n>`` are left in HBM by that kernel and only converted to the
reference's layout when the attribute is read.
"""
import numpy

from . import _device, linsys, utils
from .utils import _ctx, _is_dev

__all__ = ["DeflatedCg", "DeflatedMinres", "DeflatedGmres", "_DeflationMixin",
           "ObliqueProjection", "_Projection"]


class _Projection(utils.Projection):
    def __init__(self, linear_system, U, **kwargs):
        """Abstract base class of a projection for deflation (deflation.py:19-29)."""
        raise NotImplementedError("abstract base class cannot be instanciated")


class ObliqueProjection(_Projection):
    def __init__(self, linear_system, U, qr_reorthos=0, **kwargs):
        """Oblique projection for left deflation (krypy/deflation.py:32-56)."""
        ctx = _ctx()
        self.linear_system = ls = linear_system
        if _is_dev(U):
            d = U.shape[0]
            Ud = U.to(ls._td)
        else:
            U = numpy.asarray(U)
            (N, d) = U.shape
            Ud = ctx.to_block(U, ls._td)
        # orthogonalize U in the Minv-inner-product (deflation.py:40)
        if d > 0:
            Ud, _ = utils._qr_dev(Ud, ip_B=ls.get_ip_Minv_B(), reorthos=qr_reorthos)
        self._Ud = Ud
        # apply operator to U (deflation.py:47)
        self._AUd = ls.MlAMr._apply_dev(Ud) if d > 0 else Ud
        self._MAU = None
        super(_Projection, self).__init__(self._AUd, self._Ud, ip_B=ls.ip_B, **kwargs)

    @property
    def U(self):
        """Orthonormalised basis of the deflation space, ``(N, d)`` numpy."""
        return _ctx().to_numpy(self._Ud).astype(self.linear_system.dtype, copy=False)

    @property
    def AU(self):
        """``MlAMr U`` as ``(N, d)`` numpy."""
        return _ctx().to_numpy(self._AUd).astype(self.linear_system.dtype, copy=False)

    def _small_correct(self, c):
        """c -> WR VR^{-1} R^{-1} Q^H c  (d x d host algebra, deflation.py:64-67)."""
        import scipy.linalg
        if self.Q is not None and self.R is not None:
            c = scipy.linalg.solve_triangular(self.R, self.Q.T.conj().dot(c))
        if self.WR is not self.VR:
            c = self.WR.dot(scipy.linalg.solve_triangular(self.VR, c))
        return c

    def _correct_dev(self, zd):
        """krypy/deflation.py:58-68 on a device block (1, N)."""
        ctx = _ctx()
        ls = self.linear_system
        if self._k == 0:
            return zd
        Az = ls.A._apply_dev(zd)
        r = ctx.empty(zd.shape, zd.dtype)
        ctx.axpby(1.0, ls._b_dev[0], -1.0, Az[0], r[0])
        c = ls.Ml._apply_dev(r)
        c = utils._inner_dev(self._Wd, c, self.ip_B).cpu().numpy()       # (d, 1); synchronises
        c = numpy.ascontiguousarray(self._small_correct(c).reshape(-1), dtype=numpy.float64)
        cd = _device.torch().from_numpy(c).to(ctx.device)
        out = ctx.empty(zd.shape, zd.dtype)
        ctx.block_combine(self._Wd, self._k, cd, zd[0], out[0])          # z + W c
        return out

    def correct(self, z):
        """Correct the approximate solution ``z`` (numpy ``(N,1)``) w.r.t. the
        deflation space (krypy/deflation.py:58-68)."""
        ctx = _ctx()
        zd = z if _is_dev(z) else ctx.to_block(numpy.asarray(z), self.linear_system._td)
        out = self._correct_dev(zd)
        return out if _is_dev(z) else ctx.to_numpy(out)

    @property
    def MAU(self):
        """``M MlAMr U`` (krypy/deflation.py:70-76)."""
        if self._MAU is None:
            self._MAU = _ctx().to_numpy(self.linear_system.M._apply_dev(self._AUd))
        return self._MAU


class _DeflationMixin(object):
    """Mixin class for deflation in Krylov subspace methods
    (krypy/deflation.py:79-233)."""

    def __init__(self, linear_system, U=None, projection_kwargs=None, *args, **kwargs):
        if U is None:
            U = numpy.zeros((linear_system.N, 0))
        if projection_kwargs is None:
            projection_kwargs = {}
        if not _is_dev(U):
            U = numpy.asarray(U)
            if U.ndim == 1:
                U = U.reshape(-1, 1)
        projection = ObliqueProjection(linear_system, U, **projection_kwargs)
        self.projection = projection
        d = projection._k
        self._d = d

        # E = ip_B(U, AU) from the projection (deflation.py:104-111)
        if projection.Q is None and projection.R is None:
            E = numpy.eye(d)
        else:
            E = projection.Q.dot(projection.R)
        if projection.VR is not None and projection.WR is not None:
            E = projection.WR.T.conj().dot(E.dot(projection.VR))
        self.E = E

        self._ncols = 0        # number of projector applications inside _solve
        self._Craw = None      # device (maxcols, d): raw W^H (MlAMr v) of each application
        self._C_cache = None
        self._B_ = None
        udtype = _device.torch_to_np_dtype(U.dtype) if _is_dev(U) else U.dtype
        if numpy.dtype(udtype).kind not in "fc":
            udtype = numpy.float64
        super(_DeflationMixin, self).__init__(linear_system, dtype=udtype, *args, **kwargs)

    # -- C = <U, MlAMr V_n> ----------------------------------------------------
    def _raw_columns(self):
        """host (d, ncols) array of the raw first-application coefficients."""
        if self._d == 0 or self._ncols == 0:
            return numpy.zeros((self._d, self._ncols))
        raw = self._Craw[: self._ncols].cpu().numpy().T.copy()       # (d, ncols)
        WR = self.projection.WR
        return WR.T.conj().dot(raw) if WR is not None else raw       # Ya = WR^H c, utils.py:544-545

    @property
    def C(self):
        r""":math:`C=\langle U,M_lAM_rV_n\rangle` (deflation.py:114-119, 142)."""
        if self._C_cache is None or self._C_cache.shape[1] != self._ncols:
            self._C_cache = self._raw_columns()
        return self._C_cache

    def _solve(self):
        """krypy/deflation.py:127-133."""
        N = self.linear_system.N
        ctx = _ctx()
        self._Craw = ctx.scalars(max((self.maxiter + 2) * max(self._d, 1), 1)).reshape(
            self.maxiter + 2, max(self._d, 1))
        P = utils._FunctionDeviceOperator((N, N), self.linear_system.dtype, self._apply_projection)
        self.MlAMr = P * self.linear_system.MlAMr
        super(_DeflationMixin, self)._solve()

    def _apply_projection(self, Av):
        """Apply the projection to the device block ``Av`` in place and leave
        ``<U, Av>`` in HBM (krypy/deflation.py:135-143)."""
        if self._d == 0:
            self._ncols += 1
            return Av
        j = self._ncols
        if j >= self._Craw.shape[0]:
            raise utils.RuntimeError("more projector applications than maxiter+2")
        self.projection._complement_dev(Av, c_first=self._Craw[j], out=Av)
        self._ncols += 1
        return Av

    def _discard_speculative(self):
        # the look-ahead step of Gmres applied the projector once more than the reference would
        if self._ncols > 0:
            self._ncols -= 1

    def _get_initial_residual(self, x0):
        """Projected initial residual M P Ml (b - A x0) (krypy/deflation.py:145-159)."""
        ls = self.linear_system
        ctx = _ctx()
        if x0 is None:
            Mlr = ls._Mlb_dev
        else:
            Ax = ls.A._apply_dev(x0)
            r = ctx.empty(x0.shape, x0.dtype)
            ctx.axpby(1.0, ls._b_dev[0], -1.0, Ax[0], r[0])
            Mlr = ls.Ml._apply_dev(r)
        PMlr, self.UMlr = self.projection._complement_dev(Mlr, return_Ya=True) \
            if self._d > 0 else (Mlr.clone(), numpy.zeros((0, 1)))
        MPMlr = ls.M._apply_dev(PMlr)
        MPMlr_norm = linsys._norm_dev(PMlr, MPMlr, ls.ip_B)
        return MPMlr, PMlr, MPMlr_norm

    def _get_xk(self, yk):
        """krypy/deflation.py:161-163."""
        xk = super(_DeflationMixin, self)._get_xk(yk)
        return self.projection._correct_dev(xk)

    @property
    def B_(self):
        r""":math:`\underline{B}=\langle V_{n+1},M_lAM_rU\rangle` (deflation.py:165-189)."""
        (n_, n) = self.H.shape
        ls = self.linear_system
        if self._B_ is None or self._B_.shape[1] < n_:
            if ls.self_adjoint:
                self._B_ = self.C.T.conj()
                if n_ > n:
                    self._B_ = numpy.vstack(
                        [self._B_, utils.inner(self.V[:, [-1]], self.projection.AU, ip_B=ls.ip_B)])
            else:
                self._B_ = utils.inner(self.V, self.projection.AU, ip_B=ls.ip_B)
        return self._B_

    def estimate_time(self, nsteps, ndefl, deflweight=1.0):
        """krypy/deflation.py:191-233."""
        solver_ops = self.operations(nsteps)
        proj_ops = {
            "A": ndefl, "M": ndefl, "Ml": ndefl, "Mr": ndefl,
            "ip_B": (ndefl * (ndefl + 1) / 2 + ndefl ** 2 + 2 * ndefl * solver_ops["Ml"]),
            "axpy": (ndefl * (ndefl + 1) / 2 + ndefl * ndefl + (2 * ndefl + 2) * solver_ops["Ml"]),
        }
        if not isinstance(self.linear_system, linsys.TimedLinearSystem):
            raise utils.RuntimeError("A `TimedLinearSystem` has to be used in order to obtain timings.")
        timings = self.linear_system.timings
        return timings.get_ops(solver_ops) + deflweight * timings.get_ops(proj_ops)


class DeflatedCg(_DeflationMixin, linsys.Cg):
    """Deflated preconditioned CG (krypy/deflation.py:236-263)."""

    def __init__(self, *args, **kwargs):
        self._rho_snap = []
        super(DeflatedCg, self).__init__(*args, **kwargs)

    def _apply_projection(self, Av):
        # remember the rhos the reference's recurrence would see at this call
        # (deflation.py:253-260); C itself is assembled lazily from the raw columns
        self._rho_snap.append((self.iter, tuple(self.rhos[-3:])))
        return super(DeflatedCg, self)._apply_projection(Av)

    @property
    def C(self):
        """Three-term recurrence of krypy/deflation.py:247-263, evaluated on the
        host from the per-application coefficients kept in HBM."""
        if self._C_cache is not None and self._C_cache.shape[1] == self._ncols:
            return self._C_cache
        UAps = self._raw_columns()
        C = numpy.zeros((self._d, 0))
        for j, (it, rh) in enumerate(self._rho_snap[: self._ncols]):
            c = UAps[:, [j]].copy()
            if it > 0:
                c -= (1 + rh[-1] / rh[-2]) * UAps[:, [j - 1]]
            if it > 1:
                c += rh[-2] / rh[-3] * UAps[:, [j - 2]]
            c *= ((-1) ** it) / numpy.sqrt(rh[-1])
            if it > 0:
                c -= numpy.sqrt(rh[-2] / rh[-1]) * C[:, [-1]]
            C = numpy.column_stack([C, c])
        self._C_cache = C
        return C


class DeflatedMinres(_DeflationMixin, linsys.Minres):
    """Deflated preconditioned MINRES (krypy/deflation.py:266-273)."""


class DeflatedGmres(_DeflationMixin, linsys.Gmres):
    """Deflated preconditioned GMRES (krypy/deflation.py:276-283)."""
