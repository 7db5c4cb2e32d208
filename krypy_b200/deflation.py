"""Host-side mirror of the deflation projector and mixin of ``krypy.deflation``
(krypy/deflation.py:19-283).

The projector ``P = I - AU <U,AU>^{-1} <U, .>`` is applied right after the
operator inside the Arnoldi/Lanczos step by ONE fused cooperative kernel
(kry_project: block dots, small QR solve, block update, twice); the columns of
``C = <U, MlAMr V_n>`` are left in HBM by that kernel and only converted to the
reference's layout when the attribute is read.
"""
import numpy

from . import _cplx, _device, linsys, utils
from .utils import _ctx, _is_dev

__all__ = ["DeflatedCg", "DeflatedMinres", "DeflatedGmres", "_DeflationMixin",
           "ObliqueProjection", "_Projection", "Ritz", "Arnoldifyer", "bound_pseudo"]


class _Projection(utils.Projection):
    def __init__(self, linear_system, U, **kwargs):
        """Abstract base class of a projection for deflation (deflation.py:19-29)."""
        raise NotImplementedError("abstract base class cannot be instanciated")


class ObliqueProjection(_Projection):
    def __init__(self, linear_system, U, qr_reorthos=0, **kwargs):
        """Oblique projection for left deflation (krypy/deflation.py:32-56)."""
        ctx = _ctx()
        self.linear_system = ls = linear_system
        if isinstance(U, utils.DeviceBlock):
            Ud = U.block.to(ls._td)                # vector-major (d, N) block already in HBM
            d = Ud.shape[0]
        else:
            if not _is_dev(U):
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
        c = self._small_correct(c).reshape(-1)
        out = ctx.empty(zd.shape, zd.dtype)
        utils._combine(ctx, self._Wd, self._k, c, zd[0], out[0])         # z + W c
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
        if not _is_dev(U) and not isinstance(U, utils.DeviceBlock):
            U = numpy.asarray(U)
            if U.ndim == 1:
                U = U.reshape(-1, 1)
        udt = _device.torch_to_np_dtype(U.dtype) if _is_dev(U) else U.dtype
        if numpy.dtype(udt).kind == "c" and numpy.dtype(linear_system.dtype).kind != "c":
            # complex deflation vectors (e.g. Ritz vectors of a real nonsymmetric problem):
            # the projector and the whole solve are complex, deflation.py:123-125
            linear_system = linear_system._as_dtype(numpy.complex128)
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
        udtype = _device.torch_to_np_dtype(U.dtype) if _is_dev(U) else U.dtype    # DeviceBlock.dtype is numpy
        if numpy.dtype(udtype).kind not in "fc":
            udtype = numpy.float64
        super(_DeflationMixin, self).__init__(linear_system, dtype=udtype, *args, **kwargs)

    # -- C = <U, MlAMr V_n> ----------------------------------------------------
    def _raw_columns(self):
        """host (d, ncols) array of the raw first-application coefficients."""
        if self._d == 0 or self._ncols == 0:
            return numpy.zeros((self._d, self._ncols))
        raw = self._Craw[: self._ncols].cpu().numpy()                # (ncols, d doubles | 2d interleaved)
        if self._td == _device.torch().complex128:
            raw = _cplx.from_pairs(raw)
        raw = raw.T.copy()                                           # (d, ncols)
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
        nr = 2 if self._td == _device.torch().complex128 else 1
        self._Craw = ctx.scalars(max((self.maxiter + 2) * nr * max(self._d, 1), 1)).reshape(
            self.maxiter + 2, nr * max(self._d, 1))
        # (a weak reference: the operator is stored on the solver, and a bound method would make the
        # solver -- with its bases in HBM -- cyclic garbage that only Python's cyclic collector frees)
        import weakref
        me = weakref.ref(self)
        P = utils._FunctionDeviceOperator((N, N), self.linear_system.dtype,
                                          lambda Av: me()._apply_projection(Av))
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
            Vd = self._basis_dev()[:n_]                      # the basis is still in HBM: no PCIe round trip
            AUd = self.projection._AUd
            if ls.self_adjoint:
                self._B_ = self.C.T.conj()
                if n_ > n:
                    last = utils._inner_dev(Vd[n_ - 1:n_], AUd, ls.ip_B).cpu().numpy().copy()
                    self._B_ = numpy.vstack([self._B_, last])
            else:
                self._B_ = utils._inner_dev(Vd, AUd, ls.ip_B).cpu().numpy().copy()
        return self._B_

    def _basis_dev(self):
        """device block of the Arnoldi/Lanczos basis of the finished solve"""
        if hasattr(self, "arnoldi"):
            return self.arnoldi._Vd
        if hasattr(self, "lanczos"):
            return self.lanczos._Vd
        return self._Vd

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


class Ritz(object):
    """Ritz / harmonic Ritz pairs of a deflated solve (krypy/deflation.py:737-869).

    The (n+m)-sized eigenproblem is host algebra on the small matrices the solve left behind
    (``H``, ``B_``, ``C``, ``E``; ``F = <AU, M AU>`` is one device block product); the N-sized part --
    the Ritz vectors ``[V_n, U] coeffs`` and the explicit residuals -- runs on the device.
    ``get_vectors`` returns an ``(N, k)`` numpy array like the reference; ``get_vectors_dev`` keeps
    the block in HBM (``utils.DeviceBlock``) so a recycled deflation space never crosses PCIe."""

    def __init__(self, deflated_solver, mode="ritz"):
        import scipy.linalg
        self._solver = self._deflated_solver = sv = deflated_solver
        self._cplx_bases = None
        ls = sv.linear_system
        self.values = None
        self.coeffs = None
        Hx = numpy.asarray(sv.H)
        (n1, n) = Hx.shape
        Hn = Hx[:n, :n]
        pr = sv.projection
        if not isinstance(pr, ObliqueProjection):
            raise utils.ArgumentError("Invalid projection used in deflated_solver. Valid are ObliqueProjection")
        m = pr._k
        if n + m == 0:                                           # deflation.py:766-770
            self.values = numpy.zeros((0,))
            self.coeffs = numpy.zeros((0,))
            self.resnorms = numpy.zeros((0,))
            return
        E, C = sv.E, sv.C
        EiC = numpy.linalg.solve(E, C) if m > 0 else C          # deflation.py:775-778
        Bx = sv.B_
        Bn = Bx[:n, :]
        eye, zer = numpy.eye, numpy.zeros
        Mmat = numpy.block([[Hn + Bn.dot(EiC), Bn], [C, E]])      # deflation.py:783
        if m > 0:
            MAUd = ls.M._apply_dev(pr._AUd)
            F = utils._inner_dev(pr._AUd, MAUd, ls.ip_B).cpu().numpy().copy()
        else:
            F = zer((0, 0))
        S = numpy.block([[eye(n1), Bx, zer((n1, m))],
                         [Bx.T.conj(), F, E],
                         [zer((m, n1)), E.T.conj(), eye(m)]])    # deflation.py:785-791
        eig = scipy.linalg.eigh if ls.self_adjoint else scipy.linalg.eig
        if mode == "ritz":
            self.values, self.coeffs = eig(Mmat)
        elif mode == "harmonic":
            L = numpy.block([[Hx, zer((n1, m))], [EiC, eye(m)]])
            K = numpy.block([[eye(n1), Bx], [Bx.T.conj(), F]])
            sig, self.coeffs = eig(Mmat.T.conj(), L.T.conj().dot(K.dot(L)))
            self.values = numpy.zeros(m + n, dtype=sig.dtype)
            tiny = numpy.abs(sig) < numpy.finfo(float).eps
            self.values[~tiny] = 1.0 / sig[~tiny]
            self.values[tiny] = numpy.inf
        else:
            raise utils.ArgumentError("Invalid value  '%s' for 'mode'. Valid are ritz and harmonic." % mode)
        self.coeffs = self.coeffs / numpy.linalg.norm(self.coeffs, 2, axis=0, keepdims=True)   # :814-815
        # residual norms of the Ritz pairs from the small matrices (deflation.py:817-834)
        self.resnorms = numpy.zeros(m + n)
        for i in range(n + m):
            mu = self.values[i]
            y = self.coeffs[:, [i]]
            G = numpy.block([[Hx - mu * eye(n1, n), zer((n1, m))],
                             [EiC, eye(m)],
                             [zer((m, n)), -mu * eye(m)]])
            Gy = G.dot(y)
            self.resnorms[i] = numpy.sqrt(numpy.abs((Gy.T.conj().dot(S.dot(Gy)))[0, 0]))

    # -- N-sized parts on the device -------------------------------------------------------
    def _real_coeffs(self, indices, realify):
        co = self.coeffs if indices is None else self.coeffs[:, indices]
        if co.ndim == 1:
            co = co.reshape(-1, 1)
        if numpy.iscomplexobj(co):
            if numpy.abs(co.imag).max() <= 1e-14 * max(numpy.abs(co).max(), 1e-300):
                return numpy.ascontiguousarray(co.real)
            if not realify:
                return None            # complex Ritz vectors of a real problem: the caller goes complex
            # [Re, Im] spans the same space as the selected vectors when conjugate pairs are selected
            # together; a pivoted QR of the coefficients keeps the k most independent combinations
            import scipy.linalg
            k = co.shape[1]
            RI = numpy.hstack([co.real, co.imag])
            _, _, piv = scipy.linalg.qr(RI, mode="economic", pivoting=True)
            return numpy.ascontiguousarray(RI[:, piv[:k]])
        return numpy.ascontiguousarray(co)

    def get_vectors_dev(self, indices=None, realify=False):
        """Ritz vectors as a vector-major device block wrapped in ``utils.DeviceBlock``."""
        sv = self._solver
        ctx = _ctx()
        t = _device.torch()
        (n1, n) = numpy.asarray(sv.H).shape
        pr = sv.projection
        m = pr._k
        Vd = sv._basis_dev()
        N = sv.linear_system.N
        co = None if sv._td == t.complex128 else self._real_coeffs(indices, realify)
        if co is None:
            # complex coefficients over the twin storage of V_n and U (complex system, or complex Ritz
            # vectors of a real nonsymmetric problem, deflation.py:792-796 + SURVEY F10)
            co = self.coeffs if indices is None else self.coeffs[:, indices]
            co = co.reshape(-1, 1) if co.ndim == 1 else co
            k = co.shape[1]
            Ud = pr._Ud
            if sv._td != t.complex128:
                if self._cplx_bases is None:
                    self._cplx_bases = (Vd[:n].to(t.complex128), Ud.to(t.complex128))
                Vd, Ud = self._cplx_bases
            out = ctx.empty((k, N), t.complex128)
            tmp = ctx.empty((1, N), t.complex128)
            for j in range(k):
                utils._combine(ctx, Vd, n, co[:n, j], None, out[j])
                if m:
                    utils._combine(ctx, Ud, m, co[n:, j], out[j], tmp[0])
                    out[j].copy_(tmp[0])
            return utils.DeviceBlock(out)
        k = co.shape[1]
        out = ctx.empty((k, N), sv._td)
        cd = t.from_numpy(numpy.ascontiguousarray(co.T, dtype=numpy.float64)).to(ctx.device)   # (k, n+m)
        for j in range(k):
            ctx.block_combine(Vd, n, cd[j][:n] if n else None, None, out[j])                 # V[:, :n] y
            if m:
                ctx.block_axpy(pr._Ud, m, cd[j][n:], 1.0, out[j])                             # + U z
        return utils.DeviceBlock(out)

    def get_vectors(self, indices=None, realify=False):
        """krypy/deflation.py:840-847."""
        blk = self.get_vectors_dev(indices, realify)
        out = _ctx().to_numpy(blk.block)
        return out if numpy.iscomplexobj(out) else out.astype(self._solver.dtype, copy=False)

    def get_explicit_residual(self, indices=None):
        """krypy/deflation.py:849-855: ``MlAMr Z - Z diag(values)`` as an ``(N, k)`` numpy array like
        every other public accessor (the device block stays internal)."""
        return _ctx().to_numpy(self._explicit_residual_dev(indices))

    def _explicit_residual_dev(self, indices=None):
        """the explicit Ritz residuals as a vector-major (k, N) device block"""
        ctx = _ctx()
        blk = self.get_vectors_dev(indices).block
        vals = self.values if indices is None else self.values[indices]
        vals = numpy.atleast_1d(vals)
        res = self._solver.linear_system.MlAMr._apply_dev(blk)
        if numpy.iscomplexobj(vals) and not utils._is_cplx(blk):
            vals = vals.real          # real vectors come with (numerically) real values
        for j in range(blk.shape[0]):
            utils._caxpby(ctx, -vals[j], blk[j], 1.0, res[j], res[j])
        return res

    def get_explicit_resnorms(self, indices=None):
        """krypy/deflation.py:857-869."""
        ls = self._solver.linear_system
        res = self._explicit_residual_dev(indices)
        out = numpy.zeros(res.shape[0])
        for j in range(res.shape[0]):
            rj = res[j:j + 1]
            out[j] = linsys._norm_dev(rj, ls.M._apply_dev(rj), ls.ip_B)
        return out


from ._arnoldifyer import Arnoldifyer, bound_pseudo  # noqa: E402,F401  (krypy/deflation.py:286-734)
