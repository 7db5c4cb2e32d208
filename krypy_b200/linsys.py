"""Host-side mirror of ``krypy.linsys``: the
same classes, constructor arguments, attributes and exceptions as the reference;
the iteration itself runs on the B200 through libkrypy_b200.so.

Convergence control stays on the host (one pinned-mailbox read per iteration);
operator applies, Gram-Schmidt, Givens/Hessenberg updates, solution updates and
residual norms are device kernels.
"""
import os
import warnings

import numpy

from . import _cplx, _device, _lib, utils
from .utils import _ctx, _is_dev

__all__ = ["LinearSystem", "Cg", "Minres", "Gmres", "RestartedGmres", "TimedLinearSystem",
           "ConvertedTimedLinearSystem"]


_TRACE = bool(__import__("os").environ.get("KRY_TRACE"))


def _mark(obj, name):
    """KRY_TRACE=1: synchronised wall-clock marks of the solver phases (diagnostics only)"""
    if _TRACE:
        import time
        _device.torch().cuda.synchronize()
        obj.__dict__.setdefault("_trace", []).append((name, time.perf_counter()))


def _host_vec(x):
    """numpy view (N,1) of a public vector argument, or None."""
    if x is None:
        return None
    if _is_dev(x):
        x = x.detach().cpu().numpy()
    x = numpy.asarray(x)
    return x.reshape(x.shape[0], -1)


class _LazyVec(object):
    """Descriptor: numpy (N,1) view of a device block attribute, materialised on access."""

    def __init__(self, name):
        self.dev = "_" + name + "_dev"
        self.cache = "_" + name + "_np"

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        c = obj.__dict__.get(self.cache)
        if c is not None:
            return c
        d = obj.__dict__.get(self.dev)
        if d is None:
            return None
        c = _ctx().to_numpy(d).astype(obj.dtype, copy=False)
        obj.__dict__[self.cache] = c
        return c

    def __set__(self, obj, value):
        if value is None:
            obj.__dict__[self.dev] = None
            obj.__dict__[self.cache] = None
        elif _is_dev(value):
            obj.__dict__[self.dev] = value
            obj.__dict__[self.cache] = None
        else:
            value = numpy.asarray(value)
            obj.__dict__[self.cache] = value
            obj.__dict__[self.dev] = None


class _LazyBasis(object):
    """Descriptor for the ``V`` / ``P`` attributes of a solve with ``store_arnoldi=True``: the basis
    stays in HBM (vector-major) and the reference's ``(N, k)`` numpy array is materialised on first
    access.  The reference builds these arrays eagerly (linsys.py:694-696, 860-862, 1003-1006); at
    N = 4M, k = 61 that is a 2 GB device-to-host copy plus a transpose that the recycling
    work-flow (deflation.Ritz, the factories) never looks at -- it reads the device basis."""

    def __init__(self, name, fallback=None):
        self.name = name
        self.fallback = fallback

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        d = obj.__dict__
        v = d.get("_%s_np" % self.name)
        if v is not None:
            return v
        thunk = d.get("_%s_thunk" % self.name)
        if thunk is not None:
            v = d["_%s_np" % self.name] = thunk()
            d["_%s_thunk" % self.name] = None
            return v
        if self.fallback is not None:
            return self.fallback(obj)
        raise AttributeError(self.name)

    def __set__(self, obj, value):
        if callable(value) and not hasattr(value, "shape"):
            obj.__dict__["_%s_thunk" % self.name] = value
            obj.__dict__["_%s_np" % self.name] = None
        else:
            obj.__dict__["_%s_np" % self.name] = value
            obj.__dict__["_%s_thunk" % self.name] = None


def _store_arnoldi_lazily(solver, ar, with_P):
    """V, H[, P] of ``Arnoldi.get()`` (utils.py:1050-1061) with the N-sized parts deferred"""
    k = ar.iter
    ncols = k if ar.invariant else k + 1
    ctx = _ctx()
    solver.H = ar.H[:k, :k] if ar.invariant else ar.H[: k + 1, :k]
    solver.V = lambda: ctx.to_numpy(ar._Vd[:ncols]).astype(ar.dtype, copy=False)
    if with_P:
        solver.P = lambda: ctx.to_numpy(ar._Pd[:ncols]).astype(ar.dtype, copy=False)
    if getattr(ar, "_ws", None) is not None:
        # restart cycles share their basis buffers through the workspace: the next cycle overwrites
        # them, so this cycle's arrays are taken now
        solver.V
        if with_P:
            solver.P


class LinearSystem(object):
    """krypy/linsys.py:11-201.  ``dtype`` (new, optional): storage/compute dtype
    of the device path; the default follows the reference's promotion rule
    (>= float64); ``dtype=numpy.float32`` opts into fp32 storage."""

    def __init__(self, A, b, M=None, Minv=None, Ml=None, Mr=None, ip_B=None, normal=None,
                 self_adjoint=False, positive_definite=False, exact_solution=None, dtype=None):
        self.N = N = len(b)
        shape = (N, N)
        self.A = utils.get_linearoperator(shape, A)
        self.M = utils.get_linearoperator(shape, M)
        self.Minv = utils.get_linearoperator(shape, Minv)
        self.Ml = utils.get_linearoperator(shape, Ml)
        self.Mr = utils.get_linearoperator(shape, Mr)
        self.MlAMr = self.Ml * self.A * self.Mr
        try:
            self.ip_B = utils.get_linearoperator(shape, ip_B)
        except TypeError:
            self.ip_B = ip_B

        b_dev_in = _is_dev(b)
        if b_dev_in:
            self.flat_vecs = b.dim() == 1
            b_np = None
        else:
            self.flat_vecs, (b_np, exact_solution) = utils.shape_vecs(numpy.asarray(b), exact_solution)
        if _is_dev(exact_solution):
            exact_solution = _host_vec(exact_solution)
        self.exact_solution = exact_solution

        self.self_adjoint = self_adjoint
        if self_adjoint:
            if normal is not None and not normal:
                warnings.warn("Setting normal=True because self_adjoint=True is provided.")
            normal = True
        if normal is None:
            normal = False
        self.normal = normal
        self.positive_definite = positive_definite
        if self_adjoint and not normal:
            raise utils.ArgumentError("self-adjointness implies normality")

        # common dtype (linsys.py:115-117); the identities contribute float64
        bdt = _device.torch_to_np_dtype(b.dtype) if b_dev_in else b_np.dtype
        self.dtype = utils._common_type(
            [utils.find_common_dtype(self.A, self.M, self.Ml, self.Mr, self.ip_B), bdt])
        if dtype is not None:
            self.dtype = numpy.dtype(dtype)
        self._td = utils._compute_dtype(self.dtype)

        ctx = _ctx()
        self._b_dev = ctx.to_block(b if b_dev_in else b_np, self._td)
        self._b_np = b_np
        self._exact_dev = None if exact_solution is None else ctx.to_block(exact_solution, self._td)
        # ||M Ml b||_{M^-1}  (linsys.py:120-122)
        self._Mlb_dev = self.Ml._apply_dev(self._b_dev)
        self._MMlb_dev = self.M._apply_dev(self._Mlb_dev)
        self.MMlb_norm = _norm_dev(self._Mlb_dev, self._MMlb_dev, self.ip_B)

    b = _LazyVec("b")
    Mlb = _LazyVec("Mlb")
    MMlb = _LazyVec("MMlb")

    def _as_dtype(self, npdtype):
        """This system for a solver whose dtype is wider than the system's (complex ``x0`` or complex
        deflation vectors with real ``A, b``: the whole solve becomes complex, linsys.py:370-372,
        deflation.py:123-125; or a float64 ``x0`` with an fp32-storage system).  Returns ``self`` when
        the device dtype is unchanged, else a cached shallow copy that shares the operators (their
        device data is cached per block dtype) and holds promoted copies of the right-hand side."""
        td = utils._compute_dtype(npdtype)
        if td == self._td:
            return self
        views = self.__dict__.setdefault("_dtype_views", {})
        view = views.get(td)
        if view is None:
            if getattr(self, "part", None) is not None:
                raise NotImplementedError("dtype promotion of a row-partitioned system")
            view = object.__new__(type(self))
            view.__dict__.update(self.__dict__)
            view.__dict__.pop("_dtype_views", None)
            view.dtype = numpy.dtype(_device.torch_to_np_dtype(td))
            view._td = td
            for name in ("_b_dev", "_exact_dev", "_Mlb_dev", "_MMlb_dev"):
                blk = self.__dict__.get(name)
                view.__dict__[name] = None if blk is None else blk.to(td)
            views[td] = view
        return view

    # -- device residual -------------------------------------------------------
    def _get_residual_dev(self, zd, compute_norm=False, before_sync=None):
        """(M Ml (b - A z), Ml (b - A z)[, norm]) on device blocks (linsys.py:130-161).
        ``before_sync(MMlr, Mlr, norm_dev)``: see _norm_dev."""
        if zd is None:
            if compute_norm:
                return self._MMlb_dev, self._Mlb_dev, self.MMlb_norm
            return self._MMlb_dev, self._Mlb_dev
        ctx = _ctx()
        Az = self.A._apply_dev(zd)
        r = ctx.empty(zd.shape, zd.dtype)
        ctx.axpby(1.0, self._b_dev[0], -1.0, Az[0], r[0])           # b - A z
        Mlr = self.Ml._apply_dev(r)
        MMlr = self.M._apply_dev(Mlr)
        if compute_norm:
            hook = None if before_sync is None else (lambda nd: before_sync(MMlr, Mlr, nd))
            return MMlr, Mlr, _norm_dev(Mlr, MMlr, self.ip_B, hook)
        return MMlr, Mlr

    def get_residual(self, z, compute_norm=False):
        """krypy/linsys.py:130-161 for a numpy ``(N,1)`` z."""
        if z is None:
            if compute_norm:
                return self.MMlb, self.Mlb, self.MMlb_norm
            return self.MMlb, self.Mlb
        ctx = _ctx()
        zd = z if _is_dev(z) else ctx.to_block(numpy.asarray(z), self._td)
        res = self._get_residual_dev(zd, compute_norm)
        out = (ctx.to_numpy(res[0]), ctx.to_numpy(res[1]))
        return out + ((res[2],) if compute_norm else ())

    def get_ip_Minv_B(self):
        """krypy/linsys.py:163-176."""
        if not isinstance(self.M, utils.IdentityLinearOperator):
            if isinstance(self.Minv, utils.IdentityLinearOperator):
                raise utils.ArgumentError(
                    "Minv has to be provided for the evaluation of the inner "
                    "product that is implicitly defined by M.")
            if isinstance(self.ip_B, utils.LinearOperator):
                return self.Minv * self.ip_B
            else:
                return lambda x, y: self.ip_B(x, self.Minv * y)
        return self.ip_B

    def __repr__(self):
        ret = "LinearSystem {\n"

        def add(k):
            op = getattr(self, k)
            if op is not None and not isinstance(op, utils.IdentityLinearOperator):
                return "  " + k + ": " + op.__repr__() + "\n"
            return ""

        for k in ["A", "b", "M", "Minv", "Ml", "Mr", "ip_B", "normal", "self_adjoint",
                  "positive_definite", "exact_solution"]:
            ret += add(k)
        return ret + "}"


def _norm_dev(xd, yd, ip_B, before_sync=None):
    """sqrt(<x, y>_B) of single-vector device blocks; one host synchronisation.  ``before_sync(norm_dev)``
    runs after the reduction is enqueued and before the host waits for it (work the device result does not
    gate: deferred bookkeeping, the speculative launch of the next restart cycle)."""
    ctx = _ctx()
    tmp = ctx.scalars(1)
    utils._ip_coef(xd, xd if yd is None else yd, ip_B, tmp, post=1)
    if before_sync is not None:
        # the read-back is enqueued BEFORE the hook's work (which may be a whole restart cycle), so the host
        # gets the norm as soon as the reduction is done and runs ahead of the device from here on
        t = _device.torch()
        host = ctx.__dict__.get("_pinned_norm")
        if host is None:
            host = t.zeros(1, dtype=t.float64)
            host = ctx.__dict__["_pinned_norm"] = host.pin_memory() if t.cuda.is_available() else host
        host.copy_(tmp, non_blocking=True)
        ev = ctx.event()
        ev.record()
        before_sync(tmp)
        ev.synchronize()
        return numpy.float64(host[0].item())
    return numpy.float64(tmp[0].item())


class TimedLinearSystem(LinearSystem):
    """A linear system that times every application of its operators and of the inner product
    (``timings['A' | 'M' | 'Minv' | 'Ml' | 'Mr' | 'ip_B']``, per vector; krypy/linsys.py:204-252).  The
    timers are CUDA-event based (utils.Timer): no host synchronisation per application."""

    _TIMED = ("A", "M", "Minv", "Ml", "Mr")

    def __init__(self, A, b, M=None, Minv=None, Ml=None, Mr=None, ip_B=None, normal=None,
                 self_adjoint=False, positive_definite=False, exact_solution=None, dtype=None):
        self.timings = tm = utils.Timings()
        shape = (len(b), len(b))
        ops = dict(A=A, M=M, Minv=Minv, Ml=Ml, Mr=Mr)
        timed = {name: utils.get_linearoperator(shape, ops[name], timer=tm[name]) for name in self._TIMED}
        try:
            timed_ip = utils.get_linearoperator(shape, ip_B, timer=tm["ip_B"])
        except TypeError:
            timer = tm["ip_B"]

            def timed_ip(X, Y):
                # callable inner product: time per entry of the (m, n) result
                entries = X.shape[1] * Y.shape[1]
                if entries == 0:
                    return ip_B(X, Y)
                with timer:
                    G = ip_B(X, Y)
                timer.scale_last(1.0 / entries)
                return G
        super(TimedLinearSystem, self).__init__(
            b=b, ip_B=timed_ip, normal=normal, self_adjoint=self_adjoint, positive_definite=positive_definite,
            exact_solution=exact_solution, dtype=dtype, **timed)


class ConvertedTimedLinearSystem(TimedLinearSystem):
    """krypy/linsys.py:255-274."""

    def __init__(self, linear_system):
        kwargs = {k: getattr(linear_system, k) for k in
                  ["A", "b", "M", "Minv", "Ml", "Mr", "ip_B", "normal", "self_adjoint",
                   "positive_definite", "exact_solution"]}
        super(ConvertedTimedLinearSystem, self).__init__(**kwargs)


class _KrylovSolver(object):
    """krypy/linsys.py:277-517."""

    x0 = _LazyVec("x0")
    xk = _LazyVec("xk")
    MMlr0 = _LazyVec("MMlr0")
    Mlr0 = _LazyVec("Mlr0")

    def __init__(self, linear_system, x0=None, tol=1e-5, maxiter=None, explicit_residual=False,
                 store_arnoldi=False, dtype=None, _x0_residual=None):
        if not isinstance(linear_system, LinearSystem):
            raise utils.ArgumentError("linear_system is not an instance of LinearSystem")
        self.linear_system = ls = linear_system
        self._ctx = ctx = _ctx()
        ctx.use_current_stream()
        _mark(self, "start")
        N = ls.N
        # (row-partitioned runs: N is the local length, the default maxiter the global dimension)
        self.maxiter = getattr(ls, "N_global", N) if maxiter is None else maxiter
        self.explicit_residual = explicit_residual
        self.store_arnoldi = store_arnoldi
        self.tol = tol

        # dtype (linsys.py:370-372); a missing x0 does not promote (zeros in ls.dtype)
        x0dt = None
        if x0 is not None:
            x0dt = _device.torch_to_np_dtype(x0.dtype) if _is_dev(x0) else numpy.asarray(x0).dtype
        self.dtype = utils._common_type([ls.dtype, x0dt, dtype])
        self._td = utils._compute_dtype(self.dtype)
        if self._td != ls._td:
            # e.g. complex x0 with real A, b: the whole solve becomes complex
            self.linear_system = ls = ls._as_dtype(self.dtype)
            self._td = ls._td

        if x0 is None:
            self.flat_vecs = True
            x0d = None
        else:
            self.flat_vecs = (x0.dim() == 1) if _is_dev(x0) else (numpy.asarray(x0).ndim == 1)
            x0d = ctx.to_block(x0, self._td)
            if x0d.shape[1] != N:
                raise utils.ArgumentError("x0 has the wrong length")
        x0d = self._get_initial_guess(x0d)

        # initial residual (linsys.py:359).  A restart hands over the explicit residual the previous
        # cycle computed for exactly this x0 (linsys.py:460 and :359 evaluate the same expression).
        self._last_residual = None
        if _x0_residual is not None and x0d is not None and type(self)._get_initial_residual is \
                _KrylovSolver._get_initial_residual:
            self.MMlr0, self.Mlr0, self.MMlr0_norm = _x0_residual
        else:
            self.MMlr0, self.Mlr0, self.MMlr0_norm = self._get_initial_residual(x0d)
        _mark(self, "initial_residual")
        if x0d is None:
            x0d = ctx.zeros((1, N), self._td)
        self.x0 = x0d
        self.xk = None
        self.MlAMr = ls.MlAMr
        self.iter = 0
        self.resnorms = []
        if ls.MMlb_norm == 0:                                   # linsys.py:385-387
            self.xk = self.x0 = ctx.zeros((1, N), self._td)
            self.resnorms.append(0.0)
        else:
            self.resnorms.append(self.MMlr0_norm / ls.MMlb_norm)
        if ls.exact_solution is not None:                       # linsys.py:393-402
            self.errnorms = []
            self.errnorms.append(self._errnorm(self._get_xk(None)))
        try:
            self._solve()
        finally:
            _mark(self, "solve_end")
            if self.__dict__.get("_ws") is None:
                ctx.l2_window(None)                 # (a restarted solve keeps its window until its last cycle)
        self._finalize()

    # -- hooks -----------------------------------------------------------------
    def _get_initial_guess(self, x0):
        return x0

    def _get_initial_residual(self, x0):
        return self.linear_system._get_residual_dev(x0, compute_norm=True)

    def _errnorm(self, xkd):
        ctx = self._ctx
        ls = self.linear_system
        e = ctx.empty(xkd.shape, xkd.dtype)
        ctx.axpby(1.0, ls._exact_dev[0], -1.0, xkd[0], e[0])
        return _norm_dev(e, None, ls.ip_B)

    def _get_xk(self, yk):
        """krypy/linsys.py:423-428 on device blocks."""
        x0d = self.__dict__["_x0_dev"]
        if yk is not None:
            ctx = self._ctx
            Mry = self.linear_system.Mr._apply_dev(yk)
            out = ctx.empty(x0d.shape, x0d.dtype)
            ctx.axpby(1.0, x0d[0], 1.0, Mry[0], out[0])
            return out
        return x0d

    def _finalize_iteration(self, yk, resnorm):
        """krypy/linsys.py:430-493."""
        ls = self.linear_system
        self.xk = None
        self._last_residual = None
        if ls.exact_solution is not None:
            self.xk = self._get_xk(yk)
            self.errnorms.append(self._errnorm(self.__dict__["_xk_dev"]))
        rkn = None
        if (self.explicit_residual or resnorm / ls.MMlb_norm <= self.tol
                or self.iter + 1 == self.maxiter):
            if self.__dict__.get("_xk_dev") is None:
                self.xk = self._get_xk(yk)
            hook = self.__dict__.pop("_before_residual_sync", None)
            if hook is not None:
                MMlrk, Mlrk, rkn = ls._get_residual_dev(self.__dict__["_xk_dev"], compute_norm=True, before_sync=hook)
            else:
                MMlrk, Mlrk, rkn = ls._get_residual_dev(self.__dict__["_xk_dev"], compute_norm=True)
            self._last_residual = (MMlrk, Mlrk, rkn)
            self.resnorms.append(rkn / ls.MMlb_norm)
            if self.resnorms[-1] > self.tol:
                if self.iter + 1 == self.maxiter:
                    self._finalize()
                    raise utils.ConvergenceError(
                        ("No convergence in last iteration "
                         "(maxiter: %s, residual: %s)." % (self.maxiter, self.resnorms[-1])), self)
                elif not self.explicit_residual and resnorm / ls.MMlb_norm <= self.tol:
                    warnings.warn(
                        "updated residual is below tolerance, explicit residual is NOT! "
                        "(upd=%s <= tol=%s < exp=%s)" % (resnorm, self.tol, self.resnorms[-1]))
        else:
            self.resnorms.append(resnorm / ls.MMlb_norm)
        return rkn

    def _finalize(self):
        pass

    @staticmethod
    def operations(nsteps):
        raise NotImplementedError("operations() has to be overridden by the derived solver class.")

    def _solve(self):
        raise NotImplementedError("_solve has to be overridden by the derived solver class.")

    def _repr(self, name, extra=()):
        s = "krypy %s object\n" % name
        s += "    MMlr0 = [{}, ..., {}]\n".format(self.MMlr0[0], self.MMlr0[-1])
        s += "    MMlr0_norm = {}\n".format(self.MMlr0_norm)
        s += "    MlAMr: {} x {} matrix\n".format(*self.MlAMr.shape)
        s += "    Mlr0: [{}, ..., {}]\n".format(self.Mlr0[0], self.Mlr0[-1])
        for line in extra:
            s += line
        s += "    flat_vecs: {}\n".format(self.flat_vecs)
        s += "    store_arnoldi: {}\n".format(self.store_arnoldi)
        if hasattr(self, "ortho"):
            s += "    ortho: {}\n".format(self.ortho)
        s += "    tol: {}\n".format(self.tol)
        s += "    maxiter: {}\n".format(self.maxiter)
        s += "    iter: {}\n".format(self.iter)
        s += "    explicit residual: {}\n".format(self.explicit_residual)
        s += "    resnorms: [{}, ..., {}]\n".format(self.resnorms[0], self.resnorms[-1])
        s += "    x0: [{}, ..., {}]\n".format(self.x0[0], self.x0[-1])
        s += "    xk: [{}, ..., {}]".format(self.xk[0], self.xk[-1])
        return s


def _diag_of(op):
    if isinstance(op, utils.DiagonalLinearOperator) and numpy.dtype(op._d.dtype).kind != "c":
        return op
    return None


class Cg(_KrylovSolver):
    """Preconditioned CG (krypy/linsys.py:520-708).

    Device iteration (Euclidean inner product, identity or diagonal ``M``):
    ``p = z + beta p`` (kry_axpby), ``Ap = A p`` with ``<p,Ap>`` fused into the
    SpMV epilogue (kry_spmv_csr), and ONE fused sweep for ``x, r, z, rho``
    (kry_cg_update); the host reads ``rho`` from the pinned mailbox."""

    def __init__(self, linear_system, **kwargs):
        if not linear_system.self_adjoint or not linear_system.positive_definite:
            warnings.warn("Cg applied to a non-self-adjoint or non-definite "
                          "linear system. Consider using Minres or Gmres.")
        super(Cg, self).__init__(linear_system, **kwargs)

    def __repr__(self):
        return self._repr("CG")

    Mlrk = _LazyVec("Mlrk")
    MMlrk = _LazyVec("MMlrk")
    V = _LazyBasis("V")
    P = _LazyBasis("P")

    def _apply_op_dot(self, p, Ap, pAp):
        """Ap = MlAMr p and pAp[0] = <p, Ap>_B (linsys.py:631-634)."""
        ctx = self._ctx
        ls = self.linear_system
        op = self.MlAMr
        euclid = utils._is_identity_ip(ls.ip_B)
        if euclid and type(op) is utils.MatrixLinearOperator:
            A = op._dev(p.dtype)
            if isinstance(A, _device.CsrDev) and not getattr(A, "native_z", False):
                ctx.spmv(A, p[0], Ap[0], w=p[0], dot_out=pAp)      # fused <p,Ap> epilogue
                return
        if euclid and hasattr(op, "_apply_dot_dev"):
            op._apply_dot_dev(p, Ap, pAp)     # row-partitioned: halo + SpMV with epilogue + peer all-reduce
            return
        op._apply_dev(p, out=Ap)
        utils._ip_coef(p, Ap, ls.ip_B, pAp)

    def _solve(self):
        """krypy/linsys.py:593-689."""
        ctx = self._ctx
        ls = self.linear_system
        N = ls.N
        td = self._td
        yk = ctx.zeros((1, N), td)
        self.rhos = rhos = [self.MMlr0_norm ** 2]
        Mlr0d, MMlr0d = self.__dict__["_Mlr0_dev"], self.__dict__["_MMlr0_dev"]
        r = Mlr0d.clone()                          # Mlrk
        M_is_id = isinstance(ls.M, utils.IdentityLinearOperator)
        Mdiag = _diag_of(ls.M)
        euclid = utils._is_identity_ip(ls.ip_B)
        # Row-partitioned runs: the fused update publishes the LOCAL rho into the mailbox, which is
        # then summed in place over NVLink (the mailbox is mapped pinned memory: a device address).
        dist = ctx.comm is not None
        fast = euclid and (M_is_id or Mdiag is not None) and (not dist or (
            hasattr(self.MlAMr, "_apply_dot_dev") and ctx.comm.reduce == "peer"))
        z = r if M_is_id else MMlr0d.clone()       # MMlrk (aliases Mlrk when M is the identity)
        dinv = Mdiag._dev(td) if (fast and Mdiag is not None) else None
        def pvec():
            # row-partitioned: peer-mapped, the neighbours read its halo in place
            return self.MlAMr._alloc_vec(td) if dist else ctx.empty((1, N), td)
        if fast:
            pp = (pvec(), pvec())                  # p_k lives in pp[k & 1] (see the look-ahead below)
            pp[0].copy_(MMlr0d)
            p = pp[0]
        else:
            p = MMlr0d.clone()
        Ap = ctx.empty((1, N), td)
        pAp = ctx.scalars(1)
        tmp = ctx.scalars(1)
        self.iter = 0
        store = self.store_arnoldi
        if store:
            ld = (N + 31) // 32 * 32
            self._Vs = ctx.zeros((self.maxiter + 1, ld), td)
            self._Vd = self._Vs[:, :N]
            self._Pd = None
            if self.MMlr0_norm > 0:
                tmp.fill_(float(self.MMlr0_norm))
                ctx.scale_dev(tmp, 1, 1.0, MMlr0d[0], self._Vd[0])
            if not M_is_id:
                self._Ps = ctx.zeros((self.maxiter + 1, ld), td)
                self._Pd = self._Ps[:, :N]
                if self.MMlr0_norm > 0:
                    ctx.scale_dev(tmp, 1, 1.0, Mlr0d[0], self._Pd[0])
            self.H = numpy.zeros((self.maxiter + 1, self.maxiter))
            alpha_old = 0
        mb = ctx.mailbox

        if fast:
            # The scalars of the recurrence (rho, <p,Ap>, alpha, beta) live in device memory, so the FRONT
            # half of iteration k+1 -- p_{k+1} = z + beta p_k and A p_{k+1} with <p,Ap> in the SpMV
            # epilogue, two thirds of an iteration's bytes -- is enqueued BEFORE the host waits for
            # iteration k's residual (look-ahead): the device works through the host's bookkeeping, and
            # on row-partitioned runs no rank's host sits on the critical path of the others.  The
            # speculative front only writes p_{k+1} (the other buffer), A p and <p,Ap>, none of which
            # the solver exposes; x, r, z change in the BACK half, which is launched after the decision.
            st = ctx.scalars(8)                    # [rho_{k-1}, rho_k, <p,Ap>, alpha, beta, local rho share]
            st[1:2].fill_(float(rhos[-1]))
            # (deflated CG keeps its bookkeeping per operator application in step with the iteration
            # counter, deflation.py:247-263: no speculation there)
            lookahead = (not self.explicit_residual and ls.exact_solution is None
                         and not hasattr(self, "projection"))
            ev = ctx.event()

            def front(k):
                if k > 0:
                    ctx.xpby_dev(z[0], st[4:], pp[(k - 1) & 1][0], pp[k & 1][0])      # linsys.py:627
                self._apply_op_dot(pp[k & 1], Ap, st[2:3])                           # linsys.py:631-634

            launched = -1
            while self.resnorms[-1] > self.tol and self.iter < self.maxiter:
                k = self.iter
                if launched < k:
                    front(k)
                    launched = k
                if k > 0 and store:
                    omega = rhos[-1] / rhos[-2]
                ctx.cg_update_dev(Ap[0], pp[k & 1][0], yk[0], r[0], z[0] if dinv is not None else None, dinv, st)
                ctx.cg_scalars(st, 0)                                                # linsys.py:655-665
                ev.record()
                if lookahead and k + 1 < self.maxiter:
                    front(k + 1)
                    launched = k + 1
                ev.synchronize()
                rho_new, alpha = float(mb[0]), float(mb[1])
                if not numpy.isfinite(rho_new) or rho_new < 0:
                    rho_new = abs(rho_new)
                MMlrk_norm = numpy.sqrt(rho_new)
                rhos.append(MMlrk_norm ** 2)                                         # linsys.py:665
                if store:
                    if k > 0:
                        self.H[k - 1, k] = self.H[k, k - 1]
                        self.H[k, k] = (1.0 + alpha * omega / alpha_old) / alpha
                    else:
                        self.H[k, k] = 1.0 / alpha
                    tmp.fill_(float(MMlrk_norm))
                    sgn = (-1.0) ** (k + 1)
                    ctx.scale_dev(tmp, 1, sgn, z[0], self._Vd[k + 1])                # linsys.py:669
                    if self._Pd is not None:
                        ctx.scale_dev(tmp, 1, sgn, r[0], self._Pd[k + 1])            # linsys.py:671
                    self.H[k + 1, k] = numpy.sqrt(rhos[-1] / rhos[-2]) / alpha
                    alpha_old = alpha
                self._Mlrk_dev, self._MMlrk_dev = r, z
                rkn = self._finalize_iteration(yk, MMlrk_norm)                       # linsys.py:678
                if rkn is not None:
                    # the explicit residual replaces rho (linsys.py:681-683): on the device too, and a
                    # front half that was already enqueued with the old beta is redone (p is double
                    # buffered for exactly this)
                    rhos[-1] = rkn ** 2
                    st[1:2].fill_(float(rhos[-1]))
                    st[4:5].fill_(float(rhos[-1] / rhos[-2]))
                    launched = min(launched, k)
                self.iter += 1
            p = pp[(self.iter - 1) & 1] if self.iter > 0 else pp[0]

        while (not fast) and self.resnorms[-1] > self.tol and self.iter < self.maxiter:
            k = self.iter
            if k > 0:
                ctx.axpby(1.0, z[0], rhos[-1] / rhos[-2], p[0], p[0])       # linsys.py:627
                if store:
                    omega = rhos[-1] / rhos[-2]
            self._apply_op_dot(p, Ap, pAp)                                 # linsys.py:631-634
            ctx.cg_update(Ap[0], p[0], yk[0], r[0], None, None, rhos[-1], pAp, 0)
            zz = ls.M._apply_dev(r, out=None if M_is_id else z)           # linsys.py:661
            if M_is_id:
                z = r
            utils._ip_coef(r, zz, ls.ip_B, tmp, post=1)                    # linsys.py:664
            MMlrk_norm = numpy.float64(tmp[0].item())
            alpha = float(mb[1])
            rhos.append(MMlrk_norm ** 2)                                   # linsys.py:665
            if store:
                if k > 0:
                    self.H[k - 1, k] = self.H[k, k - 1]
                    self.H[k, k] = (1.0 + alpha * omega / alpha_old) / alpha
                else:
                    self.H[k, k] = 1.0 / alpha
                tmp.fill_(float(MMlrk_norm))
                sgn = (-1.0) ** (k + 1)
                ctx.scale_dev(tmp, 1, sgn, z[0], self._Vd[k + 1])          # linsys.py:669
                if self._Pd is not None:
                    ctx.scale_dev(tmp, 1, sgn, r[0], self._Pd[k + 1])      # linsys.py:671
                self.H[k + 1, k] = numpy.sqrt(rhos[-1] / rhos[-2]) / alpha
                alpha_old = alpha
            self._Mlrk_dev, self._MMlrk_dev = r, z
            rkn = self._finalize_iteration(yk, MMlrk_norm)                 # linsys.py:678
            if rkn is not None:
                rhos[-1] = rkn ** 2                                        # linsys.py:681-683
            self.iter += 1

        self.Mlrk, self.MMlrk = r, z
        if self.__dict__.get("_xk_dev") is None:
            self.xk = self._get_xk(yk)

    def _finalize(self):
        """krypy/linsys.py:691-696."""
        if self.store_arnoldi:
            ctx = self._ctx
            nc, dt = self.iter + 1, self.dtype
            Vd, Pd = self._Vd, self._Pd           # (no reference to self in the thunks: no cycle)
            self.V = lambda: ctx.to_numpy(Vd[:nc]).astype(dt, copy=False)
            if Pd is not None:
                self.P = lambda: ctx.to_numpy(Pd[:nc]).astype(dt, copy=False)
            self.H = self.H[: self.iter + 1, : self.iter]

    @staticmethod
    def operations(nsteps):
        """krypy/linsys.py:698-708."""
        return {"A": 1 + nsteps, "M": 2 + nsteps, "Ml": 2 + nsteps, "Mr": 1 + nsteps,
                "ip_B": 2 + 2 * nsteps, "axpy": 2 + 2 * nsteps}


class Minres(_KrylovSolver):
    """Preconditioned MINRES (krypy/linsys.py:711-874).

    Per iteration on the device: operator apply, ONE fused Lanczos kernel
    (three-term recurrence + dot + update + norm + normalised store,
    kry_orth_fused), the sliding-QR recurrence (kry_minres_recur) and ONE fused
    sweep for ``z, W, y`` (kry_minres_update)."""

    def __init__(self, linear_system, ortho="lanczos", **kwargs):
        if not linear_system.self_adjoint:
            warnings.warn("Minres applied to a non-self-adjoint linear system. Consider using Gmres.")
        self.ortho = ortho
        super(Minres, self).__init__(linear_system, **kwargs)

    def __repr__(self):
        return self._repr("MINRES")

    V = _LazyBasis("V")
    P = _LazyBasis("P")

    def _solve(self):
        """krypy/linsys.py:791-853."""
        ctx = self._ctx
        ls = self.linear_system
        N = ls.N
        td = self._td
        self.lanczos = lz = utils.Arnoldi(
            self.MlAMr, self.__dict__["_Mlr0_dev"], maxiter=self.maxiter, ortho=self.ortho, M=ls.M,
            Mv=self.__dict__["_MMlr0_dev"], Mv_norm=self.MMlr0_norm, ip_B=ls.ip_B, dtype=self.dtype)
        W0 = ctx.zeros((1, N), td)
        W1 = ctx.zeros((1, N), td)
        yk = ctx.zeros((1, N), td)
        st = ctx.scalars(16)
        h3 = ctx.scalars(3)
        st[6:7].fill_(float(self.MMlr0_norm))                       # y = [||r0||, 0], linsys.py:809
        mb = ctx.mailbox
        is_lanczos = self.ortho == "lanczos"
        # Look-ahead (three-term Lanczos): the operator apply and the Lanczos kernel of step k+1 -- three
        # quarters of an iteration's bytes -- are enqueued BEFORE the host waits for step k's residual,
        # so the device works through the host's bookkeeping.  A Lanczos step does not depend on the
        # host's decision; if the loop ends at k the speculative step only wrote V[k+2] and the
        # three-entry accumulator, which nothing reads afterwards.  The solution update of step k+1
        # (which changes y, W) is launched only after the decision.
        lookahead = is_lanczos
        ev = ctx.event()
        launched = -1
        while (self.resnorms[-1] > self.tol and lz.iter < lz.maxiter and not lz.invariant):
            k = self.iter = lz.iter
            if launched < k:
                lz._enqueue(k)                                       # linsys.py:823
                launched = k
            if is_lanczos:
                ctx.minres_recur(k, lz._lz, st, 1, 0)                # linsys.py:827-841, 847
                ev.record()
            elif lz._cplx:
                # the recurrence takes the real parts (linsys.py:828-833): gather them from the
                # interleaved column
                h3.copy_(lz._hcol_store[2 * k:2 * k + 6:2])
                ctx.minres_recur(k, h3, st, 0, 0)
            else:
                ctx.minres_recur(k, lz._hcol_store[k:], st, 0, 0)    # [H[k-1,k], H[k,k], H[k+1,k]]
            ctx.minres_update(lz._Vd[k], W0[0], W1[0], yk[0], st)    # linsys.py:844-846
            W0, W1 = W1, W0
            if is_lanczos:
                if lookahead and k + 1 < lz.maxiter:
                    lz._enqueue(k + 1)
                    launched = k + 1
                ev.synchronize()
                resid = float(mb[0])
                lz._finish(k, mb[6:8].copy())
            else:
                hc = lz._read_hcol(k)                                # synchronises
                resid = float(mb[0])
                lz._finish(k, hc)
            self._finalize_iteration(yk, resid)                      # linsys.py:849
        if launched >= 0 and launched == lz.iter:      # step lz.iter was enqueued but not consumed
            self._discard_speculative()
        if self.__dict__.get("_xk_dev") is None:
            self.xk = self._get_xk(yk)

    def _discard_speculative(self):
        """hook: a look-ahead Lanczos step was enqueued but not consumed"""
        pass

    def _finalize(self):
        """krypy/linsys.py:855-862."""
        if self.store_arnoldi:
            _store_arnoldi_lazily(self, self.lanczos,
                                  not isinstance(self.linear_system.M, utils.IdentityLinearOperator))

    @staticmethod
    def operations(nsteps):
        """krypy/linsys.py:864-874."""
        return {"A": 1 + nsteps, "M": 2 + nsteps, "Ml": 2 + nsteps, "Mr": 1 + nsteps,
                "ip_B": 2 + 2 * nsteps, "axpy": 4 + 8 * nsteps}


class Gmres(_KrylovSolver):
    """Preconditioned GMRES (krypy/linsys.py:877-1018).

    Per iteration on the device: operator apply (kry_spmv_csr), ONE fused
    Gram-Schmidt kernel (kry_orth_fused) and the Givens/Hessenberg update
    (kry_givens_update), which publishes ``|y[k+1]|`` and the H and R columns
    through the pinned mailbox -- the single host synchronisation of the step.

    ``ortho``: 'mgs' (default, the reference's exact order), 'dmgs', 'lanczos',
    and the block variants 'cgs' / 'cgs2' (see utils.Arnoldi).
    """

    def __init__(self, linear_system, ortho="mgs", _workspace=None, _prelaunch=False, **kwargs):
        self.ortho = ortho
        self._ws = _workspace
        # (restart driver) another cycle follows over the same workspace unless this one converges: the end of
        # this cycle may launch it speculatively, see _solve
        self._prelaunch = bool(_prelaunch)
        super(Gmres, self).__init__(linear_system, **kwargs)

    def __repr__(self):
        return self._repr("GMRES", ["    R: {} x {} matrix\n".format(*self.R.shape),
                                    "    V: {} x {} matrix\n".format(*self.V.shape)])

    # without store_arnoldi, V aliases the Arnoldi object's full-width basis (linsys.py:981)
    V = _LazyBasis("V", fallback=lambda self: self.arnoldi.V)
    P = _LazyBasis("P")

    def _get_xk(self, y, k=None):
        """krypy/linsys.py:941-949: y is a device vector holding y[:k] (or None); k defaults to the
        number of Arnoldi steps booked so far."""
        x0d = self.__dict__["_x0_dev"]
        if y is None:
            return x0d
        ctx = self._ctx
        if k is None:
            k = self.arnoldi.iter
        if k > 0:
            t = _device.torch()
            ar = self.arnoldi
            nr = ar._nr
            Rt = self.__dict__.get("_Rt_dev")
            yy = ctx.scalars(nr * k) if Rt is None else self._ws.tensor("yy", (Rt.shape[0],),
                                                                        lambda: ctx.scalars(Rt.shape[0]))
            if Rt is not None:
                ctx.tri_solve_t(k, Rt, y, yy)                              # R never left the device
            elif ar._cplx:
                Rk = t.from_numpy(_cplx.to_pairs(self.R[:k, :k])).to(ctx.device)   # (k, 2k) interleaved
                ctx.tri_solve_z(k, Rk, y, yy)
            else:
                Rk = t.from_numpy(numpy.ascontiguousarray(self.R[:k, :k], dtype=numpy.float64)).to(ctx.device)
                ctx.tri_solve(k, Rk, y, yy)                                # linsys.py:946
            out = ctx.empty(x0d.shape, x0d.dtype)
            Mr = self.linear_system.Mr
            # (complex: the interleaved yy are the real coefficients over the twin rows v_j, i v_j)
            if isinstance(Mr, utils.IdentityLinearOperator):
                ctx.block_combine(ar._Vt, nr * k, yy, x0d[0], out[0])      # x0 + V[:, :k] yy
            else:
                yk = ctx.empty(x0d.shape, x0d.dtype)
                ctx.block_combine(ar._Vt, nr * k, yy, None, yk[0])         # linsys.py:947
                Mry = Mr._apply_dev(yk)
                ctx.axpby(1.0, x0d[0], 1.0, Mry[0], out[0])                 # linsys.py:948
            return out
        return x0d

    def _solve(self):
        """krypy/linsys.py:951-997."""
        ctx = self._ctx
        ls = self.linear_system
        # The Givens / back-substitution recurrences run in ONE CTA with the column and the rotations in
        # shared memory: at most 2000 (complex: 1000) steps per cycle.  The reference has no such limit;
        # say so before any work is done instead of failing at step 2001 (maxiter defaults to N).
        cap = 1000 if self._td == _device.torch().complex128 else 2000
        if self.maxiter > cap:
            raise utils.ArgumentError(
                "Gmres on the device path supports at most %d steps per cycle (maxiter=%d; the default is "
                "N): pass maxiter<=%d or use RestartedGmres(ls, maxiter=m, max_restarts=r)"
                % (cap, self.maxiter, cap))
        m = self.maxiter
        ws = self._ws
        # a cycle the previous solver over this workspace launched speculatively for exactly this start
        pre = None
        if ws is not None:
            pre, ws.prelaunched = getattr(ws, "prelaunched", None), None
            if pre is not None and not (pre["Mlr"] is self.__dict__.get("_Mlr0_dev") and pre["m"] == m
                                        and pre["ortho"] == self.ortho and pre["ls"] is ls
                                        and not self.explicit_residual and not self.store_arnoldi
                                        and type(self) is Gmres
                                        and os.environ.get("KRY_CYCLE_AHEAD", "1") not in ("0", "")):
                # not this start: the usual set-up below is enqueued behind the stale cycle and overwrites
                # what it left
                pre = None
        if pre is not None:
            ws.prelaunch_hits = getattr(ws, "prelaunch_hits", 0) + 1
        self.arnoldi = ar = utils.Arnoldi(
            self.MlAMr, self.__dict__["_Mlr0_dev"], maxiter=self.maxiter, ortho=self.ortho, M=ls.M,
            Mv=self.__dict__["_MMlr0_dev"], Mv_norm=self.MMlr0_norm, ip_B=ls.ip_B, dtype=self.dtype,
            _workspace=self._ws, _prelaunched=pre is not None)
        self.R = numpy.zeros([m + 1, m], dtype=utils._common_type([self.dtype, numpy.float64]))
        cplx, nr = ar._cplx, ar._nr            # complex: every small quantity is an interleaved pair
        if ws is not None:
            self._y_dev = y = ws.tensor("y", (nr * (m + 2),), lambda: ctx.scalars(nr * (m + 2)))
            cs = ws.tensor("cs", (2 * nr * (m + 1),), lambda: ctx.scalars(2 * nr * (m + 1)))
            rcol = ws.tensor("rcol", (nr * (m + 2),), lambda: ctx.scalars(nr * (m + 2)))
            if pre is None:
                y.zero_()
        else:
            self._y_dev = y = ctx.scalars(nr * (m + 2))
            cs = ctx.scalars(2 * nr * (m + 1))
            rcol = ctx.scalars(nr * (m + 2))
        givens = ctx.givens_update_z if cplx else ctx.givens_update
        if pre is None:
            y[0:1].fill_(float(self.MMlr0_norm))                           # linsys.py:969
        # CUDA graphs: from the second cycle over the same workspace on, step k is one graph launch
        use_graphs = (ws is not None and type(self) is Gmres and ws.graphs_enabled(ctx)
                      and self.ortho not in ("lanczos", "house"))
        ws_warm = ws is not None and ws.uses >= 1      # buffers and lazy kernel set-up exist already
        if ws is not None:
            ws.uses += 1
        mb = ctx.mailbox
        is_lanczos = self.ortho == "lanczos"
        t = _device.torch()
        HALF = 8192                          # two mailbox halves: step k uses half k & 1
        events = (ctx.event(), ctx.event())
        # Look-ahead: step k+1 is enqueued BEFORE the host waits for step k, so the device
        # never idles on the host's convergence test.  An Arnoldi step does not depend on the
        # host's decision; if the loop ends at k the speculative step is simply discarded
        # (it only touched V[k+2], y[k+1:], cs[2k+2:], which nothing reads afterwards).
        lookahead = (not self.explicit_residual and ls.exact_solution is None and not is_lanczos
                     and 2 * nr * (m + 2) + 1 <= HALF)

        # Whole cycle ahead (restarted runs, from the third cycle over one workspace on): ALL maxiter steps of
        # the cycle are enqueued at once -- as ONE CUDA graph where graphs are on (row-partitioned runs), step by
        # step otherwise -- with one mailbox record per step: look-ahead taken to its end.  An Arnoldi step never
        # depends on the host's decisions, so steps past the one the host stops at are speculative work that
        # nothing reads (as with the one-step look-ahead).  The host synchronises once per cycle and books the
        # records of the steps that neither converge nor look invariant in bulk.  Measured on 8 B200 (C2):
        # 92 us/step inside the cycle graph against 115 us paced by the host
        # (profiles/r2_cycle_graph_probe.txt).  KRY_CYCLE_AHEAD=0 keeps the step-by-step pace.
        rec = [2 * nr * (j + 2) + 1 for j in range(m)]
        cyc_offs = numpy.concatenate([[0], numpy.cumsum(rec)]).astype(numpy.int64)
        cycle_mode = (ws is not None and type(self) is Gmres and getattr(ctx, "cycle_ahead", False)
                      and self.ortho not in ("lanczos", "house") and lookahead and ws.uses >= 3 and not cplx
                      and int(cyc_offs[-1]) <= _lib.KRY_MAILBOX_DOUBLES
                      and os.environ.get("KRY_CYCLE_AHEAD", "1") not in ("0", "")
                      # (nothing to run ahead when the start already meets the tolerance or is the zero vector)
                      and (pre is not None or (self.resnorms[-1] > self.tol and not ar.invariant)))

        def off_of(k):
            if cycle_mode:
                return int(cyc_offs[k])
            return (k & 1) * HALF if lookahead else 0

        # restarted runs keep R on the device: step k's rotated column goes to row k of Rt (column after column),
        # and the solution update solves against it in place (kry_tri_solve_t) instead of uploading the host copy
        Rt = None
        if (ws is not None and type(self) is Gmres and getattr(ctx, "cycle_ahead", False) and not cplx
                and self.ortho not in ("lanczos", "house")):
            Rt = self._Rt_dev = ws.tensor("Rt", (m, m + 2), lambda: ctx.zeros((m, m + 2), t.float64))

        def step(k):
            # Arnoldi step (linsys.py:978) + Givens / Hessenberg update (linsys.py:982-991); row-partitioned
            # block-CGS runs fold the latter into the step's last kernel
            rk_dev = rcol if Rt is None else Rt[k]
            tail = None if cplx else (rk_dev, cs, y, off_of(k))
            if not ar._enqueue(k, givens=tail):
                givens(k, ar._hcol, rk_dev, cs, y, off_of(k))

        def launch(k):
            g = ws.graphs.get(k) if use_graphs else None
            if g is None and use_graphs and ws_warm:
                g = t.cuda.CUDAGraph()
                with t.cuda.graph(g):
                    ctx.use_current_stream()
                    step(k)
                ctx.use_current_stream()
                ws.graphs[k] = g
            if g is not None:
                g.replay()
            else:
                step(k)
            events[k & 1].record()

        launched = -1
        k = -1
        _mark(self, "arnoldi_init")
        def launch_cycle():
            """all m steps of a cycle over the workspace's buffers: one graph replay, or m eager steps"""
            if use_graphs:
                g = ws.graphs.get("cycle")
                if g is None:
                    g = t.cuda.CUDAGraph()
                    with t.cuda.graph(g):
                        ctx.use_current_stream()
                        for j in range(m):
                            step(j)
                    ctx.use_current_stream()
                    ws.graphs["cycle"] = g
                g.replay()
            else:
                for j in range(m):
                    step(j)
            ev = ctx.event()
            ev.record()
            return ev

        if cycle_mode:
            if pre is not None:
                pre["event"].synchronize()                     # launched by the previous cycle's solver
            else:
                launch_cycle().synchronize()
            launched = m - 1
            resid_all = mb[cyc_offs[:m]] / ls.MMlb_norm
            if m >= 2 and bool(numpy.all(resid_all[:m - 1] > self.tol)):
                # The usual cycle: no step before the last one meets the tolerance.  The device work of the cycle's
                # end (solution update, explicit residual, its norm and the read-back of that norm) is enqueued
                # FIRST, for all m steps; while it runs the host books the records (H, R, history, invariance
                # test), launches the next cycle speculatively on the device-side norm when the restart driver
                # announced one, and only then waits for the norm.  If the bookkeeping finds a step that needs a
                # decision after all (an invariant-looking subspace), the early results are dropped and the
                # step-by-step loop below takes over -- nothing has been overwritten at that point.
                # (speculation only while the cycle's last updated residual is well above the tolerance: the
                #  explicit residual then cannot meet it, and no cycle is ever launched for nothing)
                can_pre = (self._prelaunch and Rt is not None and ar.M is None and ar._euclid
                           and not self.store_arnoldi and bool(resid_all[m - 1] > 2.0 * self.tol)
                           and os.environ.get("KRY_PRELAUNCH", "1") not in ("0", ""))
                state = {"ok": False}

                def before_sync(MMlr, Mlr, nrm_dev):
                    snap = mb[: int(cyc_offs[m])].copy()        # the next cycle reuses the record slots
                    if can_pre and MMlr is Mlr and self._no_invariance_in_sight(ar, snap, cyc_offs, m):
                        # nothing below can end the cycle early any more: the next cycle goes first
                        y.zero_()
                        y[0:1].copy_(nrm_dev[0:1])
                        ctx.scale_dev(nrm_dev, 1, 1.0, Mlr[0], ar._Vd[0])          # v_0 = r / ||r||
                        if ctx.comm is not None:
                            ctx.comm.halo_ready = None
                        ws.prelaunched = dict(Mlr=Mlr, m=m, ortho=self.ortho, ls=ls, event=launch_cycle())
                    fill = self._book_cycle_records(ar, snap, cyc_offs, m)
                    if fill is None:
                        return
                    kl = m - 1
                    off = off_of(kl)
                    nh = kl + 2
                    hcol = snap[off + 1:off + 1 + nh].copy()
                    self.R[: kl + 2, kl] = snap[off + 1 + nh:off + 1 + 2 * nh]
                    self.iter = kl
                    _mark(self, "last_iteration_begin")
                    ar._finish(kl, hcol)
                    fill()
                    state["ok"] = True

                xk_early = self._get_xk(y, k=m)
                MMlrk, Mlrk, rkn = ls._get_residual_dev(xk_early, compute_norm=True, before_sync=before_sync)
                if state["ok"]:
                    # krypy/linsys.py:430-493 for the last iteration of the cycle
                    k = m - 1
                    self.xk = xk_early
                    self._last_residual = (MMlrk, Mlrk, rkn)
                    self.resnorms.append(rkn / ls.MMlb_norm)
                    if self.resnorms[-1] > self.tol:
                        self._finalize()
                        raise utils.ConvergenceError(
                            ("No convergence in last iteration "
                             "(maxiter: %s, residual: %s)." % (self.maxiter, self.resnorms[-1])), self)
                else:
                    self.xk = None
                    self._last_residual = None
            else:
                fill = self._book_cycle_records(ar, mb, cyc_offs, m)
                if fill is not None:
                    fill()
        elif pre is not None:
            raise RuntimeError("a speculative cycle was accepted outside the cycle-ahead mode")      # (same conditions)
        while (self.resnorms[-1] > self.tol and ar.iter < ar.maxiter and not ar.invariant):
            k = self.iter = ar.iter
            if k == ar.maxiter - 1:
                _mark(self, "last_iteration_begin")
            if is_lanczos:
                ar._enqueue(k)
                # tridiagonal column from the three Lanczos entries
                ctx.minres_recur(k, ar._lz, ar._lz_st, 1, 16)
                ctx.sync()
                hcol = numpy.zeros(nr * (k + 2))
                if k > 0:
                    hcol[nr * (k - 1)] = mb[16 + 5]
                hcol[nr * k], hcol[nr * (k + 1)] = mb[16 + 6], mb[16 + 7]
                ar._hcol[: nr * (k + 2)].copy_(t.from_numpy(hcol))
                givens(k, ar._hcol, rcol, cs, y, off_of(k))
                events[k & 1].record()
                launched = k
            if launched < k:
                launch(k)
                launched = k
            if lookahead and k + 1 < ar.maxiter and launched < k + 1:
                launch(k + 1)
                launched = k + 1
            if not cycle_mode:
                events[k & 1].synchronize()
            off = off_of(k)
            resid = float(mb[off])
            nh = nr * (k + 2)
            hcol = mb[off + 1:off + 1 + nh].copy()
            rk = mb[off + 1 + nh:off + 1 + 2 * nh].copy()
            if cplx:
                hcol = _cplx.from_pairs(hcol.reshape(1, -1))[0]
                rk = _cplx.from_pairs(rk.reshape(1, -1))[0]
            self.R[: k + 2, k] = rk
            if is_lanczos:
                ar._finish(k, numpy.real(hcol[k:k + 2]))
            else:
                ar._finish(k, hcol)
            self._finalize_iteration(y, resid)                             # linsys.py:993
        if launched > k:
            self._discard_speculative()
        if self.__dict__.get("_xk_dev") is None:
            self.xk = self._get_xk(y if ar.iter > 0 else None)

    def _no_invariance_in_sight(self, ar, mb, offs, m):
        """cheap sufficient condition for "the invariant-subspace test (utils.py:1035-1039) fires at no step of
        this cycle": every H[j+1, j] stays above 1e-14 times the Frobenius norm of the WHOLE Hessenberg matrix
        (an upper bound of the norm the test uses at step j)"""
        idx = self._ws.bufs.get(("cycle_hpos", m))
        if idx is None:
            hpos = numpy.concatenate([offs[j] + 1 + numpy.arange(j + 2) for j in range(m)])
            sub = numpy.array([offs[j] + 1 + j + 1 for j in range(m)])
            idx = self._ws.bufs[("cycle_hpos", m)] = (hpos, sub)
        hpos, sub = idx
        h = mb[hpos]
        total = ar._hfro2 + float(numpy.dot(h, h))
        return bool(numpy.all(mb[sub] > 1e-14 * numpy.sqrt(total)))

    def _book_cycle_records(self, ar, mb, offs, m):
        """Whole-cycle graph: books, in bulk, the mailbox records of the leading steps of the cycle that need no
        decision -- the updated residual stays above the tolerance, the step is not the last one and the
        invariant-subspace test (utils.py:1035-1039) is far from firing.  Exactly what the loop in _solve does
        step by step (H, R, resnorms, the running Frobenius norm, iteration counters); the loop continues with
        the first step that is not that simple.  In the usual case (all steps but the last) only the counters
        are set here and a callable is returned that fills H, R and the history later, while the device is busy
        with the end of the cycle; otherwise everything is done at once and None is returned."""
        ls = self.linear_system
        resid = mb[offs[:m]] / ls.MMlb_norm
        stop = numpy.nonzero(~(resid > self.tol))[0]                   # (also catches NaN)
        kb = min(int(stop[0]) if stop.size else m, m - 1)
        if kb <= 0:
            return None
        idx = self._ws.bufs.get(("cycle_index", m))
        if idx is None:
            # positions of H[:j+2, j] / R[:j+2, j] inside the mailbox, column after column
            rows = numpy.concatenate([numpy.arange(j + 2) for j in range(m)])
            cols = numpy.concatenate([numpy.full(j + 2, j) for j in range(m)])
            hpos = numpy.concatenate([offs[j] + 1 + numpy.arange(j + 2) for j in range(m)])
            rpos = numpy.concatenate([offs[j] + 1 + (j + 2) + numpy.arange(j + 2) for j in range(m)])
            ends = numpy.cumsum([j + 2 for j in range(m)])
            idx = self._ws.bufs[("cycle_index", m)] = (rows, cols, hpos, rpos, ends)
        rows, cols, hpos, rpos, ends = idx
        hvals = mb[hpos[: ends[kb - 1]]]
        col2 = numpy.add.reduceat(hvals * hvals, numpy.concatenate([[0], ends[: kb - 1]]))
        hfro2 = ar._hfro2 + numpy.cumsum(col2)
        hk = mb[offs[:kb] + 1 + numpy.arange(1, kb + 1)]               # H[j+1, j]
        sus = numpy.nonzero(~(hk > 1e-14 * numpy.sqrt(hfro2)))[0]      # invariant-looking (or NaN) steps
        if sus.size:
            kb = int(sus[0])
            if kb <= 0:
                return None
        ne = int(ends[kb - 1])
        ar._hfro2 = float(hfro2[kb - 1])
        ar.iter = kb
        self.iter = kb - 1
        self.xk = None
        self._last_residual = None
        hist = [float(v) for v in resid[:kb]]
        self.resnorms.append(hist[-1])          # (the loop condition looks at the latest entry)
        rvals = mb[rpos[:ne]]                           # (fancy indexing copies: safe against the next cycle's records)
        pos = len(self.resnorms) - 1

        ar.H[rows[:ne], cols[:ne]] = hvals[:ne]         # (the invariance test of the next step may look at H)

        def fill():
            self.R[rows[:ne], cols[:ne]] = rvals
            self.resnorms[pos:pos + 1] = hist
        if kb == m - 1:
            return fill
        fill()
        return None

    def _discard_speculative(self):
        """hook: a look-ahead Arnoldi step was enqueued but not consumed"""
        pass

    def _finalize(self):
        """krypy/linsys.py:999-1006."""
        if self.store_arnoldi:
            _store_arnoldi_lazily(self, self.arnoldi,
                                  not isinstance(self.linear_system.M, utils.IdentityLinearOperator))

    @staticmethod
    def operations(nsteps):
        """krypy/linsys.py:1008-1018."""
        return {"A": 1 + nsteps, "M": 2 + nsteps, "Ml": 2 + nsteps, "Mr": 1 + nsteps,
                "ip_B": 2 + nsteps + nsteps * (nsteps + 1) / 2,
                "axpy": 4 + 2 * nsteps + nsteps * (nsteps + 1) / 2}


class _RestartedSolver(object):
    """krypy/linsys.py:1021-1072."""

    def __init__(self, Solver, linear_system, max_restarts=0, **kwargs):
        self.xk = None
        kwargs = dict(kwargs)
        self.resnorms = [numpy.inf]
        if linear_system.exact_solution is not None:
            self.errnorms = [numpy.inf]
        tol = None
        restart = 0
        xk_dev = None
        if Solver is Gmres and "_workspace" not in kwargs:
            kwargs["_workspace"] = utils.SolverWorkspace()
        self._workspace = kwargs.get("_workspace")
        try:
            while restart == 0 or (self.resnorms[-1] > tol and restart <= max_restarts):
                try:
                    if xk_dev is not None:
                        kwargs.update({"x0": xk_dev})        # stays in HBM between cycles
                    if Solver is Gmres:
                        kwargs["_prelaunch"] = restart < max_restarts      # another cycle follows unless this one converges
                    sol = Solver(linear_system, **kwargs)
                except utils.ConvergenceError as e:
                    sol = e.solver
                xk_dev = sol.__dict__["_xk_dev"]
                if xk_dev is None:
                    xk_dev = _ctx().to_block(sol.xk, sol._td)
                xk_dev = xk_dev.reshape(-1)                  # flat: keeps flat_vecs semantics neutral
                kwargs["_x0_residual"] = sol.__dict__.get("_last_residual")
                self._last = sol
                tol = sol.tol
                del self.resnorms[-1]
                self.resnorms += sol.resnorms
                if linear_system.exact_solution is not None:
                    del self.errnorms[-1]
                    self.errnorms += sol.errnorms
                restart += 1
        finally:
            _ctx().l2_window(None)
        self.xk = sol.xk
        self.tol = tol
        if self.resnorms[-1] > tol:
            raise utils.ConvergenceError("No convergence after %d restarts." % max_restarts, self)


class RestartedGmres(_RestartedSolver):
    """Restarted GMRES (krypy/linsys.py:1075-1081)."""

    def __init__(self, *args, **kwargs):
        super(RestartedGmres, self).__init__(Gmres, *args, **kwargs)
