"""Host-side mirror of ``krypy.utils`` for the hot path (SURVEY.md section 8a:
rows a1-a13, a24): same names, argument meaning and error behaviour as the
reference, with every N-sized operation executed by the sm_100a kernels of
libkrypy_b200.so through ctypes.

Public functions take and return ``(N, k)`` numpy arrays exactly like the
reference; internally blocks live in HBM as ``(k, N)`` torch tensors
("vector-major") and the solver classes call the ``*_dev`` methods
directly so nothing N-sized crosses PCIe inside an iteration.

Complex systems run on the same (real) kernels by real embedding: a complex block is a
``torch.complex128`` tensor whose interleaved real view is what the kernels see, and every block
that is the left operand of inner products or a combination basis is kept in *twin storage*
(``_Twin``: rows ``v_0, i v_0, v_1, i v_1, ...``), which turns complex Gram-Schmidt against k
vectors into real Gram-Schmidt against 2k vectors (krypy_b200/_cplx.py).  The Arnoldi hot loop of a
complex solve runs on native complex kernels instead (``kry_orth_fused_z`` on the even rows of the twin
storage, ``kry_spmv_csr_z`` on the matrix as it is; ``_NATIVE_Z``).

There is no CPU fallback: anything the device path cannot do (complex or Householder
row-partitioned runs) raises ``NotImplementedError``.
"""
import time
import warnings
from collections import defaultdict

import numpy

from . import _cplx, _device
from ._lib import KRY_ORTH_CGS, KRY_ORTH_MGS

__all__ = [
    "ArgumentError", "AssumptionError", "ConvergenceError", "LinearOperatorError",
    "InnerProductError", "RuntimeError", "Arnoldi", "Givens", "House",
    "IdentityLinearOperator", "ZeroLinearOperator", "LinearOperator",
    "MatrixLinearOperator", "DiagonalLinearOperator", "DeviceLinearOperator", "TimedLinearOperator", "Projection",
    "Timer", "Timings", "arnoldi", "arnoldi_res", "get_linearoperator", "inner", "ip_euclid",
    "norm", "norm_squared", "orthonormality", "qr", "shape_vec", "shape_vecs",
    "find_common_dtype", "DeviceBlock", "SolverWorkspace",
    # host-side analysis helpers (krypy_b200/_analysis.py)
    "BoundCG", "BoundMinres", "NormalizedRootsPolynomial", "Interval", "Intervals", "angles",
    "arnoldi_projected", "bound_perturbed_gmres", "gap", "hegedus", "norm_MMlr", "ritz", "strakos",
    "get_residual_norms",
]


# --------------------------------------------------------------------------
# exceptions -- krypy/utils.py:62-103
# --------------------------------------------------------------------------
class ArgumentError(Exception):
    """Raised when an argument is invalid (krypy/utils.py:62-67)."""


class AssumptionError(Exception):
    """Raised when an assumption is not satisfied (krypy/utils.py:70-78)."""


class ConvergenceError(Exception):
    """Raised when a method did not converge; ``solver`` holds the populated
    solver object (krypy/utils.py:81-91)."""

    def __init__(self, msg, solver):
        super(ConvergenceError, self).__init__(msg)
        self.solver = solver


class LinearOperatorError(Exception):
    """Raised when a LinearOperator cannot be applied (krypy/utils.py:94-95)."""


class InnerProductError(Exception):
    """Raised when the inner product is indefinite (krypy/utils.py:98-99)."""


class RuntimeError(Exception):
    """Errors that fit nowhere else (krypy/utils.py:102-103; shadows the builtin
    inside this module exactly like the reference)."""


# --------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------
def _ctx():
    return _device.Context.get()


def _is_dev(x):
    t = _device._torch
    return t is not None and isinstance(x, t.Tensor)


def _isspmatrix(A):
    import scipy.sparse as sp
    return sp.issparse(A)   # accepts csr_array as well (superset of the reference)


def _common_type(dtypes):
    dtypes = [numpy.dtype(d) for d in dtypes if d is not None]
    if not dtypes:
        return numpy.dtype(None)      # float64, as numpy.find_common_type([], []) did
    return numpy.result_type(*dtypes)


def find_common_dtype(*args):
    """krypy/utils.py:106-122: common dtype of ndarray / spmatrix / LinearOperator
    arguments; everything else (notably None) is ignored."""
    dtypes = []
    for arg in args:
        if type(arg) is numpy.ndarray or _isspmatrix(arg) or isinstance(arg, LinearOperator):
            if hasattr(arg, "dtype"):
                dtypes.append(arg.dtype)
            else:
                warnings.warn("object %s does not have a dtype." % arg.__repr__)
        elif _is_dev(arg):
            dtypes.append(_device.torch_to_np_dtype(arg.dtype))
    return _common_type(dtypes)


def _compute_dtype(npdtype):
    """numpy dtype -> torch dtype of the device path (complex64 promotes to complex128)."""
    npdtype = numpy.dtype(npdtype)
    if npdtype.kind == "c":
        return _device.np_to_torch_dtype(numpy.complex128)
    if npdtype == numpy.float32:
        return _device.np_to_torch_dtype(numpy.float32)
    return _device.np_to_torch_dtype(numpy.float64)   # ints, bools, float16 promote to fp64


def shape_vec(x):
    """Take a (n,) ndarray and return it as (n,1) ndarray (krypy/utils.py:125-127)."""
    return numpy.reshape(x, (x.shape[0], 1))


def shape_vecs(*args):
    """krypy/utils.py:130-143."""
    ret_args = []
    flat_vecs = True
    for arg in args:
        if type(arg) is numpy.ndarray:
            if len(arg.shape) == 1:
                arg = shape_vec(arg)
            else:
                flat_vecs = False
        ret_args.append(arg)
    return flat_vecs, ret_args


def _isintlike(x):
    try:
        return bool(int(x) == x) and numpy.ndim(x) == 0
    except (TypeError, ValueError):
        return False


# --------------------------------------------------------------------------
# complex blocks: twin storage (krypy_b200/_cplx.py)
# --------------------------------------------------------------------------
def _is_cplx(x):
    return _device.is_complex(x)


class _Twin(object):
    """Twin storage of k complex vectors of length N: a real ``(2k, ld)`` tensor whose row ``2j``
    is the interleaved ``v_j`` and whose row ``2j+1`` is ``i v_j``.

    ``T``: real ``(2k, 2N)`` view for the kernels (real Gram-Schmidt / block dots / combinations
    against these 2k rows ARE the complex operations against the k vectors, with interleaved
    complex coefficients); ``C``: complex ``(k, N)`` view of the even rows (the block itself)."""

    def __init__(self, ctx, k, N, op=None):
        t = _device.torch()
        self.k, self.N = int(k), int(N)
        store = ctx.alloc_basis(2 * max(self.k, 1), 2 * self.N, t.float64, None)
        self.store = store
        self.T = store[: 2 * self.k, : 2 * self.N]

    @property
    def C(self):
        """complex (k, N) view of the even rows.  The view points back at its storage (so
        ``_twin_for`` finds it and the storage lives as long as the view); the storage does not
        reference the view, so there is no reference cycle holding HBM until a cyclic GC."""
        t = _device.torch()
        c = self.store[0::2].view(t.complex128)[: self.k, : self.N]
        c._kry_twin = self
        return c

    def refresh(self, ctx, j):
        """row 2j+1 <- i * row 2j (after vector j was written)"""
        ctx.rot90(self.T[2 * j], self.T[2 * j + 1])

    @classmethod
    def of(cls, ctx, Xc):
        """twin storage holding a copy of the complex block Xc (k, N)"""
        tw = cls(ctx, Xc.shape[0], Xc.shape[1])
        if tw.k:
            tw.C.copy_(Xc)
            for j in range(tw.k):
                tw.refresh(ctx, j)
        return tw


def _twin_for(ctx, Xc):
    """the twin storage behind a complex block: the one it is a view of (blocks produced by
    ``_qr_dev``, Arnoldi bases), else a freshly built one cached on the tensor object.  Only for
    blocks that are no longer modified."""
    tw = getattr(Xc, "_kry_twin", None)
    if tw is None:
        tw = _Twin.of(ctx, Xc)
        try:
            Xc._kry_twin = tw
        except Exception:
            pass
    return tw


def _coefs_dev(ctx, c):
    """host coefficient vector (real or complex) -> device doubles (complex: interleaved)"""
    t = _device.torch()
    c = numpy.asarray(c).reshape(-1)
    if numpy.iscomplexobj(c):
        c = _cplx.to_pairs(c.reshape(1, -1)).reshape(-1)
    return t.from_numpy(numpy.ascontiguousarray(c, dtype=numpy.float64)).to(ctx.device)


def _combine(ctx, Vd, nv, coef_host, x0, out):
    """out = x0 + sum_j coef[j] V_j for a device block Vd and HOST coefficients (real block: real
    coefficients; complex block: complex coefficients over its twin storage)."""
    if _is_cplx(Vd):
        tw = _twin_for(ctx, Vd)
        c = numpy.zeros(nv, dtype=numpy.complex128)
        c[:] = numpy.asarray(coef_host).reshape(-1)[:nv]
        ctx.block_combine(tw.T, 2 * nv, _coefs_dev(ctx, c), x0, out)
    else:
        c = numpy.asarray(coef_host).reshape(-1)[:nv]
        if numpy.iscomplexobj(c):
            raise NotImplementedError("complex coefficients on a real block")
        ctx.block_combine(Vd, nv, _coefs_dev(ctx, c), x0, out)


def _caxpby(ctx, a, x, b, y, z):
    """z = a*x + b*y with possibly complex scalars on (real or complex) device vectors"""
    a, b = complex(a), complex(b)
    if (a.imag == 0.0 and b.imag == 0.0) or not _is_cplx(x):
        if a.imag != 0.0 or b.imag != 0.0:
            raise NotImplementedError("complex scaling of a real block")
        ctx.axpby(a.real, x, b.real, y, z)
        return
    t = _device.torch()
    ix = t.empty_like(x)
    ctx.rot90(x, ix)
    if y is None or b == 0:
        ctx.axpby(a.real, x, a.imag, ix, z)
        return
    if z.data_ptr() == y.data_ptr() and b == 1.0:
        ctx.axpby(a.real, x, 1.0, y, z)
        ctx.axpby(a.imag, ix, 1.0, z, z)
        return
    iy = t.empty_like(y)
    ctx.rot90(y, iy)
    tmp = t.empty_like(x)
    ctx.axpby(a.real, x, a.imag, ix, tmp)
    ctx.axpby(b.real, y, b.imag, iy, z)
    ctx.axpby(1.0, tmp, 1.0, z, z)


# --------------------------------------------------------------------------
# LinearOperator algebra -- krypy/utils.py:1365-1636
# --------------------------------------------------------------------------
class LinearOperator(object):
    """krypy/utils.py:1365-1456.  ``dot``/``dot_adj`` are user callbacks on
    ``(N, k)`` numpy arrays; subclasses living on the device override
    ``_apply_dev``.  A generic (host-callback) operator is applied to device
    blocks by a D2H copy, the user's callback and an H2D copy -- the callback is
    user code, the round trip is the price of it not being device code."""

    _device_native = False

    def __init__(self, shape, dtype, dot=None, dot_adj=None):
        if len(shape) != 2 or not _isintlike(shape[0]) or not _isintlike(shape[1]):
            raise LinearOperatorError("shape must be (m,n) with m and n integer")
        self.shape = shape
        self.dtype = numpy.dtype(dtype)  # defaults to float64
        if dot is None and dot_adj is None:
            raise LinearOperatorError("dot or dot_adj have to be defined")
        self._dot = dot
        self._dot_adj = dot_adj

    # -- public numpy API (reference semantics) --
    def dot(self, X):
        X = numpy.asanyarray(X)
        m, n = self.shape
        if X.shape[0] != n:
            raise LinearOperatorError("dimension mismatch")
        if self._dot is None:
            raise LinearOperatorError("dot undefined")
        if X.shape[1] == 0:
            return numpy.zeros(X.shape)
        return self._dot(X)

    def dot_adj(self, X):
        X = numpy.asanyarray(X)
        m, n = self.shape
        if X.shape[0] != m:
            raise LinearOperatorError("dimension mismatch")
        if self._dot_adj is None:
            raise LinearOperatorError("dot_adj undefined")
        if X.shape[1] == 0:
            return numpy.zeros(X.shape)
        return self._dot_adj(X)

    # -- device API (internal): Xd is a (k, N) tensor; result written to out --
    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        fn = self._dot_adj if adj else self._dot
        if fn is None:
            raise LinearOperatorError("dot_adj undefined" if adj else "dot undefined")
        Y = fn(ctx.to_numpy(Xd))
        Yd = ctx.to_block(numpy.asarray(Y), Xd.dtype)
        if out is not None:
            out.copy_(Yd)
            return out
        return Yd

    @property
    def adj(self):
        return _AdjointLinearOperator(self)

    def __mul__(self, X):
        try:
            if isinstance(X, IdentityLinearOperator):
                return self
            elif isinstance(self, IdentityLinearOperator):
                return X
            elif isinstance(X, LinearOperator):
                return _ProductLinearOperator(self, X)
            elif numpy.isscalar(X):
                return _ScaledLinearOperator(self, X)
            elif _is_dev(X):
                return self._apply_dev(X)
            else:
                return self.dot(X)
        except LinearOperatorError:
            return NotImplemented

    def __rmul__(self, X):
        try:
            return _ScaledLinearOperator(self, X)
        except LinearOperatorError:
            return NotImplemented

    def __pow__(self, X):
        try:
            return _PowerLinearOperator(self, X)
        except LinearOperatorError:
            return NotImplemented

    def __add__(self, X):
        try:
            return _SumLinearOperator(self, X)
        except LinearOperatorError:
            return NotImplemented

    def __neg__(self):
        try:
            return _ScaledLinearOperator(self, -1)
        except LinearOperatorError:
            return NotImplemented

    def __sub__(self, X):
        return self + (-X)

    def __repr__(self):
        m, n = self.shape
        return "<%dx%d %s with dtype=%s>" % (m, n, self.__class__.__name__, str(self.dtype))


class _DeviceOperator(LinearOperator):
    """Base of operators whose action is a kernel: the numpy-facing ``dot`` is
    H2D -> kernel -> D2H, the solvers call ``_apply_dev`` directly."""

    _device_native = True

    def __init__(self, shape, dtype):
        super(_DeviceOperator, self).__init__(shape, dtype, _DeviceOperator._dot, _DeviceOperator._dot_adj)
        # the class-level methods below take over: storing BOUND methods on the instance would make
        # every operator a reference cycle, and a cycle keeps its device arrays (the CSR matrix of a
        # MatrixLinearOperator: 0.8 GB on config C2) in HBM until Python's cyclic collector runs
        del self._dot, self._dot_adj

    def _dot(self, X):
        return self._dot_np(X)

    def _dot_adj(self, X):
        return self._dot_adj_np(X)

    def _np(self, X, adj):
        ctx = _ctx()
        X = numpy.asarray(X)
        dt = _compute_dtype(_common_type([self.dtype, X.dtype]))
        Yd = self._apply_dev(ctx.to_block(X, dt), adj=adj)
        return ctx.to_numpy(Yd)

    def _dot_np(self, X):
        return self._np(X, False)

    def _dot_adj_np(self, X):
        return self._np(X, True)

    def _apply_dev(self, Xd, out=None, adj=False):
        raise NotImplementedError


def _get_dtype(operators, dtypes=None):
    """krypy/utils.py:1459-1465."""
    if dtypes is None:
        dtypes = []
    for obj in operators:
        if obj is not None and hasattr(obj, "dtype"):
            dtypes.append(obj.dtype)
    return _common_type(dtypes)


class _SumLinearOperator(_DeviceOperator):
    """krypy/utils.py:1468-1483."""

    def __init__(self, A, B):
        if not isinstance(A, LinearOperator) or not isinstance(B, LinearOperator):
            raise LinearOperatorError("both operands have to be a LinearOperator")
        if A.shape != B.shape:
            raise LinearOperatorError("shape mismatch")
        self.args = (A, B)
        super(_SumLinearOperator, self).__init__(A.shape, _get_dtype([A, B]))

    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        Y0 = self.args[0]._apply_dev(Xd, adj=adj)
        Y1 = self.args[1]._apply_dev(Xd, adj=adj)
        if out is None:
            out = ctx.empty(Xd.shape, Xd.dtype)
        ctx.axpby(1.0, Y0, 1.0, Y1, out)
        return out


class _ProductLinearOperator(_DeviceOperator):
    """krypy/utils.py:1486-1501."""

    def __init__(self, A, B):
        if not isinstance(A, LinearOperator) or not isinstance(B, LinearOperator):
            raise LinearOperatorError("both operands have to be a LinearOperator")
        if A.shape[1] != B.shape[0]:
            raise LinearOperatorError("shape mismatch")
        self.args = (A, B)
        super(_ProductLinearOperator, self).__init__((A.shape[0], B.shape[1]), _get_dtype([A, B]))

    def _apply_dev(self, Xd, out=None, adj=False):
        if adj:
            T = self.args[0]._apply_dev(Xd, adj=True)
            return self.args[1]._apply_dev(T, out=out, adj=True)
        if getattr(self.args[0], "_inplace", False) and out is not None:
            # outer factor works in place on its argument (deflation projector): no temporary
            T = self.args[1]._apply_dev(Xd, out=out)
            return self.args[0]._apply_dev(T, out=out)
        T = self.args[1]._apply_dev(Xd)
        return self.args[0]._apply_dev(T, out=out)


class _ScaledLinearOperator(_DeviceOperator):
    """krypy/utils.py:1504-1519."""

    def __init__(self, A, alpha):
        if not isinstance(A, LinearOperator):
            raise LinearOperatorError("LinearOperator expected as A")
        if not numpy.isscalar(alpha):
            raise LinearOperatorError("scalar expected as alpha")
        self.args = (A, alpha)
        super(_ScaledLinearOperator, self).__init__(A.shape, _get_dtype([A], [type(alpha)]))

    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        alpha = self.args[1]
        Y = self.args[0]._apply_dev(Xd, adj=adj)
        if out is None:
            out = Y if Y is not Xd else ctx.empty(Xd.shape, Xd.dtype)
        if numpy.iscomplexobj(alpha) and complex(alpha).imag != 0.0:
            alpha = numpy.conj(alpha) if adj else alpha          # utils.py:1516-1519
            for j in range(Y.shape[0]):
                _caxpby(ctx, alpha, Y[j], 0.0, None, out[j])
            return out
        ctx.axpby(float(numpy.real(alpha)), Y, 0.0, None, out)
        return out


class _PowerLinearOperator(_DeviceOperator):
    """krypy/utils.py:1522-1545."""

    def __init__(self, A, p):
        if not isinstance(A, LinearOperator):
            raise LinearOperatorError("LinearOperator expected as A")
        if A.shape[0] != A.shape[1]:
            raise LinearOperatorError("square LinearOperator expected as A")
        if not _isintlike(p):
            raise LinearOperatorError("integer expected as p")
        self.args = (A, p)
        super(_PowerLinearOperator, self).__init__(A.shape, A.dtype)

    def _apply_dev(self, Xd, out=None, adj=False):
        res = Xd.clone()
        for _ in range(self.args[1]):
            res = self.args[0]._apply_dev(res, adj=adj)
        if out is not None:
            out.copy_(res)
            return out
        return res


class _AdjointLinearOperator(_DeviceOperator):
    """krypy/utils.py:1548-1556."""

    def __init__(self, A):
        if not isinstance(A, LinearOperator):
            raise LinearOperatorError("LinearOperator expected as A")
        self.args = (A,)
        m, n = A.shape
        super(_AdjointLinearOperator, self).__init__((n, m), A.dtype)

    def _apply_dev(self, Xd, out=None, adj=False):
        return self.args[0]._apply_dev(Xd, out=out, adj=not adj)


class IdentityLinearOperator(_DeviceOperator):
    """krypy/utils.py:1559-1569 (dtype float64, takes part in promotion, F3)."""

    def __init__(self, shape):
        super(IdentityLinearOperator, self).__init__(shape, numpy.dtype(None))

    def _dot_np(self, X):
        return X

    def _dot_adj_np(self, X):
        return X

    def _apply_dev(self, Xd, out=None, adj=False):
        if out is not None:
            if out.data_ptr() != Xd.data_ptr():
                out.copy_(Xd)
            return out
        return Xd


class ZeroLinearOperator(_DeviceOperator):
    """krypy/utils.py:1572-1582."""

    def __init__(self, shape):
        super(ZeroLinearOperator, self).__init__(shape, numpy.dtype(None))

    def _dot_np(self, X):
        return numpy.zeros(X.shape)

    def _dot_adj_np(self, X):
        return numpy.zeros(X.shape)

    def _apply_dev(self, Xd, out=None, adj=False):
        if out is not None:
            out.zero_()
            return out
        return _ctx().zeros(Xd.shape, Xd.dtype)


class MatrixLinearOperator(_DeviceOperator):
    """krypy/utils.py:1585-1602: a matrix as operator.  Sparse matrices are
    uploaded once per compute dtype as device CSR (kry_spmv_csr), dense arrays as
    row-major device matrices (kry_gemv_dense); the adjoint is materialised
    lazily like the reference's cached ``A.T.conj()`` (:1596-1599)."""

    def __init__(self, A):
        if _is_dev(A):
            if A.layout == _device.torch().sparse_csr:
                import scipy.sparse as sp
                A = sp.csr_matrix((A.values().cpu().numpy(), A.col_indices().cpu().numpy(),
                                   A.crow_indices().cpu().numpy()), shape=tuple(A.shape))
            else:
                A = A.detach().cpu().numpy()
        super(MatrixLinearOperator, self).__init__(A.shape, A.dtype)
        self._A = A
        self._A_adj = None
        self._devcache = {}

    def _dev(self, tdtype, adj=False):
        ctx = _ctx()
        native_z = bool(_NATIVE_Z) and ctx.comm is None and tdtype == _device.torch().complex128
        key = (tdtype, adj, native_z)
        obj = self._devcache.get(key)
        if obj is None:
            if adj and self._A_adj is None:
                self._A_adj = self._A.T.conj()
            A = self._A_adj if adj else self._A
            t = _device.torch()
            if tdtype == t.complex128:
                # complex vectors: the real kernels apply the 2N x 2N real embedding of A to the
                # interleaved real views (entry a+ib -> [[a, -b], [b, a]], krypy_b200/_cplx.py)
                # (sparse, default: the matrix as it is for the native complex SpMV kry_spmv_csr_z)
                if _isspmatrix(A) and native_z:
                    obj = ctx.upload_csr_z(A)
                elif _isspmatrix(A):
                    obj = ctx.upload_csr(_cplx.expand_sparse(A), t.float64)
                else:
                    obj = t.from_numpy(_cplx.expand_dense(numpy.asarray(A))).to(ctx.device)
            elif numpy.dtype(A.dtype).kind == "c":
                raise NotImplementedError("complex matrix applied to a real block")
            elif _isspmatrix(A):
                obj = ctx.upload_csr(A, tdtype)
            else:
                npdt = _device.torch_to_np_dtype(tdtype)
                obj = t.from_numpy(numpy.ascontiguousarray(numpy.asarray(A), dtype=npdt)).to(ctx.device)
            self._devcache[key] = obj
        return obj

    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        A = self._dev(Xd.dtype, adj)
        k = Xd.shape[0]
        if out is None:
            out = ctx.empty((k, self.shape[1] if adj else self.shape[0]), Xd.dtype)
        sparse = isinstance(A, _device.CsrDev)
        native_z = sparse and getattr(A, "native_z", False)
        for j in range(k):
            if native_z:
                ctx.spmv_z(A, Xd[j], out[j])
            elif sparse:
                ctx.spmv(A, Xd[j], out[j])
            else:
                ctx.gemv(A, Xd[j], out[j])
        return out

    def __repr__(self):
        return self._A.__repr__()


class DiagonalLinearOperator(_DeviceOperator):
    """A diagonal operator stored as one device vector (new; a sparse matrix
    whose pattern is exactly the main diagonal -- e.g. the Jacobi ``M`` of config
    C3 -- is turned into this by ``get_linearoperator``)."""

    def __init__(self, d):
        d = numpy.asarray(d).reshape(-1)
        super(DiagonalLinearOperator, self).__init__((d.shape[0], d.shape[0]), d.dtype)
        self._d = d
        self._devcache = {}
        self._as_matrix = None

    def _dev(self, tdtype):
        obj = self._devcache.get(tdtype)
        if obj is None:
            t = _device.torch()
            if tdtype == t.complex128:
                # real diagonal acting on interleaved complex data: every entry twice
                obj = _ctx().to_block(numpy.repeat(self._d.astype(numpy.float64), 2), t.float64)[0]
            else:
                obj = _ctx().to_block(self._d, tdtype)[0]
            self._devcache[tdtype] = obj
        return obj

    def _apply_dev(self, Xd, out=None, adj=False):
        ctx = _ctx()
        if numpy.dtype(self._d.dtype).kind == "c":
            # complex diagonal: the general (embedded sparse) path
            if self._as_matrix is None:
                import scipy.sparse as sp
                self._as_matrix = MatrixLinearOperator(sp.diags(self._d).tocsr())
            return self._as_matrix._apply_dev(Xd, out=out, adj=adj)
        d = self._dev(Xd.dtype)
        if out is None:
            out = ctx.empty(Xd.shape, Xd.dtype)
        for j in range(Xd.shape[0]):
            ctx.diag_mul(d, Xd[j], out[j])
        return out


class _FunctionDeviceOperator(_DeviceOperator):
    """Operator defined by a function on device blocks (used for the deflation
    projector, krypy/deflation.py:129-131)."""

    _inplace = True     # fn may overwrite its argument and return it

    def __init__(self, shape, dtype, fn):
        super(_FunctionDeviceOperator, self).__init__(shape, dtype)
        self._fn = fn

    def _apply_dev(self, Xd, out=None, adj=False):
        if adj:
            raise LinearOperatorError("dot_adj undefined")
        if out is None:
            out = Xd.clone()
        elif out.data_ptr() != Xd.data_ptr():
            out.copy_(Xd)
        Y = self._fn(out)
        if Y.data_ptr() != out.data_ptr():
            out.copy_(Y)
        return out


class DeviceLinearOperator(_DeviceOperator):
    """A user operator that runs on the device (new): ``dot_dev`` / ``dot_adj_dev``
    receive a ``(k, N)`` torch CUDA tensor -- k vectors, one per row (vector-major), complex128 for
    complex systems -- on the current stream and return a tensor of the same layout (or write into
    and return the ``out`` tensor passed as second argument).  Unlike ``LinearOperator(dot=...)``
    with a numpy callback nothing crosses PCIe.  The numpy-facing ``dot``/``*`` still work
    (H2D -> callback -> D2H)."""

    def __init__(self, shape, dtype, dot_dev=None, dot_adj_dev=None):
        if dot_dev is None and dot_adj_dev is None:
            raise LinearOperatorError("dot_dev or dot_adj_dev have to be defined")
        self._dot_dev, self._dot_adj_dev = dot_dev, dot_adj_dev
        super(DeviceLinearOperator, self).__init__(shape, dtype)

    def _apply_dev(self, Xd, out=None, adj=False):
        fn = self._dot_adj_dev if adj else self._dot_dev
        if fn is None:
            raise LinearOperatorError("dot_adj undefined" if adj else "dot undefined")
        m = self.shape[1] if adj else self.shape[0]
        if out is None:
            out = _ctx().empty((Xd.shape[0], m), Xd.dtype)
        Y = fn(Xd, out)
        if Y is None:
            Y = out
        if tuple(Y.shape) != (Xd.shape[0], m):
            raise LinearOperatorError("device callback returned shape %s, expected %s"
                                      % (tuple(Y.shape), (Xd.shape[0], m)))
        if Y.data_ptr() != out.data_ptr():
            out.copy_(Y)
        return out


class Timer(list):
    """``with timer: ...`` appends the duration of the block in seconds (krypy/utils.py:1289-1318).

    On the device path a block is bracketed by two CUDA events recorded on the current stream and the
    duration is read back LAZILY -- when the list is looked at -- so timing an operator application
    costs two event records and no host synchronisation (the reference's wall clock would need a
    device synchronise per application to mean anything here).  For host-only blocks on an idle
    stream the two event time stamps are host times, i.e. the wall-clock duration.  Without a CUDA
    device the wall clock is used."""

    def __init__(self):
        super(Timer, self).__init__()
        self._open = None
        self._pending = []           # [index, start event, end event, factor]

    def __enter__(self):
        t = _device._torch
        if t is not None and t.cuda.is_available():
            e0 = t.cuda.Event(enable_timing=True)
            e0.record()
            self._open = e0
        else:
            self._open = time.time()

    def __exit__(self, a, b, c):
        start, self._open = self._open, None
        if isinstance(start, float):
            list.append(self, time.time() - start)
            return
        e1 = _device._torch.cuda.Event(enable_timing=True)
        e1.record()
        self._pending.append([list.__len__(self), start, e1, 1.0])
        list.append(self, float("nan"))

    def scale_last(self, factor):
        """multiply the most recent entry by ``factor`` (per-vector time of a block application)
        without forcing its read-back"""
        n = list.__len__(self)
        if self._pending and self._pending[-1][0] == n - 1:
            self._pending[-1][3] *= factor
        elif n:
            list.__setitem__(self, n - 1, list.__getitem__(self, n - 1) * factor)

    def _resolve(self):
        if self._pending:
            pend, self._pending = self._pending, []
            pend[-1][2].synchronize()
            for idx, e0, e1, f in pend:
                list.__setitem__(self, idx, 1e-3 * e0.elapsed_time(e1) * f)

    def __getitem__(self, i):
        self._resolve()
        return list.__getitem__(self, i)

    def __iter__(self):
        self._resolve()
        return list.__iter__(self)

    def __repr__(self):
        self._resolve()
        return list.__repr__(self)

    def __eq__(self, other):
        self._resolve()
        return list.__eq__(self, other)

    __hash__ = None


class Timings(defaultdict):
    """A dictionary of timers keyed by operation name; ``get`` = the best (minimal) time seen
    (krypy/utils.py:1321-1362)."""

    def __init__(self):
        super(Timings, self).__init__(Timer)

    def get(self, key):
        times = list(self[key]) if key in self else []
        return min(times) if times else 0

    def get_ops(self, ops):
        """total time of ``{operation: number of applications}``"""
        return float(sum(self.get(op) * count for op, count in ops.items()))

    def __repr__(self):
        return "Timings(" + ", ".join("%s: %s" % (key, self.get(key)) for key in self) + ")"


class TimedLinearOperator(LinearOperator):
    """A linear operator whose applications are timed per vector (krypy/utils.py:1605-1636): the host
    entry points (``dot``, ``dot_adj``) and the device entry point the solvers use (``_apply_dev``)
    go through the same bracket."""

    def __init__(self, linear_operator, timer=None):
        self._linear_operator = op = linear_operator
        super(TimedLinearOperator, self).__init__(shape=op.shape, dtype=op.dtype, dot=op.dot, dot_adj=op.dot_adj)
        self._timer = Timer() if timer is None else timer

    def _timed(self, nvec, fn, *args, **kwargs):
        if nvec == 0:
            return fn(*args, **kwargs)
        with self._timer:
            ret = fn(*args, **kwargs)
        self._timer.scale_last(1.0 / nvec)
        return ret

    def dot(self, X):
        return self._timed(X.shape[1], self._linear_operator.dot, X)

    def dot_adj(self, X):
        # (an empty block goes to ``dot`` like in the reference, utils.py:1631)
        op = self._linear_operator
        return self._timed(X.shape[1], op.dot_adj if X.shape[1] else op.dot, X)

    def _apply_dev(self, Xd, out=None, adj=False):
        return self._timed(Xd.shape[0], self._linear_operator._apply_dev, Xd, out=out, adj=adj)


def _as_diagonal(A):
    """Return the diagonal if sparse A stores exactly its main diagonal, else None."""
    import scipy.sparse as sp
    if not sp.issparse(A) or A.shape[0] != A.shape[1]:
        return None
    A = A.tocsr() if not sp.isspmatrix_csr(A) else A
    n = A.shape[0]
    if A.nnz != n or n == 0:
        return None
    if not numpy.array_equal(A.indptr, numpy.arange(n + 1)):
        return None
    if not numpy.array_equal(A.indices, numpy.arange(n)):
        return None
    return A.data


def get_linearoperator(shape, A, timer=None):
    """krypy/utils.py:241-273 (plus csr_array, torch tensors, diagonal detection)."""
    ret = None
    import scipy.sparse.linalg as scipylinalg

    if isinstance(A, LinearOperator):
        ret = A
    elif A is None:
        ret = IdentityLinearOperator(shape)
    elif isinstance(A, numpy.matrix):
        ret = MatrixLinearOperator(numpy.atleast_2d(numpy.asarray(A)))
    elif isinstance(A, numpy.ndarray) or _isspmatrix(A) or _is_dev(A):
        d = _as_diagonal(A) if _isspmatrix(A) else None
        ret = DiagonalLinearOperator(d) if d is not None else MatrixLinearOperator(A)
    elif isinstance(A, scipylinalg.LinearOperator):
        if not hasattr(A, "dtype"):
            raise ArgumentError("scipy LinearOperator has no dtype.")
        ret = LinearOperator(A.shape, dot=A.matvec, dot_adj=A.rmatvec, dtype=A.dtype)
    else:
        raise TypeError("type not understood")

    if A is not None and not isinstance(A, IdentityLinearOperator) and timer is not None:
        ret = TimedLinearOperator(ret, timer)

    if tuple(shape) != tuple(ret.shape):
        raise LinearOperatorError("shape mismatch")
    return ret


# --------------------------------------------------------------------------
# inner products and norms -- krypy/utils.py:146-238
# --------------------------------------------------------------------------
def _is_identity_ip(ip_B):
    return ip_B is None or isinstance(ip_B, IdentityLinearOperator)


def _inner_dev(Xd, Yd, ip_B=None, out=None):
    """<X, Y> for device blocks X (m, N), Y (n, N) -> device (m, n) tensor, fp64 (complex128 for
    complex blocks) (krypy/utils.py:160-193).  Column j of the result is one kry_block_dot; for
    complex blocks two: Re = <X, y>_R and Im = -<X, i y>_R on the interleaved real views."""
    ctx = _ctx()
    t = _device.torch()
    m, n = Xd.shape[0], Yd.shape[0]
    cplx = _is_cplx(Xd) or _is_cplx(Yd)
    if cplx and not (_is_cplx(Xd) and _is_cplx(Yd)):
        Xd = Xd if _is_cplx(Xd) else Xd.to(t.complex128)
        Yd = Yd if _is_cplx(Yd) else Yd.to(t.complex128)
    if out is None or cplx:
        out = ctx.scalars(max(m * n, 1))[: m * n].reshape(n, m)
    if m == 0 or n == 0:
        return out.t().to(t.complex128) if cplx else out.t()
    if not _is_identity_ip(ip_B):
        try:
            B = get_linearoperator((Xd.shape[1], Xd.shape[1]), ip_B)
        except TypeError:
            # callable inner product on host arrays (utils.py:186-189)
            G = numpy.asarray(ip_B(ctx.to_numpy(Xd), ctx.to_numpy(Yd)))
            if numpy.iscomplexobj(G) and not cplx:
                if numpy.abs(G.imag).max() > 0:
                    raise NotImplementedError("complex-valued inner product of real blocks")
                G = G.real
            gdt = numpy.complex128 if cplx else numpy.float64
            return t.from_numpy(numpy.ascontiguousarray(G, dtype=gdt)).to(ctx.device)
        if m > n:
            # (B X)^H Y: for a self-adjoint B this equals X^H (B Y) up to round-off;
            # follow the reference's choice of which side gets B (utils.py:190-193)
            Xd = B._apply_dev(Xd)
        else:
            Yd = B._apply_dev(Yd)
    if not cplx:
        if (m > 1 and n > 1 and ctx.comm is None and Xd.dtype == Yd.dtype and Xd.shape[1] >= _BLOCK_MIN_N
                and ctx.gram_fits(m, n, Xd.data_ptr() == Yd.data_ptr() and m == n)):
            # block x block: ONE pass over both blocks (kry_gram) instead of n passes over X
            G = ctx.scalars(m * n)
            ctx.gram(Xd, m, Yd, n, G)
            return G.reshape(m, n)
        for j in range(n):
            ctx.block_dot(Xd, m, Yd[j], out[j])
        return out.t()
    out_im = ctx.scalars(m * n).reshape(n, m)
    iy = ctx.empty((1, Yd.shape[1]), Yd.dtype)
    for j in range(n):
        ctx.block_dot(Xd, m, Yd[j], out[j])
        ctx.rot90(Yd[j], iy[0])
        ctx.block_dot(Xd, m, iy[0], out_im[j])
    return t.complex(out, -out_im).t()


def _ip_coef(Xd, Yd, ip_B, out, acc=None, post=0, x_twin=None):
    """out[0] = <x, y>_B for single-vector device blocks (1, N) without leaving the
    device (post=1: sqrt(|.|) as numpy.sqrt(numpy.linalg.norm(ip, 2)) gives for a
    1x1 matrix, utils.py:238); acc[0] += out[0] when given.

    Complex blocks: out[0] is the REAL part (all a norm, CG and Lanczos need, linsys.py:634-641,
    utils.py:1003-1009); with ``x_twin`` (the real row ``i x`` of x's twin storage) the imaginary
    part ``<i x, y>_R`` goes to out[1] (acc[1])."""
    ctx = _ctx()
    want_im = x_twin is not None and _is_cplx(Xd) and not post
    if _is_identity_ip(ip_B):
        ctx.block_dot(Xd, 1, Yd[0], out, post, acc)
        if want_im:
            ctx.block_dot(x_twin.reshape(1, -1), 1, Yd[0], out[1:], 0, None if acc is None else acc[1:])
        return
    try:
        B = get_linearoperator((Xd.shape[1], Xd.shape[1]), ip_B)
    except TypeError:
        val = numpy.asarray(ip_B(ctx.to_numpy(Xd), ctx.to_numpy(Yd)))[0, 0]
        im = float(numpy.imag(val))
        if numpy.iscomplexobj(val) and not _is_cplx(Xd):
            if abs(val.imag) > 1e-10 * max(abs(val), 1e-300):
                raise NotImplementedError("complex-valued inner product of real blocks")
        val = float(numpy.real(val))
        if post:
            val = float(numpy.sqrt(abs(val)))
        out[0:1].fill_(val)
        if acc is not None:
            acc[0:1].add_(val)
        if want_im:
            out[1:2].fill_(im)
            if acc is not None:
                acc[1:2].add_(im)
        return
    BY = B._apply_dev(Yd)
    ctx.block_dot(Xd, 1, BY[0], out, post, acc)
    if want_im:
        ctx.block_dot(x_twin.reshape(1, -1), 1, BY[0], out[1:], 0, None if acc is None else acc[1:])


def ip_euclid(X, Y):
    """Euclidean inner product X^* Y (krypy/utils.py:146-157)."""
    return inner(X, Y)


def inner(X, Y, ip_B=None):
    """krypy/utils.py:160-193 for numpy ``(N,m)``, ``(N,n)`` inputs -> numpy ``(m,n)``."""
    if _is_dev(X) or _is_dev(Y):
        return _inner_dev(X, Y, ip_B)
    ctx = _ctx()
    X = numpy.asarray(X)
    Y = numpy.asarray(Y)
    dt = _compute_dtype(_common_type([X.dtype, Y.dtype, getattr(ip_B, "dtype", None)]))
    res = _inner_dev(ctx.to_block(X, dt), ctx.to_block(Y, dt), ip_B)
    return res.cpu().numpy().copy()


def norm_squared(x, Mx=None, inner_product=ip_euclid):
    """krypy/utils.py:196-211."""
    assert len(x.shape) == 2
    if Mx is None:
        rho = inner_product(x, x)
    else:
        assert len(Mx.shape) == 2
        rho = inner_product(x, Mx)
    if rho.shape == (1, 1):
        if abs(rho[0, 0].imag) > abs(rho[0, 0]) * 1e-10 or rho[0, 0].real < 0.0:
            raise InnerProductError("<x,Mx> = %g. Is the inner product indefinite?" % rho[0, 0])
    return numpy.linalg.norm(rho, 2)


def norm(x, y=None, ip_B=None):
    r"""krypy/utils.py:214-238: :math:`\sqrt{\langle x,y\rangle}`."""
    if y is None:
        y = x
    ip = inner(x, y, ip_B=ip_B)
    if _is_dev(ip):
        ip = ip.cpu().numpy()
    if ip.size == 0:
        return 0.0
    nrm_diag = numpy.linalg.norm(numpy.diag(ip), 2)
    nrm_diag_imag = numpy.linalg.norm(numpy.imag(numpy.diag(ip)), 2)
    if nrm_diag_imag > nrm_diag * 1e-10:
        raise InnerProductError("inner product defined by ip_B not positive definite?")
    return numpy.sqrt(numpy.linalg.norm(ip, 2))


def orthonormality(V, ip_B=None):
    """krypy/utils.py:297-305."""
    return numpy.linalg.norm(numpy.eye(V.shape[1]) - inner(V, V, ip_B=ip_B), 2)


def arnoldi_res(A, V, H, ip_B=None):
    """krypy/utils.py:308-329."""
    N = V.shape[0]
    invariant = H.shape[0] == H.shape[1]
    A = get_linearoperator((N, N), A)
    if invariant:
        res = A * V - numpy.dot(V, H)
    else:
        res = A * V[:, :-1] - numpy.dot(V, H)
    return norm(res, ip_B=ip_B)


# --------------------------------------------------------------------------
# Givens / Householder -- krypy/utils.py:332-436
# --------------------------------------------------------------------------
def _drotg(a, b):
    """BLAS drotg (reference-BLAS 3.10 algorithm), the host twin of the device
    routine in csrc/kry_small.cu; returns (c, s)."""
    anorm, bnorm = abs(a), abs(b)
    if bnorm == 0.0:
        return 1.0, 0.0
    if anorm == 0.0:
        return 0.0, 1.0
    safmin, safmax = 2.2250738585072014e-308, 4.4942328371557898e+307
    scl = min(safmax, max(safmin, anorm, bnorm))
    sigma = numpy.copysign(1.0, a) if anorm > bnorm else numpy.copysign(1.0, b)
    r = sigma * (scl * numpy.sqrt((a / scl) ** 2 + (b / scl) ** 2))
    return a / r, b / r


def _zrotg(a, b):
    """BLAS zrotg (LAPACK 3.10 semantics), the host twin of kryc_zrotg in
    csrc/kry_small_core.h; returns (c real, s complex)."""
    a, b = complex(a), complex(b)
    if b == 0:
        return 1.0, 0j
    if a == 0:
        u = max(abs(b.real), abs(b.imag))
        bs = b / u
        return 0.0, numpy.conj(bs) / abs(bs)
    u = max(abs(a.real), abs(a.imag), abs(b.real), abs(b.imag))
    as_, bs = a / u, b / u
    f2 = as_.real ** 2 + as_.imag ** 2
    g2 = bs.real ** 2 + bs.imag ** 2
    if f2 == 0.0:
        return (abs(a) / u) / numpy.sqrt(g2), (a / abs(a)) * numpy.conj(bs) / numpy.sqrt(g2)
    h2 = f2 + g2
    return float(numpy.sqrt(f2 / h2)), numpy.conj(bs) * (as_ / numpy.sqrt(f2 * h2))


class Givens:
    """krypy/utils.py:405-436 for a (2,1) vector (drotg for real-valued input, zrotg otherwise).
    This 2x2 helper is scalar host arithmetic; inside the solvers the rotations are generated
    and applied on the device (kry_givens_update[_z] / kry_minres_recur)."""

    def __init__(self, x):
        if x.shape != (2, 1):
            raise ArgumentError("x is not a vector of shape (2,1)")
        a = x[0].item()
        b = x[1].item()
        if numpy.isreal(x).all():
            a = float(numpy.real(a))
            b = float(numpy.real(b))
            c, s = _drotg(a, b)
        else:
            c, s = _zrotg(a, b)
        self.c = c
        self.s = s
        self.r = c * a + s * b
        self.G = numpy.array([[c, s], [-numpy.conj(s), c]])

    def apply(self, x):
        return numpy.dot(self.G, x)


def _house_params(gamma, sigma, n):
    """Scalars of the Householder reflector of a vector with first entry ``gamma`` and
    ``sigma = ||x[1:]||`` (krypy/utils.py:349-377; Golub/Van Loan Alg. 5.1.1 + section 5.1.13):
    returns (v0, vscale, xnorm, alpha, beta) with ``v = [v0, x[1:]] / vscale``."""
    if n == 1 or sigma == 0:
        xnorm = numpy.abs(gamma)
        alpha = 1 if gamma == 0 else gamma / xnorm
        return 1.0, 1.0, xnorm, alpha, 0
    xnorm = numpy.sqrt(numpy.abs(gamma) ** 2 + sigma ** 2)
    if gamma == 0:
        v0, alpha = -sigma, 1
    else:
        v0 = gamma + gamma / numpy.abs(gamma) * xnorm
        alpha = -gamma / numpy.abs(gamma)
    return v0, numpy.sqrt(numpy.abs(v0) ** 2 + sigma ** 2), xnorm, alpha, 2


class House:
    """krypy/utils.py:332-402: Householder reflector ``H = I - beta v v^*`` with
    ``H x = alpha ||x|| e_1``.  A helper for small host vectors (numpy ``(n,1)``), like ``Givens``;
    inside ``Arnoldi(ortho='house')`` the reflectors live in HBM and are applied by the block
    kernels (``Arnoldi._house_*``)."""

    def __init__(self, x):
        if len(x.shape) != 2 or x.shape[1] != 1:
            raise ArgumentError("x is not a vector of dim (N,1)")
        x = numpy.asarray(x)
        n = x.shape[0]
        gamma = x[0].item()
        sigma = numpy.linalg.norm(x[1:], 2) if n > 1 else 0
        v0, vscale, self.xnorm, self.alpha, self.beta = _house_params(gamma, sigma, n)
        v = numpy.array(x, dtype=numpy.result_type(x.dtype, type(v0), numpy.float64))
        v[0] = v0
        self.v = v / vscale

    def apply(self, x):
        """krypy/utils.py:379-389."""
        if len(x.shape) != 2:
            raise ArgumentError("x is not a matrix of shape (N,*)")
        if self.beta == 0:
            return x
        return x - self.beta * self.v * numpy.dot(self.v.T.conj(), x)

    def matrix(self):
        """krypy/utils.py:391-402 (dense; for tests)."""
        n = self.v.shape[0]
        return numpy.eye(n, n) - self.beta * numpy.dot(self.v, self.v.T.conj())


# --------------------------------------------------------------------------
# QR with inner product, Projection -- krypy/utils.py:439-707
# --------------------------------------------------------------------------
def _qr_dev(Xd, ip_B=None, reorthos=1):
    """Modified Gram-Schmidt QR of a device block (k, N) in the ``ip_B`` inner
    product (krypy/utils.py:695-707).  Returns (Q (k, N) device, R numpy (k,k))."""
    ctx = _ctx()
    k, N = Xd.shape
    if _is_cplx(Xd):
        return _qr_dev_z(ctx, Xd, ip_B, reorthos)
    if (_CHOLQR and 2 <= k <= 32 and N >= _BLOCK_MIN_N and ctx.comm is None and _is_identity_ip(ip_B)
            and ctx.gram_fits(k, k, True)):
        res = _cholqr2(ctx, Xd)
        if res is not None:
            return res
    ld = (N + 31) // 32 * 32
    store = ctx.empty((max(k, 1), ld), Xd.dtype)
    Q = store[:k, :N]
    Q.copy_(Xd)
    Rdev = ctx.scalars(max(k * (k + 1), 1))          # row i: R[:, i] (column i) + norm slot
    Rrows = Rdev[: k * (k + 1)].reshape(k, k + 1) if k else None
    euclid = _is_identity_ip(ip_B)
    tmp = ctx.scalars(2)
    R = numpy.zeros((k, k))
    for i in range(k):
        qi = Q[i]
        if euclid:
            ctx.orth_fused(Q, Q, 0, i, qi, reorthos + 1, KRY_ORTH_MGS, Rrows[i], nrm=Rrows[i][i:])
        else:
            for _ in range(reorthos + 1):
                for j in range(i):
                    _ip_coef(Q[j:j + 1], Q[i:i + 1], ip_B, tmp, acc=Rrows[i][j:])
                    ctx.axpy_dev(tmp, -1.0, Q[j], qi)
            _ip_coef(Q[i:i + 1], Q[i:i + 1], ip_B, Rrows[i][i:], post=1)
        ctx.sync()
        col = Rrows[i][: i + 1].cpu().numpy().copy()
        R[: i + 1, i] = col
        if R[i, i] >= 1e-15:                                        # utils.py:705-706
            ctx.scale_dev(Rrows[i][i:], 1, 1.0, qi, qi)
    return Q, R


_CHOLQR = __import__("os").environ.get("KRY_CHOLQR", "1") not in ("0", "")
# the one-pass block kernels (kry_gram, CholQR2) take over from the vector-by-vector sequences at
# sizes where a pass over the block costs more than a launch; below, the r1-validated paths stay
_BLOCK_MIN_N = int(__import__("os").environ.get("KRY_BLOCK_MIN_N", "4096"))


def _cholqr2(ctx, Xd):
    """Orthonormalise the (k, N) device block by two rounds of Cholesky QR (Euclidean inner product):
    G = X^H X in one pass (kry_gram), R = chol(G) on the host, X <- X R^-1 (kry_block_trsm).  Six
    block passes and two host synchronisations instead of the k(k+1) dependent sweeps of the
    column-by-column Gram-Schmidt (krypy/utils.py:698-706; 67 % of the reference's deflated solve,
    SURVEY F4).  Same factorisation as MGS up to round-off (R has a positive diagonal).  Returns None
    -- the caller falls back to the reference's MGS, which also implements its rank-deficiency rule
    (utils.py:705-706) -- when the block is numerically rank deficient or too ill-conditioned for
    CholQR2 (kappa(X) >~ 1e6)."""
    t = _device.torch()
    k, N = Xd.shape
    ld = (N + 31) // 32 * 32
    store = ctx.empty((k, ld), Xd.dtype)
    Q = store[:, :N]
    Q.copy_(Xd)
    G = ctx.scalars(k * k)
    R = numpy.eye(k)
    for rnd in range(2):
        ctx.gram(Q, k, Q, k, G)
        Gh = G.cpu().numpy().reshape(k, k).copy()               # synchronises
        Gh = 0.5 * (Gh + Gh.T)
        if not numpy.all(numpy.isfinite(Gh)):
            return None
        try:
            Rr = numpy.linalg.cholesky(Gh).T
        except numpy.linalg.LinAlgError:
            return None
        dg = numpy.diag(Rr)
        if dg.min() <= (1e-6 if rnd == 0 else 0.5) * dg.max():
            return None
        ctx.block_trsm(Q, k, t.from_numpy(numpy.ascontiguousarray(Rr)).to(ctx.device), Q)
        R = Rr.dot(R)
    return Q, R


def _qr_dev_z(ctx, Xd, ip_B, reorthos):
    """complex ``_qr_dev``: the same MGS over twin storage (column i against the 2i real rows
    ``q_0, i q_0, ...`` = complex MGS against i columns; coefficients come out interleaved)."""
    k, N = Xd.shape
    tw = _Twin.of(ctx, Xd)
    Q, T = tw.C, tw.T
    Rdev = ctx.scalars(max(2 * k * (k + 1), 2))
    Rrows = Rdev[: 2 * k * (k + 1)].reshape(k, 2 * (k + 1)) if k else None
    euclid = _is_identity_ip(ip_B)
    tmp = ctx.scalars(2)
    R = numpy.zeros((k, k), dtype=numpy.complex128)
    for i in range(k):
        qi = T[2 * i]
        if euclid:
            ctx.orth_fused(T, T, 0, 2 * i, qi, reorthos + 1, KRY_ORTH_MGS, Rrows[i], nrm=Rrows[i][2 * i:])
        else:
            for _ in range(reorthos + 1):
                for j in range(i):
                    _ip_coef(Q[j:j + 1], Q[i:i + 1], ip_B, tmp, acc=Rrows[i][2 * j:], x_twin=T[2 * j + 1])
                    ctx.axpy_dev(tmp, -1.0, T[2 * j], qi)
                    ctx.axpy_dev(tmp[1:], -1.0, T[2 * j + 1], qi)
            _ip_coef(Q[i:i + 1], Q[i:i + 1], ip_B, Rrows[i][2 * i:], post=1)
        ctx.sync()
        col = _cplx.from_pairs(Rrows[i][: 2 * (i + 1)].cpu().numpy().reshape(1, -1))[0]
        R[: i + 1, i] = col
        if R[i, i].real >= 1e-15:                                   # utils.py:705-706
            ctx.scale_dev(Rrows[i][2 * i:], 1, 1.0, qi, qi)
        tw.refresh(ctx, i)
    return Q, R


def qr(X, ip_B=None, reorthos=1):
    """krypy/utils.py:680-707.  Always (re-orthogonalised) modified Gram-Schmidt
    on the device; the reference's LAPACK shortcut for ``ip_B is None`` (:692-693)
    differs only by column signs (R has a positive diagonal here)."""
    if _is_dev(X):
        return _qr_dev(X, ip_B, reorthos)
    ctx = _ctx()
    X = numpy.asarray(X)
    dt = _compute_dtype(_common_type([X.dtype]))
    if X.shape[1] == 0:
        return X.copy(), numpy.zeros((0, 0), dtype=X.dtype)
    Qd, R = _qr_dev(ctx.to_block(X, dt), ip_B, reorthos)
    return ctx.to_numpy(Qd), R.astype(_common_type([X.dtype, numpy.float64]) if X.dtype.kind not in "fc" else X.dtype)


class Projection(object):
    """krypy/utils.py:439-677: oblique projection in XQRY form, bases in HBM.

    ``V, W`` are exposed as ``(N, k)`` numpy arrays (lazy D2H); ``apply*`` accept
    numpy ``(N, m)`` arrays or device blocks.  ``apply_complement`` on a single
    vector with a Euclidean inner product is ONE fused cooperative kernel
    (kry_project)."""

    def __init__(self, X, Y=None, ip_B=None, orthogonalize=True, iterations=2):
        import scipy.linalg
        self.ip_B = ip_B
        if iterations < 1:
            raise ArgumentError("iterations < 1 not allowed")
        self.orthogonalize = orthogonalize
        self.iterations = iterations
        ctx = _ctx()
        Y = X if Y is None else Y
        if len(X.shape) != 2:
            raise ArgumentError("X does not have shape==(N,k)")
        Xs = tuple(X.shape) if not _is_dev(X) else (X.shape[1], X.shape[0])
        Ys = tuple(Y.shape) if not _is_dev(Y) else (Y.shape[1], Y.shape[0])
        if Xs != Ys:
            raise ArgumentError("X and Y have different shapes")
        self._N, self._k = Xs
        same = Y is X
        if not _is_dev(X):
            dt = _compute_dtype(_common_type([X.dtype, Y.dtype]))
            Xd = ctx.to_block(X, dt)
            Yd = Xd if same else ctx.to_block(Y, dt)
        else:
            Xd, Yd = X, Y
        self._tdtype = Xd.dtype
        self._Q_dev = self._R_dev = None
        if self._k == 0:
            self._Vd = self._Wd = Xd
            self.VR = self.WR = self.Q = self.R = None
            return
        if orthogonalize:
            self._Vd, self.VR = _qr_dev(Xd, ip_B=ip_B)
        else:
            self._Vd, self.VR = Xd, None
        if same and orthogonalize:
            self._Wd, self.WR = self._Vd, self.VR
            self.Q, self.R = None, None
        else:
            if orthogonalize:
                self._Wd, self.WR = _qr_dev(Yd, ip_B=ip_B)
            else:
                self._Wd, self.WR = Yd, None
            M = _inner_dev(self._Wd, self._Vd, ip_B).cpu().numpy()
            self.Q, self.R = scipy.linalg.qr(M)           # k x k host algebra (utils.py:520)
            t = _device.torch()
            if _is_cplx(Xd):
                # kry_project computes R_dev^-1 (Q_dev^T c) on the 2k interleaved coefficients; the
                # embedding of a triangular complex R is only BLOCK triangular, so the kernel gets
                # the embedded product  T = R^-1 Q^H  as its "Q^T" and the identity as its "R"
                Tm = scipy.linalg.solve_triangular(self.R, self.Q.T.conj())
                self._Q_dev = t.from_numpy(numpy.ascontiguousarray(_cplx.expand_dense(Tm).T)).to(ctx.device)
                self._R_dev = t.from_numpy(numpy.eye(2 * self._k)).to(ctx.device)
            else:
                self._Q_dev = t.from_numpy(numpy.ascontiguousarray(self.Q, dtype=numpy.float64)).to(ctx.device)
                self._R_dev = t.from_numpy(numpy.ascontiguousarray(self.R, dtype=numpy.float64)).to(ctx.device)

    @property
    def V(self):
        return _ctx().to_numpy(self._Vd) if self._k else numpy.zeros((self._N, 0))

    @property
    def W(self):
        return _ctx().to_numpy(self._Wd) if self._k else numpy.zeros((self._N, 0))

    # -- device implementation ------------------------------------------------
    def _coef_transform(self, c_host, adj=False):
        import scipy.linalg
        if self.Q is not None and self.R is not None:
            if adj:
                return self.Q.dot(scipy.linalg.solve_triangular(self.R.T.conj(), c_host, lower=True))
            return scipy.linalg.solve_triangular(self.R, self.Q.T.conj().dot(c_host))
        return c_host

    def _apply_dev(self, ad, return_Ya=False, adj=False, c_first=None):
        """single application on a device block (m, N) -> Pa (m, N) [, Ya numpy (k, m)]
        (krypy/utils.py:522-564).  ``c_first`` (device, k doubles per column) receives the raw
        ``<W, a>`` like the fused kernel's ``c_first_dev``."""
        ctx = _ctx()
        m = ad.shape[0]
        Wd, Vd = (self._Vd, self._Wd) if adj else (self._Wd, self._Vd)
        c = _inner_dev(Wd, ad, self.ip_B).cpu().numpy().copy()   # (k, m), small
        if c_first is not None:
            cf = _coefs_dev(ctx, c.T.reshape(-1))                # complex: interleaved, 2k per column
            c_first[: cf.numel()].copy_(cf)
        Ya = None
        if return_Ya:
            Ya = c.copy()
            if self.WR is not None:
                Ya = self.WR.T.conj().dot(Ya)
        c = self._coef_transform(c, adj=adj)                     # (k, m)
        Pa = ctx.empty(ad.shape, ad.dtype)
        for j in range(m):
            _combine(ctx, Vd, self._k, c[:, j], None, Pa[j])
        return (Pa, Ya) if return_Ya else Pa

    def _complement_dev(self, ad, return_Ya=False, c_first=None, out=None):
        """apply_complement on a device block (krypy/utils.py:604-627).  Fused
        single-kernel path for Euclidean inner products; ``c_first`` (device, k
        doubles per column) receives the raw W^H a of the first application."""
        ctx = _ctx()
        if out is None:
            out = ad.clone()
        elif out.data_ptr() != ad.data_ptr():
            out.copy_(ad)
        if self._k == 0:
            return (out, numpy.zeros((0, ad.shape[0]))) if return_Ya else out
        m = ad.shape[0]
        cplx = _is_cplx(ad)
        kk = 2 * self._k if cplx else self._k            # real rows / coefficients per vector
        if _is_identity_ip(self.ip_B) and kk <= _CGS_CHUNK:
            own = c_first is None and return_Ya
            if own:
                c_first = ctx.scalars(kk * m)
            if cplx:
                Wb, Vb = _twin_for(ctx, self._Wd).T, _twin_for(ctx, self._Vd).T
            else:
                Wb, Vb = self._Wd, self._Vd
            for j in range(m):
                cf = None if c_first is None else c_first[j * kk:(j + 1) * kk]
                ctx.project(Wb, Vb, kk, out[j], self._Q_dev, self._R_dev, self.iterations, cf)
            if return_Ya:
                Ya = c_first[: kk * m].reshape(m, kk).cpu().numpy()
                Ya = (_cplx.from_pairs(Ya) if cplx else Ya).T
                if self.WR is not None:
                    Ya = self.WR.T.conj().dot(Ya)
                return out, Ya
            return out
        # generic inner product: one application at a time
        res = self._apply_dev(out, return_Ya=return_Ya, c_first=c_first)
        x, Ya = res if return_Ya else (res, None)
        ctx.axpby(1.0, out, -1.0, x, out)
        for _ in range(self.iterations - 1):
            w = self._apply_dev(out)
            ctx.axpby(1.0, out, -1.0, w, out)
        return (out, Ya) if return_Ya else out

    # -- public numpy API -----------------------------------------------------
    def _to_dev(self, a):
        if _is_dev(a):
            return a, True
        a = numpy.asarray(a)
        return _ctx().to_block(a, self._tdtype), False

    def _ret(self, xd, was_dev):
        return xd if was_dev else _ctx().to_numpy(xd)

    def _apply(self, a, return_Ya=False):
        """krypy/utils.py:522-552."""
        if self._k == 0:
            Pa = numpy.zeros(a.shape)
            if return_Ya:
                return Pa, numpy.zeros((0, a.shape[1]))
            return Pa
        ad, was_dev = self._to_dev(a)
        res = self._apply_dev(ad, return_Ya=return_Ya)
        if return_Ya:
            return self._ret(res[0], was_dev), res[1]
        return self._ret(res, was_dev)

    def _apply_adj(self, a):
        """krypy/utils.py:554-564."""
        if self._k == 0:
            return numpy.zeros(a.shape)
        ad, was_dev = self._to_dev(a)
        return self._ret(self._apply_dev(ad, adj=True), was_dev)

    def apply(self, a, return_Ya=False):
        """krypy/utils.py:566-591."""
        if self._k == 0:
            Pa = numpy.zeros(a.shape)
            if return_Ya:
                return Pa, numpy.zeros((0, a.shape[1]))
            return Pa
        ctx = _ctx()
        ad, was_dev = self._to_dev(a)
        res = self._apply_dev(ad, return_Ya=return_Ya)
        x, Ya = res if return_Ya else (res, None)
        for _ in range(self.iterations - 1):
            z = ctx.empty(ad.shape, ad.dtype)
            ctx.axpby(1.0, ad, -1.0, x, z)
            w = self._apply_dev(z)
            ctx.axpby(1.0, x, 1.0, w, x)
        if return_Ya:
            return self._ret(x, was_dev), Ya
        return self._ret(x, was_dev)

    def apply_adj(self, a):
        """krypy/utils.py:593-602."""
        if self._k == 0:
            return numpy.zeros(a.shape)
        ctx = _ctx()
        ad, was_dev = self._to_dev(a)
        x = self._apply_dev(ad, adj=True)
        for _ in range(self.iterations - 1):
            z = ctx.empty(ad.shape, ad.dtype)
            ctx.axpby(1.0, ad, -1.0, x, z)
            w = self._apply_dev(z, adj=True)
            ctx.axpby(1.0, x, 1.0, w, x)
        return self._ret(x, was_dev)

    def apply_complement(self, a, return_Ya=False):
        """krypy/utils.py:604-627."""
        if self._k == 0:
            if return_Ya:
                return a.copy() if not _is_dev(a) else a.clone(), numpy.zeros((0, a.shape[1]))
            return a.copy() if not _is_dev(a) else a.clone()
        ad, was_dev = self._to_dev(a)
        res = self._complement_dev(ad, return_Ya=return_Ya)
        if return_Ya:
            return self._ret(res[0], was_dev), res[1]
        return self._ret(res, was_dev)

    def apply_complement_adj(self, a):
        """krypy/utils.py:629-638."""
        if self._k == 0:
            return a.copy()
        ctx = _ctx()
        ad, was_dev = self._to_dev(a)
        x = self._apply_dev(ad, adj=True)
        z = ctx.empty(ad.shape, ad.dtype)
        ctx.axpby(1.0, ad, -1.0, x, z)
        for _ in range(self.iterations - 1):
            w = self._apply_dev(z, adj=True)
            ctx.axpby(1.0, z, -1.0, w, z)
        return self._ret(z, was_dev)

    def _get_operator(self, fun, fun_adj):
        N = self._N
        return LinearOperator((N, N), _device.torch_to_np_dtype(self._tdtype), fun, fun_adj)

    def operator(self):
        """krypy/utils.py:645-654."""
        if self._k == 0:
            return ZeroLinearOperator((self._N, self._N))
        return self._get_operator(self.apply, self.apply_adj)

    def operator_complement(self):
        """krypy/utils.py:656-665."""
        if self._k == 0:
            return IdentityLinearOperator((self._N, self._N))
        return self._get_operator(self.apply_complement, self.apply_complement_adj)

    def matrix(self):
        """krypy/utils.py:667-677."""
        return self.apply(numpy.eye(self._N))


# --------------------------------------------------------------------------
# Arnoldi / Lanczos -- krypy/utils.py:854-1081
# --------------------------------------------------------------------------
def _rightmost_factor(op):
    """the factor of an operator product that is applied to the basis vector first"""
    while isinstance(op, _ProductLinearOperator):
        op = op.args[1]
    return op


_ORTHO = {
    # name: (kernel algorithm, passes)
    "mgs": (KRY_ORTH_MGS, 1),        # utils.py:923-926: reorthos = 0
    "dmgs": (KRY_ORTH_MGS, 2),       # reorthos = 1
    "lanczos": (KRY_ORTH_MGS, 1),    # start = k: a single basis vector, utils.py:1000-1001
    "cgs": (KRY_ORTH_CGS, 1),        # new: fused block classical Gram-Schmidt
    "cgs2": (KRY_ORTH_CGS, 2),       # new: CGS with re-orthogonalisation
}
_CGS_CHUNK = 64   # KRY_MAX_SLOTS of csrc/kry_common.cuh
# the fused one-kernel Lanczos step for a diagonal inner-product matrix (kry_lanczos_diag) is the
# default since round 2 (validated on B200: +30 % on config C5); KRY_LANCZOS_DIAGB=0 selects the
# generic seven-launch sequence
_LANCZOS_DIAGB = __import__("os").environ.get("KRY_LANCZOS_DIAGB", "1") not in ("0", "")
# native complex128 kernels of the Arnoldi hot loop (csrc/kry_cplx.cu: kry_orth_fused_z reads every basis vector
# once instead of the vector and its twin; kry_spmv_csr_z moves 20 instead of 48 bytes per matrix entry);
# KRY_NATIVE_Z=0 selects the real-embedding kernels for these two as well (krypy_b200/_cplx.py)
_NATIVE_Z = __import__("os").environ.get("KRY_NATIVE_Z", "1") not in ("0", "")
_CGS_CHUNK_Z = 32   # complex vectors per kry_orth_fused_z call (two reduction slots each)
# L2 residency window on w = A v_k, the vector an Arnoldi step reads four times and writes twice (kry_l2_window):
# its passes after the first are served by the 126 MB L2 instead of HBM.  Single GPU, vectors of 8 MB and more.
# Default since round 2 -- same-box A/B on a B200 (profiles/r2_l2window_ab.json): exact MGS (the drop-in default
# ortho) 980 -> 1,191 it/s on C2, C5 2,282 -> 2,401 it/s, block CGS 1,605 -> 1,623 it/s, results bitwise identical.
# KRY_L2_WINDOW=0 turns it off.
_L2_WINDOW = __import__("os").environ.get("KRY_L2_WINDOW", "1") not in ("0", "")
_L2_WINDOW_MIN_BYTES = 1 << 23


class DeviceBlock(object):
    """A block of k vectors that already lives in HBM in the internal vector-major layout
    ``(k, N)``; accepted wherever the public API takes an ``(N, k)`` array of deflation vectors."""

    def __init__(self, block):
        self.block = block

    @property
    def shape(self):
        return (self.block.shape[1], self.block.shape[0])

    @property
    def dtype(self):
        return _device.torch_to_np_dtype(self.block.dtype)


class SolverWorkspace(object):
    """Device buffers (and the CUDA graphs recorded over them) that persist across the restart
    cycles of one restarted solve.  The reference re-allocates everything per cycle
    (linsys.py:1046-1051); keeping the buffers makes every Arnoldi step's launch sequence
    identical from cycle to cycle, so step k is captured once into a CUDA graph and replayed
    with a single launch afterwards (the host-bound regime of the multi-GPU configurations)."""

    def __init__(self, graphs=None):
        import os
        self.bufs = {}
        self.graphs = {}
        self.uses = 0
        if graphs is None:
            graphs = os.environ.get("KRY_GRAPHS", "auto")
        self.graph_mode = graphs          # "auto" (row-partitioned runs only) | "on" | "off"

    def tensor(self, name, key, factory):
        k = (name,) + tuple(key)
        t = self.bufs.get(k)
        if t is None:
            t = self.bufs[k] = factory()
        return t

    def graphs_enabled(self, ctx):
        if ctx.timer is not None or self.graph_mode == "off":
            return False
        if self.graph_mode == "on":
            return True
        return ctx.comm is not None


class Arnoldi(object):
    """krypy/utils.py:854-1074 with the basis resident in HBM.

    ``V`` (and ``P`` when ``M`` is given) are stored vector-major as
    ``(maxiter+1, N)`` device tensors; the ``V``/``P``/``H`` attributes and
    ``get()`` return ``(N, k)`` numpy arrays like the reference.  One step is:
    operator apply (kry_spmv_csr / kry_gemv_dense), ONE fused cooperative
    Gram-Schmidt kernel (kry_orth_fused: dots, update, norm, normalised store),
    and -- when driven by Gmres/Minres -- the small device recurrence.

    ``ortho``: 'mgs', 'dmgs', 'lanczos' as in the reference (exact MGS order),
    plus 'cgs' / 'cgs2' (block classical Gram-Schmidt, one / two passes: V is
    read twice per pass with a single grid-wide reduction).  'house' is not
    implemented.
    """

    def __init__(self, A, v, maxiter=None, ortho="mgs", M=None, Mv=None, Mv_norm=None, ip_B=None,
                 dtype=None, _workspace=None, _prelaunched=False):
        """``_prelaunched`` (linsys.Gmres over a workspace): the steps of this Arnoldi process are already
        running on the device (the previous restart cycle wrote v_0 and launched them); only the host side is
        set up here, no device buffer is touched."""
        ctx = self._ctx = _ctx()
        ws = self._ws = _workspace
        t = _device.torch()
        v_dev_in = _is_dev(v)
        N = v.shape[1] if v_dev_in else v.shape[0]
        self.N = N
        self.A = get_linearoperator((N, N), A)
        self.maxiter = N if maxiter is None else maxiter
        self.ortho = ortho
        self.M = get_linearoperator((N, N), M)
        if isinstance(self.M, IdentityLinearOperator):
            self.M = None
        self.ip_B = ip_B
        self.dtype = _common_type([find_common_dtype(self.A, v, self.M), dtype])   # utils.py:898
        if ortho == "house":
            if self.M is not None or not _is_identity_ip(ip_B):
                raise ArgumentError("Only euclidean inner product allowed with Householder orthogonalization")
            if ctx.comm is not None:
                raise NotImplementedError("ortho='house' is not implemented for row-partitioned runs")
        elif ortho not in _ORTHO:
            raise ArgumentError(
                "Invalid value '%s' for argument 'ortho'. Valid are house, mgs, dmgs, lanczos "
                "(and cgs, cgs2 on the device path)." % ortho)
        self._algo, self._passes = _ORTHO.get(ortho, (None, 0))
        td = self._td = _compute_dtype(self.dtype)
        self._cplx = cplx = td == t.complex128
        self._nr = nr = 2 if cplx else 1          # real rows per basis vector / doubles per coefficient
        self.iter = 0
        self.invariant = False
        m1 = self.maxiter + 1
        rf = _rightmost_factor(self.A)
        self._halo_op = rf if (ctx.comm is not None and hasattr(rf, "_halo_args")) else None
        if ctx.comm is not None:
            ctx.comm.halo_ready = None
        self._Pd = None
        if cplx:
            # twin storage: rows 2j, 2j+1 = v_j, i v_j; the kernels work on the real rows
            if ctx.comm is not None:
                raise NotImplementedError("complex row-partitioned runs are not implemented")
            if ws is not None:
                self._Vtw = ws.tensor("Vtw", (m1, N), lambda: _Twin(ctx, m1, N))
            else:
                self._Vtw = _Twin(ctx, m1, N)
            self._Vs = self._Vtw.store
            self._Vd, self._Vt = self._Vtw.C, self._Vtw.T
            self._Pt = None
            if self.M is not None:
                # (through the workspace like V: a CUDA graph recorded in one restart cycle is
                # replayed in the next ones and must find the same buffers)
                self._Ptw = (ws.tensor("Ptw", (m1, N), lambda: _Twin(ctx, m1, N)) if ws is not None
                             else _Twin(ctx, m1, N))
                self._Pd, self._Pt = self._Ptw.C, self._Ptw.T
        else:
            if ws is not None:
                self._Vs = ws.tensor("V", (m1, N, td), lambda: ctx.alloc_basis(m1, N, td, rf))
            else:
                self._Vs = ctx.alloc_basis(m1, N, td, rf)
            self._Vd = self._Vs[:, :N]
            self._Vt = self._Vd
            self._Pt = None
            if self.M is not None:
                self._Ps = (ws.tensor("P", (m1, N, td), lambda: ctx.alloc_basis(m1, N, td)) if ws is not None
                            else ctx.alloc_basis(m1, N, td))
                self._Pd = self._Ps[:, :N]
                self._Pt = self._Pd
        self._ld = self._Vs.stride(0)
        # small quantities are always >= fp64 on the device path (also in fp32 storage mode)
        self.H = numpy.zeros((self.maxiter + 1, self.maxiter), dtype=_common_type([self.dtype, numpy.float64]))
        self._euclid = _is_identity_ip(ip_B)
        # Row-partitioned runs: A v_k goes to one of TWO peer-mapped buffers (step parity), because the
        # neighbours gather the halo of v_{k+1} from this rank's un-normalised q (kry_dist_scale_haloq)
        # while this rank may already be writing the next A v.
        self._q2 = None
        if self._halo_op is not None and not cplx and getattr(ctx.comm, "halo_from_q", False):
            mk = lambda: ctx.alloc_basis(2, N, td, rf)
            q2 = ws.tensor("q2", (2, N, td), mk) if ws is not None else mk()
            self._q2 = (q2, q2[0:1, :N], q2[1:2, :N])
        if ws is not None:
            # buffers persist across restart cycles so that a step's launch arguments are stable
            # (CUDA-graph replay); everything that must start from zero is re-zeroed here
            self._q = ws.tensor("q", (1, N, td), lambda: ctx.empty((1, N), td))
            self._hcol_store = ws.tensor("hcol", (nr * (self.maxiter + 10),),
                                         lambda: ctx.scalars(nr * (self.maxiter + 2 + 8)))
            self._tmp = ws.tensor("tmp", (4,), lambda: ctx.scalars(4))
            self._lz = ws.tensor("lz", (3,), lambda: ctx.scalars(3))
            self._lz_st = ws.tensor("lz_st", (16,), lambda: ctx.scalars(16))
            if not _prelaunched:
                for buf in (self._hcol_store, self._tmp, self._lz, self._lz_st):
                    buf.zero_()
        else:
            self._q = ctx.empty((1, N), td)
            self._hcol_store = ctx.scalars(nr * (self.maxiter + 2 + 8))
            self._tmp = ctx.scalars(4)
            self._lz = ctx.scalars(3)          # Lanczos: [H[k-1,k], H[k,k], H[k+1,k]]
            self._lz_st = ctx.scalars(16)
        if self._euclid and self.M is None:
            self._t = None
        elif ws is not None:
            self._t = ws.tensor("t", (1, N, td), lambda: ctx.empty((1, N), td))
        else:
            self._t = ctx.empty((1, N), td)
        self._hcol = self._hcol_store[nr:]      # one leading zero: H[-1, 0] of linsys.py:828
        self._hfro2 = 0.0
        if _L2_WINDOW and ctx.comm is None and self._q.numel() * self._q.element_size() >= _L2_WINDOW_MIN_BYTES:
            ctx.l2_window(self._q)                  # (a no-op when the window already covers this buffer)

        if _prelaunched:
            self.vnorm = Mv_norm
            return
        # first basis vector: utils.py:923-952
        vd = v if v_dev_in else ctx.to_block(numpy.asarray(v), td)
        if vd.dtype != td:
            vd = vd.to(td)
        if ortho == "house":
            # utils.py:910-922: reflectors zero-padded to full length in HBM (twin storage if complex)
            self._Wtw = _Twin(ctx, m1 + 1, N) if cplx else None
            self._Wt = self._Wtw.T if cplx else ctx.alloc_basis(m1 + 1, N, td)[:, :N]
            self._houses = []
            self._house_make(0, vd)
            Mv_norm = self._houses[0][2]                       # numpy.linalg.norm(v, 2)
        if self.M is not None:
            pd = vd
            if Mv is None:
                vd = self.M._apply_dev(pd)
            else:
                vd = Mv if _is_dev(Mv) else ctx.to_block(numpy.asarray(Mv), td)
            if Mv_norm is None:
                self.vnorm = self._norm_dev(pd, vd)
            else:
                self.vnorm = Mv_norm
            if self.vnorm > 0:
                self._set_scaled(self._Pd[0], pd[0], self.vnorm)
                if cplx:
                    self._Ptw.refresh(ctx, 0)
        else:
            if Mv_norm is None:
                self.vnorm = self._norm_dev(vd, None)
            else:
                self.vnorm = Mv_norm
        if self.vnorm > 0:
            self._set_scaled(self._Vd[0], vd[0], self.vnorm)
            if cplx:
                self._Vtw.refresh(ctx, 0)
        else:
            self.invariant = True

    # -- helpers ---------------------------------------------------------------
    def _set_scaled(self, dst, src, s):
        """dst = src / s with a host scalar (utils.py:938, 950)."""
        self._tmp[3:4].fill_(float(s))
        self._ctx.scale_dev(self._tmp[3:], 1, 1.0, src, dst)

    def _norm_dev(self, xd, yd):
        """norm(x, y, ip_B) of device blocks (1, N): one sync (setup only)."""
        _ip_coef(xd, xd if yd is None else yd, self.ip_B, self._tmp, post=1)
        return float(self._tmp[0].item())

    # -- one step, enqueue only ---------------------------------------------------
    def _enqueue_dist_fused(self, k, givens):
        """Row-partitioned block-CGS step with ONE cross-GPU wait (see dist.PeerComm.fused_step): SpMV,
        kry_dist_dot with <w, w> (the local V^H w and <w, w>, w = A v_k, published to the peers) and
        kry_dist_update_scale (acquire, norm from <w, w> - sum c^2 with an exact-norm guard, update +
        normalised store in one sweep, halo of v_{k+1}, Givens update in an extra CTA).  Returns None when the
        step does not qualify, else whether the Givens update was part of it."""
        ctx = self._ctx
        comm, op = ctx.comm, self._halo_op
        if (comm is None or op is None or not getattr(comm, "fused_step", False) or comm.reduce != "peer"
                or self._q2 is None or self._algo != KRY_ORTH_CGS or self._passes != 1 or not self._euclid
                or self.M is not None or self._cplx or self.ortho in ("lanczos", "house") or k + 2 > 64):
            return None
        Vt = self._Vt
        vnext = Vt[k + 1]
        hal = op._halo_args(vnext)
        q = self._q2[1 + (k & 1)]
        halq = op._halo_src_args(q[0]) if hal is not None else None
        if halq is None:
            return None
        self.A._apply_dev(self._Vd[k:k + 1], out=q)                # utils.py:968
        ctx.dist_dot_sq(Vt, k + 1, q[0])
        fold = givens is not None
        ctx.dist_update_scale(Vt, k + 1, q[0], vnext, self._hcol.data_ptr(), self._hcol[k + 1:], hal, halq,
                              op.plan.block, givens=((k,) + tuple(givens)) if fold else None)
        comm.halo_ready = vnext.data_ptr()
        return fold

    def _enqueue(self, k, givens=None):
        """Launch the kernels of Arnoldi step k (utils.py:964-1045) without any
        host synchronisation.  Results: h[0..k] accumulated into self._hcol (or
        self._lz for Lanczos), H[k+1,k] in hcol[k+1] (lz[2]), V[k+1] (P[k+1]) stored.
        ``givens`` = (rcol, cs, y, mailbox offset): the caller's Givens update of column k may be folded
        into the step's last kernel; returns True when it was (the caller then skips its own launch)."""
        ctx = self._ctx
        if ctx.comm is not None:
            folded = self._enqueue_dist_fused(k, givens)
            if folded is not None:
                return folded
        V, P = self._Vd, self._Pd
        # Vt/Pt: the rows the kernels work on.  Real: the basis itself.  Complex: twin storage,
        # rows nr*j (+1) = v_j (i v_j); coefficients are nr doubles each (interleaved re/im).
        Vt, Pt, nr, cplx = self._Vt, self._Pt, self._nr, self._cplx
        q = self._q if self._q2 is None else self._q2[1 + (k & 1)]
        self.A._apply_dev(V[k:k + 1], out=q)                       # utils.py:968
        q0 = q[0]
        if self.ortho == "house":
            return self._enqueue_house(k)
        lanczos = self.ortho == "lanczos"
        Vsub = Pt if Pt is not None else Vt
        if lanczos:
            # three-term recurrence with REAL coefficients (alpha.real, utils.py:1003-1009): also for
            # complex data only the real row of v_k takes part
            r0, r1 = nr * k, nr * k + 1
            h_ptr = self._lz.data_ptr() + 8 * (1 - r0)              # h[k] -> lz[1]
            nrm = self._lz[2:]
            pre_vec = Vsub[nr * (k - 1)] if k > 0 else None
            pre_coef = self._lz if k > 0 else None
        else:
            r0, r1 = 0, nr * (k + 1)
            h_ptr = self._hcol.data_ptr()
            nrm = self._hcol[nr * (k + 1):]
            pre_vec = pre_coef = None
        vnext = Vt[nr * (k + 1)]
        if self._euclid and cplx and not lanczos and _NATIVE_Z and ctx.comm is None:
            # native complex sweep (kry_orth_fused_z) over the vectors themselves: the even rows of the twin
            # storage, one read per basis vector; complex coefficients interleaved in hcol as on the twin path
            fused_tail = self.M is None
            ldz = Vt.stride(0)               # complex elements between v_j and v_{j+1} (two real rows)
            j0, nvz = 0, k + 1
            while True:
                j1 = nvz if self._algo != KRY_ORTH_CGS else min(j0 + _CGS_CHUNK_Z, nvz)
                last = j1 == nvz
                ctx.orth_fused_z(Vt, Vsub, ldz, j0, j1, q0, self._passes, self._algo, h_ptr,
                                 nrm=nrm if (last and fused_tail) else None,
                                 vnext=vnext if (last and fused_tail) else None)
                if last:
                    break
                j0 = j1
        elif self._euclid:
            fused_tail = self.M is None
            if self._algo == KRY_ORTH_CGS and (r1 - r0) > _CGS_CHUNK:
                # more basis vectors than reduction slots: block-wise CGS
                j0 = r0
                while j0 < r1:
                    j1 = min(j0 + _CGS_CHUNK, r1)
                    last = j1 == r1
                    ctx.orth_fused(Vt, Vsub, j0, j1, q0, self._passes, self._algo, None,
                                   nrm=nrm if (last and fused_tail) else None,
                                   vnext=vnext if (last and fused_tail) else None, h_ptr=h_ptr)
                    j0 = j1
            else:
                ctx.orth_fused(Vt, Vsub, r0, r1, q0, self._passes, self._algo, None,
                               nrm=nrm if fused_tail else None,
                               vnext=vnext if fused_tail else None,
                               pre_vec=pre_vec, pre_coef=pre_coef, h_ptr=h_ptr, halo_op=self._halo_op)
        elif (_LANCZOS_DIAGB and lanczos and not cplx and self.M is None
              and (ctx.comm is None or ctx.comm.reduce == "peer")
              and self._passes == 1 and self._diag_ip() is not None):
            # the whole Lanczos step for a diagonal ip_B in ONE cooperative kernel instead of the seven
            # launches of the generic path below (KRY_LANCZOS_DIAGB=0 turns it off)
            ctx.lanczos_diag(pre_vec, Vt[k], self._diag_ip(), q0, pre_coef, self._lz, vnext)
            return
        else:
            # generic inner product: the reference's loop, one reduction at a time
            if pre_vec is not None:
                ctx.axpy_dev(pre_coef, -1.0, pre_vec, q0)
            for _ in range(self._passes):
                for j in range(k if lanczos else 0, k + 1):
                    hslot = self._lz[1:] if lanczos else self._hcol[nr * j:]
                    tw = Vt[nr * j + 1] if (cplx and not lanczos) else None
                    _ip_coef(V[j:j + 1], q, self.ip_B, self._tmp, acc=hslot, x_twin=tw)   # utils.py:1015, 1025
                    ctx.axpy_dev(self._tmp, -1.0, Vsub[nr * j], q0)                       # utils.py:1026-1029
                    if tw is not None:
                        ctx.axpy_dev(self._tmp[1:], -1.0, Vsub[nr * j + 1], q0)
            fused_tail = False
        if not fused_tail:
            # utils.py:1030-1045: M apply, norm, scaled stores
            if self.M is not None:
                MAv = self.M._apply_dev(q, out=self._t)
                _ip_coef(q, MAv, self.ip_B, nrm, post=1)
                ctx.scale_dev(nrm, 1, 1.0, q0, P[k + 1])
                ctx.scale_dev(nrm, 1, 1.0, MAv[0], V[k + 1])
                if cplx:
                    self._Ptw.refresh(ctx, k + 1)
            else:
                _ip_coef(q, q, self.ip_B, nrm, post=1)
                ctx.scale_dev(nrm, 1, 1.0, q0, V[k + 1])
        if cplx:
            self._Vtw.refresh(ctx, k + 1)

    def _diag_ip(self):
        """device vector of the diagonal of ip_B when it is a real positive diagonal operator, else None"""
        if "_diag_ip_cache" not in self.__dict__:
            d = None
            try:
                B = get_linearoperator((self.N, self.N), self.ip_B)
            except TypeError:
                B = None
            if isinstance(B, TimedLinearOperator):
                B = B._linear_operator
            if isinstance(B, DiagonalLinearOperator) and numpy.dtype(B._d.dtype).kind != "c" \
                    and B._d.shape[0] > 0 and float(B._d.min()) > 0.0:
                d = B._dev(self._td)
            self.__dict__["_diag_ip_cache"] = d
        return self.__dict__["_diag_ip_cache"]

    # -- Householder orthogonalisation (utils.py:970-994) ------------------------------------
    def _house_make(self, j, xd):
        """Reflector j from the device vector xd (1, N), acting on entries j..N-1 (utils.py:332-377):
        stored zero-padded in row j of the reflector block; (alpha, beta, xnorm) kept on the host.
        One host synchronisation (gamma and sigma are needed for the branch decisions)."""
        ctx, t = self._ctx, _device.torch()
        N, nr = self.N, self._nr
        w = self._Wtw.C[j] if self._cplx else self._Wt[j]
        w.copy_(xd[0])
        if j > 0:
            w[:j].zero_()
        gam = t.empty(1, dtype=w.dtype)
        gam.copy_(w[j:j + 1])                                      # D2H of one entry (synchronises)
        gamma = gam[0].item()
        sigma = 0.0
        if j + 1 < N:
            w[j:j + 1].zero_()
            ctx.block_dot(w.reshape(1, -1), 1, w, self._tmp, 1, None)      # ||x[j+1:]||
            sigma = float(self._tmp[0].item())
        v0, vscale, xnorm, alpha, beta = _house_params(gamma, sigma, N - j)
        w[j:j + 1].copy_(t.full((1,), v0, dtype=w.dtype))          # H2D of one entry
        self._tmp[3:4].fill_(float(vscale))
        ctx.scale_dev(self._tmp[3:], 1, 1.0, w, w)
        if self._cplx:
            self._Wtw.refresh(ctx, j)
        self._houses.append((alpha, beta, xnorm))

    def _house_apply(self, j, x):
        """x <- (I - beta_j w_j w_j^*) x on a full-length device vector (utils.py:379-389)"""
        beta = self._houses[j][1]
        if beta == 0:
            return
        ctx, nr = self._ctx, self._nr
        W = self._Wt[nr * j:]
        ctx.block_dot(W, nr, x, self._tmp, 0, None)               # <w_j, x> (complex: re, im over the twin rows)
        ctx.block_axpy(W, nr, self._tmp, -float(beta), x)

    def _enqueue_house(self, k):
        """Householder step k (utils.py:970-994); leaves column k of H in the device accumulator
        like the Gram-Schmidt kernels do and stores V[k+1]."""
        ctx, t = self._ctx, _device.torch()
        N, nr, cplx = self.N, self._nr, self._cplx
        q = self._q
        for j in range(k + 1):
            self._house_apply(j, q[0])
        ncoef = min(k + 2, N)
        col = numpy.zeros(k + 2, dtype=numpy.complex128 if cplx else numpy.float64)
        if k + 1 < N:
            self._house_make(k + 1, q)
            col[k + 1] = self._houses[k + 1][2]                    # |alpha^* (H_{k+1} q)[k+1]| = ||q[k+1:]||
        head = q[0][:min(k + 1, N)].cpu().numpy().astype(col.dtype)
        for j in range(head.shape[0]):
            col[j] = head[j] * numpy.conj(self._houses[j][0])      # Av[j] *= conj(alpha_j)
        self._hcol[: nr * (k + 2)].copy_(_coefs_dev(ctx, col))
        if k + 1 < N:
            # v_{k+1} = alpha_{k+1} H_0 ... H_{k+1} e_{k+1}
            v = self._t_house = getattr(self, "_t_house", None)
            if v is None:
                v = self._t_house = ctx.empty((1, N), self._td)
            v.zero_()
            v[0][k + 1:k + 2].copy_(t.ones(1, dtype=v.dtype))
            for j in range(k + 1, -1, -1):
                self._house_apply(j, v[0])
            _caxpby(ctx, self._houses[k + 1][0], v[0], 0.0, None, self._Vd[k + 1])
            if cplx:
                self._Vtw.refresh(ctx, k + 1)

    def _finish(self, k, hcol_host):
        """Host bookkeeping of step k given H[0..k+1, k] (utils.py:1025, 1032-1039, 1048)."""
        H = self.H
        if self.ortho == "lanczos":
            if k > 0:
                H[k - 1, k] = H[k, k - 1]                         # utils.py:1003
            H[k, k] = hcol_host[0]
            H[k + 1, k] = hcol_host[1]
            col2 = hcol_host[0] ** 2 + hcol_host[1] ** 2 + (abs(H[k - 1, k]) ** 2 if k > 0 else 0.0)
        else:
            H[: k + 2, k] = hcol_host[: k + 2]
            col2 = float(numpy.sum(numpy.abs(hcol_host[: k + 2]) ** 2))
        self._hfro2 += col2
        hk = float(numpy.real(H[k + 1, k]))
        # invariant-subspace test H[k+1,k]/||H[:k+2,:k+1]||_2 <= 1e-14 (utils.py:1035-1039).
        # ||H||_2 <= ||H||_F, so the SVD is only needed when the cheap bound cannot decide.
        if not numpy.isfinite(hk):
            self.invariant = True
        elif hk <= 1e-14 * numpy.sqrt(self._hfro2):
            nrm2 = numpy.linalg.norm(H[: k + 2, : k + 1], 2)
            if nrm2 == 0 or hk / nrm2 <= 1e-14:
                self.invariant = True
        self.iter = k + 1

    def advance(self):
        """Carry out one iteration of Arnoldi (krypy/utils.py:954-1048)."""
        if self.iter >= self.maxiter:
            raise ArgumentError("Maximum number of iterations reached.")
        if self.invariant:
            raise ArgumentError("Krylov subspace was found to be invariant in the previous iteration.")
        ctx = self._ctx
        k = self.iter
        self._enqueue(k)
        if self.ortho == "lanczos":
            ctx.minres_recur(k, self._lz, self._lz_st, 1, 0)        # publishes + shifts the 3 entries
            ctx.sync()
            self._finish(k, ctx.mailbox[6:8].copy())
        else:
            self._finish(k, self._read_hcol(k))

    def _read_hcol(self, k):
        """synchronising D2H of column k of H (k+2 coefficients); re-zeroes the accumulator"""
        n = self._nr * (k + 2)
        hc = self._hcol[:n].cpu().numpy().copy()
        self._hcol[:n].zero_()
        return _cplx.from_pairs(hc.reshape(1, -1))[0] if self._cplx else hc

    # -- accessors ------------------------------------------------------------------
    def _block_np(self, Bd, ncols):
        out = numpy.zeros((self.N, self.maxiter + 1), dtype=self.dtype)
        if ncols > 0:
            out[:, :ncols] = self._ctx.to_numpy(Bd[:ncols])
        return out

    def _valid_cols(self):
        return self.iter if self.invariant else self.iter + 1

    @property
    def V(self):
        return self._block_np(self._Vd, self._valid_cols())

    @property
    def P(self):
        if self._Pd is None:
            raise AttributeError("P")
        return self._block_np(self._Pd, self._valid_cols())

    def get(self):
        """krypy/utils.py:1050-1061."""
        k = self.iter
        ctx = self._ctx
        if self.invariant:
            V, H = ctx.to_numpy(self._Vd[:k]).astype(self.dtype, copy=False), self.H[:k, :k]
            if self.M is not None:
                return V, H, ctx.to_numpy(self._Pd[:k]).astype(self.dtype, copy=False)
            return V, H
        V, H = ctx.to_numpy(self._Vd[: k + 1]).astype(self.dtype, copy=False), self.H[: k + 1, :k]
        if self.M is not None:
            return V, H, ctx.to_numpy(self._Pd[: k + 1]).astype(self.dtype, copy=False)
        return V, H

    def get_last(self):
        """krypy/utils.py:1063-1074."""
        k = self.iter
        ctx = self._ctx
        if self.invariant:
            V, H = None, self.H[:k, [k - 1]]
            if self.M is not None:
                return V, H, None
            return V, H
        V, H = ctx.to_numpy(self._Vd[k:k + 1]).astype(self.dtype, copy=False), self.H[: k + 1, [k - 1]]
        if self.M is not None:
            return V, H, ctx.to_numpy(self._Pd[k:k + 1]).astype(self.dtype, copy=False)
        return V, H


def arnoldi(*args, **kwargs):
    """krypy/utils.py:1077-1081."""
    _arnoldi = Arnoldi(*args, **kwargs)
    try:
        while _arnoldi.iter < _arnoldi.maxiter and not _arnoldi.invariant:
            _arnoldi.advance()
    finally:
        _ctx().l2_window(None)        # the L2 set-aside of the process's vector goes back (an idle one costs L2)
    return _arnoldi.get()


# host-side analysis helpers under the reference's names
from ._analysis import (BoundCG, BoundMinres, Interval, Intervals, NormalizedRootsPolynomial,  # noqa: E402,F401
                        angles, arnoldi_projected, bound_perturbed_gmres, gap, get_residual_norms,
                        hegedus, norm_MMlr, ritz, strakos)
