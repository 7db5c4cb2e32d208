"""TEST DOUBLE of the device layer -- test infrastructure, never part of the product.

``krypy_b200`` has no CPU path: every N-sized operation is a CUDA kernel behind the C ABI.  To
exercise the HOST logic (solver control flow, restart/deflation bookkeeping, attribute semantics,
exception policy) in the CPU-only test tier, the tests in ``tests/test_host_logic_cpu.py`` swap
``krypy_b200._device.Context`` for the ``FakeContext`` below, whose methods restate the CONTRACT of
each C-ABI entry point (include/krypy_b200.h) in numpy on torch CPU tensors -- including the
mailbox layouts, the "+=" accumulation into h and the zeroing done by the device recurrences.

Nothing under ``krypy_b200/`` imports this file, and no GPU-marked test uses it: the parity tests
proper run the real kernels.  What a green run here proves is that the Python host layer drives the
C ABI consistently; it says nothing about the kernels.
"""
import ctypes

import numpy as np
import scipy.sparse as sp
import torch

from krypy_b200 import _device
from krypy_b200._device import realviews
from krypy_b200._lib import KRY_ORTH_CGS


def _view(x, n=None):
    """numpy float64 view of a CPU tensor or of a raw address (ints come from pointer arithmetic)"""
    if x is None:
        return None
    if isinstance(x, int):
        assert n is not None
        return np.ctypeslib.as_array(ctypes.cast(x, ctypes.POINTER(ctypes.c_double)), shape=(n,))
    return x.detach().numpy()


def _drotg(a, b):
    """csrc/kry_small.cu: kry_drotg"""
    an, bn = abs(a), abs(b)
    if bn == 0.0:
        return 1.0, 0.0
    if an == 0.0:
        return 0.0, 1.0
    scl = min(4.4942328371557898e+307, max(2.2250738585072014e-308, an, bn))
    sigma = np.copysign(1.0, a) if an > bn else np.copysign(1.0, b)
    r = sigma * (scl * np.sqrt((a / scl) ** 2 + (b / scl) ** 2))
    return a / r, b / r


def _rot(c, s, x0, x1):
    return c * x0 + s * x1, -s * x0 + c * x1


def rviewc(x):
    """complex128 view of a vector given as complex tensor or as its interleaved float64 view"""
    return x if x.dtype == torch.complex128 else x.view(torch.complex128)


class _Event(object):
    def record(self):
        pass

    def synchronize(self):
        pass


class FakeContext(object):
    def __init__(self):
        self.device = torch.device("cpu")
        self.mailbox = np.zeros(16384)
        self.timer = None
        self.comm = None
        self.sm_count = 148
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def _timed(self, tag, meta, fn):
        """the KernelTimer brackets of the real Context (bench.py, tools/bench_cplx.py)"""
        if self.timer is None:
            return False
        tm, self.timer = self.timer, None
        try:
            tm.bracket(tag, meta, fn)
        finally:
            self.timer = tm
        return True

    # ---- plumbing ----
    def l2_window(self, t):
        """kry_l2_window: a performance hint, nothing to emulate; the calls are recorded"""
        self.l2_calls = getattr(self, "l2_calls", []) + [None if t is None else (t.data_ptr(), t.numel() * t.element_size())]
        if t is None:
            self._l2win = None
            return None
        nb = t.numel() * t.element_size()
        res = (82903040, 134213632, min(nb, 82903040), nb, min(1.0, 82903040.0 / nb))     # the B200's limits
        self._l2win = ((t.data_ptr(), nb, 0), res)
        return res

    def use_current_stream(self):
        pass

    def sync(self):
        pass

    def event(self):
        return _Event()

    def launch_count(self):
        return sum(self.calls.values())

    def reset_launch_count(self):
        self.calls = {}

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def scalars(self, n):
        return torch.zeros(n, dtype=torch.float64)

    def alloc_basis(self, rows, N, dtype, op=None):
        ld = (int(N) + 31) // 32 * 32
        return torch.zeros((int(rows), ld), dtype=dtype)

    def to_block(self, X, dtype):
        if isinstance(X, torch.Tensor):
            if X.is_complex() and dtype != torch.complex128:
                raise NotImplementedError("complex vectors need a complex linear system / solver dtype")
            Xt = X.detach().to(dtype=dtype)
            Xt = Xt.reshape(1, -1) if Xt.dim() == 1 else Xt.t()
            return Xt.contiguous().clone()
        X = np.asarray(X)
        if np.iscomplexobj(X) and dtype != torch.complex128:
            raise NotImplementedError("complex vectors need a complex linear system / solver dtype")
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        return torch.from_numpy(np.ascontiguousarray(X.T, dtype=_device.torch_to_np_dtype(dtype))).clone()

    def to_numpy(self, Xd):
        return np.ascontiguousarray(Xd.detach().numpy().T)

    def upload_csr(self, A, dtype):
        A = sp.csr_matrix(A)
        npdt = _device.torch_to_np_dtype(dtype)
        return _device.CsrDev(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                              torch.from_numpy(np.ascontiguousarray(A.data, dtype=npdt)), A.shape)

    def upload_csr_z(self, A):
        A = sp.csr_matrix(A)
        npdt = np.complex128 if np.iscomplexobj(A.data) else np.float64
        obj = _device.CsrDev(torch.from_numpy(A.indptr.astype(np.int32)), torch.from_numpy(A.indices.astype(np.int32)),
                             torch.from_numpy(np.ascontiguousarray(A.data, dtype=npdt)), A.shape)
        obj.native_z = True
        return obj

    # ---- kry_spmv_csr_z ----
    def spmv_z(self, A, x, y):
        if self._timed("spmv", (A.shape[0], A.nnz), lambda: self.spmv_z(A, x, y)):
            return
        self._count("spmv_z")
        assert x.dtype == torch.complex128 and y.dtype == torch.complex128 and getattr(A, "native_z", False)
        M = sp.csr_matrix((A.vals.numpy(), A.colidx.numpy(), A.rowptr.numpy()), shape=A.shape)
        y.copy_(torch.from_numpy(np.asarray(M @ x.numpy(), dtype=np.complex128)))

    # ---- operators (kry_spmv_csr, kry_gemv_dense, kry_diag_mul) ----
    @realviews
    def spmv(self, A, x, y, w=None, dot_out=None):
        if self._timed("spmv", (A.shape[0], A.nnz), lambda: self.spmv(A, x, y, w, dot_out)):
            return
        self._count("spmv")
        M = sp.csr_matrix((A.vals.numpy().astype(np.float64), A.colidx.numpy(), A.rowptr.numpy()), shape=A.shape)
        r = M @ x.numpy().astype(np.float64)
        if y is not None:
            y.copy_(torch.from_numpy(r).to(y.dtype))
        if w is not None:
            dot_out[0] = float(w.numpy().astype(np.float64) @ r)

    @realviews
    def gemv(self, A, x, y):
        self._count("gemv")
        y.copy_(torch.from_numpy(A.numpy().astype(np.float64) @ x.numpy().astype(np.float64)).to(y.dtype))

    @realviews
    def diag_mul(self, d, x, y):
        self._count("diag_mul")
        y.copy_((d.double() * x.double()).to(y.dtype))

    # ---- elementwise (kry_axpby, kry_axpy_dev, kry_scale_dev) ----
    @realviews
    def axpby(self, a, x, b, y, z):
        self._count("axpby")
        r = float(a) * x.double()
        if y is not None:
            r = r + float(b) * y.double()
        z.copy_(r.to(z.dtype))

    @realviews
    def axpy_dev(self, coef, sign, x, y):
        self._count("axpy_dev")
        y.copy_((y.double() + float(sign) * float(coef[0]) * x.double()).to(y.dtype))

    @realviews
    def scale_dev(self, s, divide, mul, x, out):
        self._count("scale_dev")
        v = float(mul) * x.double()
        out.copy_((v / float(s[0]) if divide else v * float(s[0])).to(out.dtype))

    @realviews
    def rot90(self, x, y):
        self._count("rot90")
        xv = x.detach().clone()
        y[0::2] = -xv[1::2]
        y[1::2] = xv[0::2]

    # ---- tall-skinny (kry_block_dot, kry_block_axpy, kry_block_combine) ----
    @realviews
    def block_dot(self, V, nv, q, out, post=0, acc=None):
        self._count("block_dot")
        o = _view(out, nv)
        a = _view(acc, nv)
        qq = q.numpy().astype(np.float64)
        for j in range(int(nv)):
            s = float(V[j].numpy().astype(np.float64) @ qq)
            if post == 1:
                s = float(np.sqrt(abs(s)))
            o[j] = s
            if a is not None:
                a[j] += s

    @realviews
    def block_axpy(self, V, nv, coef, sign, q):
        self._count("block_axpy")
        r = q.double()
        for j in range(int(nv)):
            r = r + float(sign) * float(coef[j]) * V[j].double()
        q.copy_(r.to(q.dtype))

    @realviews
    def block_combine(self, V, nv, coef, x0, out):
        self._count("block_combine")
        s = torch.zeros(out.shape, dtype=torch.float64)
        for j in range(int(nv)):
            s = s + float(coef[j]) * V[j].double()
        if x0 is not None:
            s = x0.double() + s
        out.copy_(s.to(out.dtype))

    # ---- kry_orth_fused ----
    @realviews
    def orth_fused(self, Vdot, Vsub, j0, nv, q, passes, algo, h, nrm=None, vnext=None, pre_vec=None,
                   pre_coef=None, h_ptr=None, halo_op=None):
        if self._timed("orth", (q.numel(), int(nv) - int(j0), int(passes), int(algo), vnext is not None),
                       lambda: self.orth_fused(Vdot, Vsub, j0, nv, q, passes, algo, h, nrm, vnext, pre_vec, pre_coef,
                                               h_ptr, halo_op)):
            return
        self._count("orth_fused")
        hv = _view(h_ptr if h_ptr is not None else h, max(int(nv), 1)) if (h is not None or h_ptr is not None) else None
        qq = q.double()
        if pre_vec is not None:
            qq = (qq - float(pre_coef[0]) * pre_vec.double()).to(q.dtype).double()
        for _ in range(int(passes)):
            if algo == KRY_ORTH_CGS:
                cs = [float(Vdot[j].double() @ qq) for j in range(int(j0), int(nv))]
                for j, c in zip(range(int(j0), int(nv)), cs):
                    hv[j] += c
                    qq = qq - c * Vsub[j].double()
                qq = qq.to(q.dtype).double()
            else:
                for j in range(int(j0), int(nv)):
                    c = float(Vdot[j].double() @ qq)
                    hv[j] += c
                    qq = (qq - c * Vsub[j].double()).to(q.dtype).double()
        q.copy_(qq.to(q.dtype))
        if nrm is not None:
            n2 = float(np.sqrt(float(qq @ qq)))
            nrm[0] = n2
            if vnext is not None:
                vnext.copy_((qq / n2 if n2 > 0 else torch.zeros_like(qq)).to(vnext.dtype))

    # ---- kry_orth_fused_z ----
    def orth_fused_z(self, Vdot, Vsub, ldv, j0, nv, q, passes, algo, h_ptr, nrm=None, vnext=None):
        """complex vectors j0..nv-1 start ldv complex elements apart at Vdot / Vsub (tensors: row 0 of a real
        twin storage, the vectors are its even rows); h: 2 doubles per vector, +="""
        qc = rviewc(q)
        n = qc.numel()
        if self._timed("orth", (2 * n, 2 * (int(nv) - int(j0)), int(passes), int(algo), vnext is not None),
                       lambda: self.orth_fused_z(Vdot, Vsub, ldv, j0, nv, q, passes, algo, h_ptr, nrm, vnext)):
            return
        self._count("orth_fused_z")

        def vec(B, j):
            # B: real (rows, >= 2n) view whose row 2j holds complex vector j (ldv = its real row stride)
            return B[2 * j][: 2 * n].view(torch.complex128)
        assert int(ldv) == Vdot.stride(0) and int(ldv) == Vsub.stride(0)
        hv = _view(h_ptr, 2 * max(int(nv), 1)).view(np.complex128) if isinstance(h_ptr, int) else \
            h_ptr.numpy()[: 2 * max(int(nv), 1)].view(np.complex128)
        qq = qc.clone()
        for _ in range(int(passes)):
            if algo == KRY_ORTH_CGS:
                cs = [complex(torch.vdot(vec(Vdot, j), qq)) for j in range(int(j0), int(nv))]
                for j, c in zip(range(int(j0), int(nv)), cs):
                    hv[j] += c
                    qq = qq - c * vec(Vsub, j)
            else:
                for j in range(int(j0), int(nv)):
                    c = complex(torch.vdot(vec(Vdot, j), qq))
                    hv[j] += c
                    qq = qq - c * vec(Vsub, j)
        qc.copy_(qq)
        if nrm is not None:
            n2 = float(np.sqrt(float(torch.vdot(qq, qq).real)))
            nrm[0] = n2
            if vnext is not None:
                rviewc(vnext).copy_(qq / n2 if n2 > 0 else torch.zeros_like(qq))

    # ---- kry_lanczos_diag ----
    def lanczos_diag(self, vprev, vk, bdiag, q, pre_coef, h3, vnext):
        self._count("lanczos_diag")
        dt = q.dtype
        qq = q.double()
        if vprev is not None:
            qq = (qq - float(pre_coef[0]) * vprev.double()).to(dt).double()
        bq = (bdiag.double() * qq).to(dt).double()
        alpha = float(vk.double() @ bq)
        h3[1] += alpha
        qq = (qq - alpha * vk.double()).to(dt).double()
        bq = (bdiag.double() * qq).to(dt).double()
        beta = float(np.sqrt(abs(float(qq @ bq))))
        h3[2] = beta
        q.copy_(qq.to(dt))
        if vnext is not None:
            vnext.copy_((qq / beta if beta > 0 else torch.zeros_like(qq)).to(vnext.dtype))

    # ---- kry_project ----
    @realviews
    def project(self, W, V, d, a, Q, R, iterations, c_first):
        import scipy.linalg
        self._count("project")
        aa = a.double()
        Wn = W[: int(d)].double().numpy()
        Vn = V[: int(d)].double().numpy()
        for it in range(int(iterations)):
            c = Wn @ aa.numpy()
            if it == 0 and c_first is not None:
                c_first[: int(d)] = torch.from_numpy(c)
            if Q is not None:
                c = scipy.linalg.solve_triangular(R.numpy(), Q.numpy().T @ c)
            aa = aa - torch.from_numpy(Vn.T @ c)
        a.copy_(aa.to(a.dtype))

    # ---- small recurrences ----
    def givens_update(self, k, hcol, rcol, cs, y, off=0):
        self._count("givens")
        h = hcol.numpy()
        r = h[: k + 2].copy()
        mb = self.mailbox
        mb[off + 1: off + k + 3] = r
        h[: k + 2] = 0.0
        c_ = cs.numpy()
        for i in range(k):
            r[i], r[i + 1] = _rot(c_[2 * i], c_[2 * i + 1], r[i], r[i + 1])
        c, s = _drotg(r[k], r[k + 1])
        c_[2 * k], c_[2 * k + 1] = c, s
        r[k], r[k + 1] = _rot(c, s, r[k], r[k + 1])
        yy = y.numpy()
        yy[k], yy[k + 1] = _rot(c, s, yy[k], yy[k + 1])
        mb[off] = abs(yy[k + 1])
        rcol.numpy()[: k + 2] = r
        mb[off + k + 3: off + 2 * k + 5] = r

    def givens_update_z(self, k, hcol, rcol, cs, y, off=0):
        """csrc/kry_small.cu: givens_z_kernel (interleaved complex; drotg for real-valued pairs,
        zrotg otherwise; the serial core itself is tested in tests/test_small_core_cpu.py)"""
        self._count("givens_z")
        from krypy_b200.utils import _zrotg
        h = hcol.numpy()
        nr = 2 * (k + 2)
        mb = self.mailbox
        mb[off + 1: off + 1 + nr] = h[:nr]
        r = h[:nr].copy().view(np.complex128)
        h[:nr] = 0.0
        c_ = cs.numpy()

        def rot(c, s, x0, x1):
            return c * x0 + s * x1, -np.conj(s) * x0 + c * x1
        for i in range(k):
            r[i], r[i + 1] = rot(c_[4 * i], complex(c_[4 * i + 2], c_[4 * i + 3]), r[i], r[i + 1])
        if r[k].imag == 0.0 and r[k + 1].imag == 0.0:
            c, s = _drotg(r[k].real, r[k + 1].real)
            s = complex(s)
            flag = 1.0
        else:
            c, s = _zrotg(r[k], r[k + 1])
            flag = 0.0
        c_[4 * k: 4 * k + 4] = [c, flag, s.real, s.imag]
        r[k], r[k + 1] = rot(c, s, r[k], r[k + 1])
        yy = y.numpy()[:nr].view(np.complex128)
        yy[k], yy[k + 1] = rot(c, s, yy[k], yy[k + 1])
        mb[off] = abs(yy[k + 1])
        rcol.numpy()[:nr] = r.view(np.float64)
        mb[off + 1 + nr: off + 1 + 2 * nr] = r.view(np.float64)

    def tri_solve_z(self, k, R, y, out):
        import scipy.linalg
        self._count("tri_solve_z")
        Rc = np.ascontiguousarray(R.numpy()[:k, :2 * k]).view(np.complex128)
        yc = y.numpy()[:2 * k].copy().view(np.complex128)
        out.numpy()[:2 * k] = scipy.linalg.solve_triangular(Rc, yc).view(np.float64)

    def tri_solve_t(self, k, Rt, y, out):
        import scipy.linalg
        self._count("tri_solve_t")
        out[: int(k)] = torch.from_numpy(scipy.linalg.solve_triangular(Rt.numpy()[:k, :k].T, y.numpy()[:k]))

    def tri_solve(self, k, R, y, out):
        import scipy.linalg
        self._count("tri_solve")
        out[: int(k)] = torch.from_numpy(scipy.linalg.solve_triangular(R.numpy()[:k, :k], y.numpy()[:k]))

    def minres_recur(self, k, h3, st, shift=1, off=0):
        self._count("minres_recur")
        h = h3.numpy()
        s_ = st.numpy()
        R0, R1 = 0.0, (h[0] if k > 0 else 0.0)
        if s_[2] != 0.0:
            R0, R1 = _rot(s_[0], s_[1], R0, R1)
        R2, R3 = h[1], h[2]
        if s_[5] != 0.0:
            R1, R2 = _rot(s_[3], s_[4], R1, R2)
        s_[0:3] = s_[3:6]
        c, s = _drotg(R2, R3)
        s_[3], s_[4], s_[5] = c, s, 1.0
        R2 = c * R2 + s * R3
        y0, y1 = _rot(c, s, s_[6], 0.0)
        s_[8:12] = [R0, R1, R2, y0]
        s_[6] = y1
        mb = self.mailbox
        mb[off: off + 8] = [abs(y1), R0, R1, R2, y0, h[0], h[1], h[2]]
        if shift:
            h[0] = h[2]
            h[1] = 0.0

    @realviews
    def minres_update(self, v, w0, w1, yk, st):
        self._count("minres_update")
        s_ = st.numpy()
        z = ((v.double() - s_[8] * w0.double() - s_[9] * w1.double()) / s_[10]).to(w0.dtype)
        w0.copy_(z)
        yk.copy_((yk.double() + s_[11] * z.double()).to(yk.dtype))

    @realviews
    def cg_update(self, Ap, p, yk, r, z, dinv, rho, pAp, off=0):
        self._count("cg_update")
        pap = float(pAp[0])
        alpha = float(rho) / pap
        yk.copy_((yk.double() + alpha * p.double()).to(yk.dtype))
        r.copy_((r.double() - alpha * Ap.double()).to(r.dtype))
        if dinv is not None:
            z.copy_((dinv.double() * r.double()).to(z.dtype))
            rz = float(r.double() @ z.double())
        else:
            rz = float(r.double() @ r.double())
        self.mailbox[off: off + 3] = [rz, alpha, pap]

    # ---- CG with device-resident scalars: kry_cg_update_dev / kry_cg_scalars / kry_xpby_dev ----
    @realviews
    def cg_update_dev(self, Ap, p, yk, r, z, dinv, st):
        self._count("cg_update_dev")
        alpha = float(st[1]) / float(st[2])
        yk.copy_((yk.double() + alpha * p.double()).to(yk.dtype))
        r.copy_((r.double() - alpha * Ap.double()).to(r.dtype))
        if dinv is not None:
            z.copy_((dinv.double() * r.double()).to(z.dtype))
            rz = float(r.double() @ z.double())
        else:
            rz = float(r.double() @ r.double())
        st[3] = alpha
        st[5] = rz

    def cg_scalars(self, st, off=0):
        self._count("cg_scalars")
        s = float(st[5])
        nrm = float(np.sqrt(abs(s)))
        prev = float(st[1])
        st[0] = prev
        st[1] = nrm * nrm
        st[4] = (nrm * nrm) / prev
        self.mailbox[off: off + 3] = [s, float(st[3]), float(st[2])]

    @realviews
    def xpby_dev(self, x, beta, y, out):
        self._count("xpby_dev")
        out.copy_((x.double() + float(beta[0]) * y.double()).to(out.dtype))

    # ---- kry_gram / kry_block_trsm ----
    @staticmethod
    def gram_fits(kx, ky, same):
        return ((kx if same else kx + ky) <= 64) and ((kx + 3) // 4) * ((ky + 3) // 4) <= 32

    @realviews
    def gram(self, X, kx, Y, ky, out):
        self._count("gram")
        G = X[:kx].double() @ Y[:ky].double().T
        out[: kx * ky].copy_(G.reshape(-1))

    @realviews
    def block_trsm(self, X, d, R, Q):
        self._count("block_trsm")
        import scipy.linalg
        Rm = R[: d * d].reshape(d, d).numpy()
        Qn = scipy.linalg.solve_triangular(Rm, X[:d].double().numpy(), trans="T", lower=False)
        Q[:d].copy_(torch.from_numpy(np.ascontiguousarray(Qn)).to(Q.dtype))


def install(monkeypatch):
    """swap the product's device context for the test double (pytest monkeypatch fixture)"""
    fake = FakeContext()
    monkeypatch.setattr(_device.Context, "get", classmethod(lambda cls, device=None: fake))
    return fake
