"""Device plumbing: context handle, torch tensors as device-array containers,
thin typed wrappers over the C ABI (include/krypy_b200.h).

PyTorch is used only to own device memory and expose the current CUDA stream;
every N-sized arithmetic operation goes through libkrypy_b200.so.

Internal layout: a block of k vectors of length N is a contiguous torch tensor
of shape ``(k, N)`` ("vector-major"); the public API exposes ``(N, k)`` numpy
arrays like the reference does.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import KRY_F32, KRY_F64, check

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def np_to_torch_dtype(dt):
    t = torch()
    dt = np.dtype(dt)
    if dt == np.float64:
        return t.float64
    if dt == np.float32:
        return t.float32
    if dt.kind == "c":
        return t.complex128
    raise NotImplementedError(
        "krypy_b200: dtype %s is not supported by the device path (float32/float64/complex128; "
        "there is no CPU fallback)" % dt)


def torch_to_np_dtype(dt):
    t = torch()
    if dt == t.float64:
        return np.dtype(np.float64)
    if dt == t.float32:
        return np.dtype(np.float32)
    if dt == t.complex128:
        return np.dtype(np.complex128)
    raise NotImplementedError("unsupported torch dtype %s" % dt)


def code(t):
    if t.dtype == torch().float64:
        return KRY_F64
    if t.dtype == torch().float32:
        return KRY_F32
    raise TypeError("kernel argument of dtype %s (complex blocks are passed as real views)" % t.dtype)


def is_complex(t):
    return t is not None and not isinstance(t, int) and t.dtype == torch().complex128


def rview(t):
    """Real view of a complex128 tensor: (..., N) complex -> (..., 2N) float64, re/im interleaved
    (same memory).  Real tensors, ints (raw addresses) and None pass through."""
    if t is None or isinstance(t, int) or t.dtype != torch().complex128:
        return t
    return t.view(torch().float64)


def realviews(fn):
    """Kernel wrappers see complex blocks as their interleaved real views: the kernels are real
    (krypy_b200/_cplx.py: complex arithmetic by real embedding + twin storage)."""
    import functools
    T = torch

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        tens = T().Tensor
        args = [rview(a) if isinstance(a, tens) else a for a in args]
        for k, v in kwargs.items():
            if isinstance(v, tens):
                kwargs[k] = rview(v)
        return fn(self, *args, **kwargs)
    return wrapper


def _p(t):
    if t is None or isinstance(t, int):
        return t
    return t.data_ptr()


class CsrDev(object):
    """CSR matrix resident in HBM (int32 row pointers / column indices)."""

    def __init__(self, rowptr, colidx, vals, shape):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(vals.shape[0])

    @property
    def dtype(self):
        return self.vals.dtype

    def nbytes(self):
        return (self.rowptr.numel() + self.colidx.numel()) * 4 + self.vals.numel() * self.vals.element_size()


class KernelTimer(object):
    """CUDA-event brackets around selected kernel launches (bench.py roofline):
    events are recorded on the launching stream; ``summary()`` synchronises."""

    def __init__(self):
        self.events = {}

    def bracket(self, tag, meta, fn):
        t = torch()
        e0 = t.cuda.Event(enable_timing=True)
        e1 = t.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.events.setdefault(tag, []).append((e0, e1, meta))

    def summary(self):
        torch().cuda.synchronize()
        out = {}
        for tag, lst in self.events.items():
            ms = [e0.elapsed_time(e1) for (e0, e1, _) in lst]
            out[tag] = dict(launches=len(lst), ms_total=float(sum(ms)), meta=[m for (_, _, m) in lst], ms=ms)
        return out


class Context(object):
    """One per (process, device): owns the kry_ctx handle and the pinned mailbox."""

    _instances = {}

    @classmethod
    def get(cls, device=None):
        t = torch()
        if not t.cuda.is_available():
            raise RuntimeError(
                "krypy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        idx = t.cuda.current_device() if device is None else t.device(device).index
        if idx is None:
            idx = t.cuda.current_device()
        inst = cls._instances.get(idx)
        if inst is None:
            inst = cls._instances[idx] = cls(idx)
        return inst

    def __init__(self, idx):
        t = torch()
        self.lib = lib = _lib.load()
        self.index = idx
        self.device = t.device("cuda", idx)
        self.stream_handle = t.cuda.current_stream(self.device).cuda_stream
        h = ctypes.c_void_p()
        check(lib.kry_ctx_create(idx, ctypes.c_void_p(self.stream_handle), ctypes.byref(h)))
        self.h = h
        host = lib.kry_mailbox_host(h)
        self.mailbox = np.ctypeslib.as_array(
            ctypes.cast(host, ctypes.POINTER(ctypes.c_double)), shape=(_lib.KRY_MAILBOX_DOUBLES,))
        self.mailbox_dev = int(lib.kry_mailbox_dev(h))      # device alias of the mapped mailbox
        info = (ctypes.c_longlong * 8)()
        check(lib.kry_device_info(h, info))
        self.sm_count, self.cc, self.l2_bytes = int(info[0]), int(info[1]), int(info[2])
        self.orth_blocks = int(info[5])
        self.cycle_ahead = True    # linsys.Gmres may enqueue whole restart cycles ahead of the host (real device only)
        self.timer = None          # set to a KernelTimer by bench.py
        self.comm = None           # set to a dist.PeerComm: reductions become global sums
        self._tmpc = None

    # ---- stream / sync ------------------------------------------------
    def use_current_stream(self):
        s = torch().cuda.current_stream(self.device).cuda_stream
        if s != self.stream_handle:
            self.stream_handle = s
            check(self.lib.kry_ctx_set_stream(self.h, ctypes.c_void_p(s)))

    def sync(self):
        check(self.lib.kry_sync(self.h))

    def event(self):
        """a CUDA event on the context's stream (record() / synchronize())"""
        return torch().cuda.Event()

    def launch_count(self):
        return int(self.lib.kry_launch_count(self.h))

    def reset_launch_count(self):
        self.lib.kry_reset_launch_count(self.h)

    def l2_window(self, t):
        """Keep the device tensor ``t`` resident in the L2 set-aside for the kernels launched from now on
        (kry_l2_window); ``None`` removes the window.  A performance hint only; returns the info tuple
        (max set-aside, max window, set-aside, window bytes, hit ratio) or None when nothing was set."""
        cur = getattr(self, "_l2win", None)
        if t is None:
            if cur is not None:
                self._l2win = None
                self.lib.kry_l2_window(self.h, None, 0, None)        # (a hint: its failure is not an error)
            return None
        key = (t.data_ptr(), t.numel() * t.element_size(), self.stream_handle)
        if cur is not None and cur[0] == key:
            return cur[1]
        info = (ctypes.c_longlong * 5)()
        rc = self.lib.kry_l2_window(self.h, ctypes.c_void_p(key[0]), key[1], info)
        if rc != 0:
            msg = self.lib.kry_last_error()
            self.l2_window_error = msg.decode() if msg else "?"
            self._l2win = (key, None)                                # do not retry for the same buffer
            return None
        res = (int(info[0]), int(info[1]), int(info[2]), int(info[3]), info[4] * 1e-6)
        self._l2win = (key, res)
        return res

    # ---- allocation / conversion ---------------------------------------
    def empty(self, shape, dtype):
        return torch().empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch().zeros(shape, dtype=dtype, device=self.device)

    def scalars(self, n):
        """n zeroed device doubles (small coefficients always live in fp64)."""
        return torch().zeros(n, dtype=torch().float64, device=self.device)

    def to_block(self, X, dtype):
        """numpy/torch array (N,) or (N,k) -> contiguous device tensor (k, N)."""
        t = torch()
        if isinstance(X, t.Tensor):
            if X.is_complex() and dtype != t.complex128:
                raise NotImplementedError("complex vectors need a complex linear system / solver dtype")
            Xt = X.detach().to(device=self.device, dtype=dtype)
            Xt = Xt.reshape(1, -1) if Xt.dim() == 1 else Xt.t()
            return Xt.contiguous().clone()   # inputs are never mutated (ownership)
        X = np.asarray(X)
        if np.iscomplexobj(X) and dtype != t.complex128:
            raise NotImplementedError("complex vectors need a complex linear system / solver dtype")
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        if X.shape[1] > 1 and X.shape[0] * X.shape[1] > (1 << 20):
            # wide blocks (deflation spaces): upload as is, transpose on the device
            Xd = t.from_numpy(np.ascontiguousarray(X)).to(self.device)
            return Xd.to(dtype).t().contiguous()
        Xh = np.ascontiguousarray(X.T, dtype=torch_to_np_dtype(dtype))
        return t.from_numpy(Xh).to(self.device)

    def to_numpy(self, Xd):
        """device (k, N) -> numpy (N, k)."""
        t = torch()
        Xd = Xd.detach()
        k = Xd.shape[0]
        if Xd.numel() * Xd.element_size() < (1 << 20):
            return np.ascontiguousarray(Xd.cpu().numpy().T)
        # Large results: DMA into page-locked memory (torch's caching host allocator keeps the
        # block for the next solve) instead of a pageable .cpu() copy, which runs at a tenth of
        # the PCIe rate.  The returned array owns the pinned block (fresh per call: results are
        # never aliased, SURVEY 8b ownership).  (k, N) -> (N, k): transposed on the device.
        src = Xd.reshape(-1, 1) if k == 1 else Xd.t()
        if not src.is_contiguous():
            src = src.contiguous()
        host = t.empty(src.shape, dtype=src.dtype, pin_memory=True)
        host.copy_(src, non_blocking=True)
        t.cuda.current_stream(self.device).synchronize()
        return host.numpy()

    def upload_csr(self, A, dtype):
        """scipy.sparse matrix (any format) -> CsrDev."""
        import scipy.sparse as sp
        t = torch()
        A = sp.csr_matrix(A) if not sp.isspmatrix_csr(A) else A
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        if A.nnz >= 2 ** 31 - 1:
            raise NotImplementedError("nnz >= 2^31 needs 64-bit row pointers (not built)")
        npdt = torch_to_np_dtype(dtype)
        rowptr = t.from_numpy(np.ascontiguousarray(A.indptr, dtype=np.int32)).to(self.device)
        colidx = t.from_numpy(np.ascontiguousarray(A.indices, dtype=np.int32)).to(self.device)
        vals = t.from_numpy(np.ascontiguousarray(A.data, dtype=npdt)).to(self.device)
        return CsrDev(rowptr, colidx, vals, A.shape)

    def upload_csr_z(self, A):
        """scipy.sparse matrix -> CsrDev for the NATIVE complex SpMV (kry_spmv_csr_z): complex128 values for
        a complex matrix (20 bytes per entry), float64 for a real matrix applied to complex vectors (12);
        the real embedding (_cplx.expand_sparse) holds 48 bytes per entry."""
        import scipy.sparse as sp
        t = torch()
        A = sp.csr_matrix(A) if not sp.isspmatrix_csr(A) else A
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        if A.nnz >= 2 ** 31 - 1:
            raise NotImplementedError("nnz >= 2^31 needs 64-bit row pointers (not built)")
        npdt = np.complex128 if np.iscomplexobj(A.data) else np.float64
        rowptr = t.from_numpy(np.ascontiguousarray(A.indptr, dtype=np.int32)).to(self.device)
        colidx = t.from_numpy(np.ascontiguousarray(A.indices, dtype=np.int32)).to(self.device)
        vals = t.from_numpy(np.ascontiguousarray(A.data, dtype=npdt)).to(self.device)
        obj = CsrDev(rowptr, colidx, vals, A.shape)
        obj.native_z = True
        return obj

    def spmv_z(self, A, x, y):
        """y = A x on complex128 vectors, natively (A from upload_csr_z)"""
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("spmv", (A.shape[0], A.nnz), lambda: self.spmv_z(A, x, y))
            self.timer = tm
            return
        t = torch()
        if x.dtype != t.complex128 or y.dtype != t.complex128:
            raise TypeError("spmv_z works on complex128 vectors")
        check(self.lib.kry_spmv_csr_z(self.h, 1 if A.vals.dtype == t.complex128 else 0, A.shape[0], A.shape[1],
                                      A.nnz, A.rowptr.data_ptr(), A.colidx.data_ptr(), A.vals.data_ptr(),
                                      x.data_ptr(), y.data_ptr()))

    # ---- operators -------------------------------------------------------
    @realviews
    def spmv(self, A, x, y, w=None, dot_out=None):
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("spmv", (A.shape[0], A.nnz), lambda: self.spmv(A, x, y, w, dot_out))
            self.timer = tm
            return
        check(self.lib.kry_spmv_csr(self.h, code(A.vals), A.shape[0], A.shape[1], A.nnz,
                                    A.rowptr.data_ptr(), A.colidx.data_ptr(), A.vals.data_ptr(),
                                    x.data_ptr(), _p(y), _p(w), _p(dot_out)))
        if self.comm is not None and dot_out is not None:
            self.comm.allreduce(dot_out, 1)

    @realviews
    def gemv(self, A, x, y):
        check(self.lib.kry_gemv_dense(self.h, code(A), A.shape[0], A.shape[1], A.data_ptr(),
                                      A.stride(0), x.data_ptr(), y.data_ptr()))

    @realviews
    def diag_mul(self, d, x, y):
        check(self.lib.kry_diag_mul(self.h, code(x), x.numel(), d.data_ptr(), x.data_ptr(), y.data_ptr()))

    # ---- elementwise -------------------------------------------------------
    @realviews
    def axpby(self, a, x, b, y, z):
        check(self.lib.kry_axpby(self.h, code(x), x.numel(), float(a), x.data_ptr(), float(b), _p(y),
                                 z.data_ptr()))

    @realviews
    def axpy_dev(self, coef, sign, x, y):
        check(self.lib.kry_axpy_dev(self.h, code(x), x.numel(), coef.data_ptr(), float(sign),
                                    x.data_ptr(), y.data_ptr()))

    @realviews
    def scale_dev(self, s, divide, mul, x, out):
        check(self.lib.kry_scale_dev(self.h, code(x), x.numel(), s.data_ptr(), int(divide), float(mul),
                                     x.data_ptr(), out.data_ptr()))

    @realviews
    def rot90(self, x, y):
        """y = i*x for interleaved complex data held in real tensors (len(x) = 2 * #complex)"""
        check(self.lib.kry_rot90(self.h, code(x), x.numel() // 2, x.data_ptr(), y.data_ptr()))

    # ---- tall-skinny -----------------------------------------------------
    @realviews
    def block_dot(self, V, nv, q, out, post=0, acc=None):
        """out[j] = <V[j], q>, j < nv.  V: (>=nv, N) tensor (row stride = ld)."""
        if self.comm is not None:
            # row-partitioned run: local partial sums, then the global sum over NVLink peer memory
            check(self.lib.kry_block_dot(self.h, code(q), q.numel(), _p(V), V.stride(0) if V is not None else 0,
                                         int(nv), q.data_ptr(), _p(out), 0, None))
            self.comm.allreduce(out, int(nv), post=int(post), acc=acc)
            return
        check(self.lib.kry_block_dot(self.h, code(q), q.numel(), _p(V), V.stride(0) if V is not None else 0,
                                     int(nv), q.data_ptr(), _p(out), int(post), _p(acc)))

    @realviews
    def block_axpy(self, V, nv, coef, sign, q):
        check(self.lib.kry_block_axpy(self.h, code(q), q.numel(), V.data_ptr(), V.stride(0), int(nv),
                                      coef.data_ptr(), float(sign), q.data_ptr()))

    @realviews
    def block_combine(self, V, nv, coef, x0, out):
        check(self.lib.kry_block_combine(self.h, code(out), out.numel(), _p(V),
                                         V.stride(0) if V is not None else 0, int(nv), _p(coef), _p(x0),
                                         out.data_ptr()))

    @realviews
    def orth_fused(self, Vdot, Vsub, j0, nv, q, passes, algo, h, nrm=None, vnext=None, pre_vec=None,
                   pre_coef=None, h_ptr=None, halo_op=None):
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("orth", (q.numel(), int(nv) - int(j0), int(passes), int(algo), vnext is not None),
                       lambda: self.orth_fused(Vdot, Vsub, j0, nv, q, passes, algo, h, nrm, vnext, pre_vec,
                                               pre_coef, h_ptr, halo_op))
            self.timer = tm
            return
        if self.comm is not None:
            return self._orth_split(Vdot, Vsub, j0, nv, q, passes, algo, h, nrm, vnext, pre_vec, pre_coef, h_ptr,
                                    halo_op)
        ld = Vdot.stride(0) if Vdot is not None else 0
        hp = h_ptr if h_ptr is not None else _p(h)
        check(self.lib.kry_orth_fused(self.h, code(q), q.numel(), _p(Vdot), _p(Vsub), ld, int(j0), int(nv),
                                      q.data_ptr(), int(passes), int(algo), _p(pre_vec), _p(pre_coef),
                                      hp, _p(nrm), _p(vnext)))

    def orth_fused_z(self, Vdot, Vsub, ldv, j0, nv, q, passes, algo, h_ptr, nrm=None, vnext=None):
        """kry_orth_fused_z: the fused Gram-Schmidt step on complex128 vectors, natively.  Vdot / Vsub: tensor
        (or address) of complex vector 0 of the basis, ``ldv`` complex elements between consecutive vectors
        (twin storage: the even rows, ldv = the real row stride); q, vnext: complex vectors (any view of
        their 2N doubles); h_ptr: address of the interleaved coefficient array (2 doubles per vector, +=);
        nrm: device double."""
        n = q.numel() if q.dtype == torch().complex128 else q.numel() // 2
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("orth", (2 * n, 2 * (int(nv) - int(j0)), int(passes), int(algo), vnext is not None),
                       lambda: self.orth_fused_z(Vdot, Vsub, ldv, j0, nv, q, passes, algo, h_ptr, nrm, vnext))
            self.timer = tm
            return
        if self.comm is not None:
            raise NotImplementedError("complex row-partitioned runs are not implemented")
        check(self.lib.kry_orth_fused_z(self.h, n, _p(Vdot), _p(Vsub), int(ldv), int(j0), int(nv), q.data_ptr(),
                                        int(passes), int(algo), _p(h_ptr), _p(nrm), _p(vnext)))

    def _orth_split(self, Vdot, Vsub, j0, nv, q, passes, algo, h, nrm, vnext, pre_vec, pre_coef, h_ptr,
                    halo_op=None):
        """Row-partitioned Gram-Schmidt step: the phases of kry_orth_fused as separate kernels
        with the global sums (NVLink peer all-reduce) between them.  CGS: one reduction of nv
        values per pass; MGS: one per basis vector (exact reference order, latency bound)."""
        from ._lib import KRY_ORTH_CGS
        if self._tmpc is None:
            self._tmpc = self.scalars(64)
        tmp = self._tmpc
        hbase = h_ptr if h_ptr is not None else h.data_ptr()
        if pre_vec is not None:
            self.axpy_dev(pre_coef, -1.0, pre_vec, q)
        comm = self.comm
        if comm.reduce == "peer" and comm.orth_mode == "coop":
            # ONE cooperative kernel per call: the reductions are completed over NVLink inside it
            lib, w, r = self.lib, comm.world, comm.rank
            ep, sl, fl = comm.epoch_dev.data_ptr(), comm.slots.peer_table.data_ptr(), comm.flags.peer_table.data_ptr()
            ld = Vdot.stride(0) if Vdot is not None else 0
            if algo != KRY_ORTH_CGS or int(nv) - int(j0) <= 64:
                check(lib.kry_orth_fused_dist(self.h, code(q), q.numel(), _p(Vdot), _p(Vsub), ld, int(j0), int(nv),
                                              q.data_ptr(), int(passes), int(algo), _p(pre_vec), _p(pre_coef),
                                              hbase, _p(nrm), _p(vnext), w, r, ep, sl, fl))
                return
            j = int(j0)
            first = True
            while j < nv:                                  # block-wise CGS over chunks of 64 basis vectors
                j1 = min(j + 64, int(nv))
                lastc = j1 == int(nv)
                check(lib.kry_orth_fused_dist(self.h, code(q), q.numel(), _p(Vdot), _p(Vsub), ld, j, j1,
                                              q.data_ptr(), int(passes), int(algo),
                                              _p(pre_vec) if first else None, _p(pre_coef) if first else None,
                                              hbase, _p(nrm) if lastc else None, _p(vnext) if lastc else None,
                                              w, r, ep, sl, fl))
                first = False
                j = j1
            return
        fused = (algo == KRY_ORTH_CGS and comm.reduce == "peer" and int(nv) > int(j0))
        if fused:
            # exchange fused into the kernels: dot(+publish) -> update(+acquire, +||q||^2 publish) -> scale
            lib, w, r = self.lib, comm.world, comm.rank
            ep, sl, fl = comm.epoch_dev.data_ptr(), comm.slots.peer_table.data_ptr(), comm.flags.peer_table.data_ptr()
            dt, n, ld = code(q), q.numel(), Vdot.stride(0)
            es = q.element_size()
            for p in range(int(passes)):
                j = int(j0)
                while j < nv:
                    c = min(64, int(nv) - j)
                    lastc = (p == int(passes) - 1) and (j + c == int(nv))
                    check(lib.kry_dist_dot(self.h, dt, n, Vdot.data_ptr() + j * ld * es, ld, c, q.data_ptr(), 0,
                                           w, r, ep, sl, fl))
                    check(lib.kry_dist_update(self.h, dt, n, Vsub.data_ptr() + j * ld * es, ld, c, q.data_ptr(),
                                              hbase + 8 * j, 1 if (lastc and nrm is not None) else 0,
                                              w, r, ep, sl, fl))
                    j += c
            if nrm is not None:
                hal = halo_op._halo_args(vnext) if (halo_op is not None and vnext is not None) else None
                halq = halo_op._halo_src_args(q) if hal is not None else None
                if halq is not None and comm.halo_from_q:
                    # scale + halo of v_next gathered from the peers' un-normalised q: the norm's flag is the
                    # only cross-GPU wait of this kernel (q alternates between two buffers, see utils.Arnoldi)
                    _, _, hp, ho, nhalo, dst = hal
                    q_tab, q_off = halq
                    check(lib.kry_dist_scale_haloq(self.h, dt, n, q.data_ptr(), vnext.data_ptr(), nrm.data_ptr(),
                                                   nhalo, q_tab, q_off, hp, ho, dst, w, r, ep, sl, fl))
                    comm.halo_ready = vnext.data_ptr()
                elif hal is not None:
                    # scale + "segment complete" handshake + halo gather of v_next in one kernel
                    peer_tab, off, hp, ho, nhalo, dst = hal
                    check(lib.kry_dist_scale_halo(self.h, dt, n, q.data_ptr(), vnext.data_ptr(), nrm.data_ptr(),
                                                  nhalo, peer_tab, off, hp, ho, dst, w, r, ep, sl, fl))
                    comm.halo_ready = vnext.data_ptr()
                else:
                    check(lib.kry_dist_scale(self.h, dt, n, q.data_ptr(), _p(vnext), nrm.data_ptr(), w, r, ep, sl, fl))
            return
        for _ in range(int(passes)):
            if algo == KRY_ORTH_CGS:
                j = int(j0)
                while j < nv:
                    c = min(64, int(nv) - j)
                    self.block_dot(Vdot[j:], c, q, tmp, 0, hbase + 8 * j)
                    self.block_axpy(Vsub[j:], c, tmp, -1.0, q)
                    j += c
            else:
                for j in range(int(j0), int(nv)):
                    self.block_dot(Vdot[j:], 1, q, tmp, 0, hbase + 8 * j)
                    self.axpy_dev(tmp, -1.0, Vsub[j], q)
        if nrm is not None:
            self.block_dot(q.reshape(1, -1), 1, q, nrm, 1, None)
            if vnext is not None:
                self.scale_dev(nrm, 1, 1.0, q, vnext)

    def spmv_mdot(self, A, x, y, B, nb, want_sq, out=None):
        """y = A x with c[j] = <B[j], y> (j < nb) and c[nb] = <y, y> in the SpMV epilogue (kry_spmv_csr_mdot).
        Single GPU: the sums go to ``out``; row-partitioned: the local sums are published to the peers
        (consumer: dist_update_scale).  Returns False when the matrix is not on the staged short-row path."""
        from ._lib import KRY_ERR_UNSUPPORTED
        if self.timer is not None:
            tm, self.timer = self.timer, None
            res = []
            tm.bracket("spmv", (A.shape[0], A.nnz, int(nb)),
                       lambda: res.append(self.spmv_mdot(A, x, y, B, nb, want_sq, out)))
            self.timer = tm
            return res[0]
        c = self.comm
        if c is not None:
            w, r = c.world, c.rank
            ep, sl, fl = c.epoch_dev.data_ptr(), c.slots.peer_table.data_ptr(), c.flags.peer_table.data_ptr()
        else:
            w, r, ep, sl, fl = 1, 0, None, None, None
        rc = self.lib.kry_spmv_csr_mdot(self.h, code(A.vals), A.shape[0], A.shape[1], A.nnz, A.rowptr.data_ptr(),
                                        A.colidx.data_ptr(), A.vals.data_ptr(), x.data_ptr(), y.data_ptr(),
                                        _p(B), B.stride(0) if B is not None else 0, int(nb), int(want_sq), _p(out),
                                        w, r, ep, sl, fl)
        if rc == KRY_ERR_UNSUPPORTED:
            return False
        check(rc)
        return True

    def dist_dot_sq(self, V, nv, q):
        """local V^H q and <q, q> published to the peers (kry_dist_dot with want_sq); consumer: dist_update_scale"""
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("orth", (q.numel(), int(nv), 1, 101, False), lambda: self.dist_dot_sq(V, nv, q))
            self.timer = tm
            return
        c = self.comm
        check(self.lib.kry_dist_dot(self.h, code(q), q.numel(), V.data_ptr(), V.stride(0), int(nv), q.data_ptr(), 1,
                                    c.world, c.rank, c.epoch_dev.data_ptr(), c.slots.peer_table.data_ptr(),
                                    c.flags.peer_table.data_ptr()))

    def dist_update_scale(self, V, nv, q, vnext, h_ptr, nrm, halo, halo_q, halo_base, givens=None):
        """the rest of a row-partitioned block-CGS Arnoldi step after dist_dot_sq (kry_dist_update_scale):
        ``halo`` = (peer table, offset, halo_peer, halo_off, nhalo, dst) of vnext (DistCsrOperator._halo_args),
        ``halo_q`` = (peer table, element offset) of q, ``givens`` = (k, rcol, cs, y, mailbox offset) or None"""
        if self.timer is not None:
            tm, self.timer = self.timer, None
            tm.bracket("orth", (q.numel(), int(nv), 1, 100, True),      # algo 100: the fused two-kernel step
                       lambda: self.dist_update_scale(V, nv, q, vnext, h_ptr, nrm, halo, halo_q, halo_base, givens))
            self.timer = tm
            return
        c = self.comm
        _, _, hp, ho, nhalo, dst = halo
        q_tab, q_off = halo_q
        if givens is None:
            gk, rcol, cs, y, off = -1, None, None, None, 0
        else:
            gk, rcol, cs, y, off = givens
        check(self.lib.kry_dist_update_scale(
            self.h, code(q), q.numel(), V.data_ptr(), V.stride(0), int(nv), q.data_ptr(), vnext.data_ptr(),
            h_ptr, nrm.data_ptr(), nhalo, q_tab, q_off, hp, ho, int(halo_base), dst, int(gk), _p(rcol), _p(cs), _p(y),
            int(off), c.world, c.rank, c.epoch_dev.data_ptr(), c.slots.peer_table.data_ptr(),
            c.flags.peer_table.data_ptr()))

    def lanczos_diag(self, vprev, vk, bdiag, q, pre_coef, h3, vnext):
        """fused Lanczos step for a diagonal inner-product matrix (kry_lanczos_diag)"""
        if self.comm is not None:
            c = self.comm
            check(self.lib.kry_lanczos_diag_dist(
                self.h, code(q), q.numel(), _p(vprev), vk.data_ptr(), bdiag.data_ptr(), q.data_ptr(), _p(pre_coef),
                h3.data_ptr(), _p(vnext), c.world, c.rank, c.epoch_dev.data_ptr(), c.slots.peer_table.data_ptr(),
                c.flags.peer_table.data_ptr()))
            return
        check(self.lib.kry_lanczos_diag(self.h, code(q), q.numel(), _p(vprev), vk.data_ptr(), bdiag.data_ptr(),
                                        q.data_ptr(), _p(pre_coef), h3.data_ptr(), _p(vnext)))

    def alloc_basis(self, rows, N, dtype, op=None):
        """(rows, ld) storage for a vector-major basis.  Row-partitioned runs place it in a
        peer-mapped region with room for the halo behind each row, so the SpMV exchange reads
        the peers' basis rows in place (no staging copy)."""
        ext = getattr(op, "_ext_len", None) if op is not None else None
        if self.comm is not None and ext is not None:
            ld = (max(int(ext), int(N)) + 31) // 32 * 32
            return self.comm.shared_basis(int(rows), ld, dtype)
        ld = (int(N) + 31) // 32 * 32
        return self.empty((int(rows), ld), dtype)

    @realviews
    def project(self, W, V, d, a, Q, R, iterations, c_first):
        if self.comm is not None:
            if self._tmpc is None:
                self._tmpc = self.scalars(64)
            if int(d) > 32:
                raise NotImplementedError("row-partitioned projector supports d <= 32")
            tmp, tmp2 = self._tmpc[:32], self._tmpc[32:]
            for it in range(int(iterations)):
                self.block_dot(W, d, a, tmp, 0, None)
                if it == 0 and c_first is not None:
                    c_first[: int(d)].copy_(tmp[: int(d)])
                coef = tmp
                if Q is not None:
                    check(self.lib.kry_small_qr_apply(self.h, int(d), Q.data_ptr(), R.data_ptr(), tmp.data_ptr(),
                                                      tmp2.data_ptr()))
                    coef = tmp2
                self.block_axpy(V, d, coef, -1.0, a)
            return
        check(self.lib.kry_project(self.h, code(a), a.numel(), W.data_ptr(), W.stride(0), V.data_ptr(),
                                   V.stride(0), int(d), a.data_ptr(), _p(Q), _p(R), int(iterations),
                                   _p(c_first)))

    # ---- small recurrences -------------------------------------------------
    def givens_update(self, k, hcol, rcol, cs, y, off=0):
        check(self.lib.kry_givens_update(self.h, int(k), hcol.data_ptr(), rcol.data_ptr(), cs.data_ptr(),
                                         y.data_ptr(), int(off)))

    def tri_solve(self, k, R, y, out):
        check(self.lib.kry_tri_solve(self.h, int(k), R.data_ptr(), R.stride(0), y.data_ptr(), out.data_ptr()))

    def tri_solve_t(self, k, Rt, y, out):
        """tri_solve with R held column after column on the device (row j of ``Rt`` = column j of R)"""
        check(self.lib.kry_tri_solve_t(self.h, int(k), Rt.data_ptr(), Rt.stride(0), y.data_ptr(), out.data_ptr()))

    def givens_update_z(self, k, hcol, rcol, cs, y, off=0):
        """complex twin of givens_update: hcol/rcol/y hold k+2 interleaved complex numbers,
        cs 4 doubles per rotation"""
        check(self.lib.kry_givens_update_z(self.h, int(k), hcol.data_ptr(), rcol.data_ptr(), cs.data_ptr(),
                                           y.data_ptr(), int(off)))

    def tri_solve_z(self, k, R, y, out):
        """R: (k, 2k) float64 = k x k complex row-major, interleaved"""
        check(self.lib.kry_tri_solve_z(self.h, int(k), R.data_ptr(), R.stride(0) // 2, y.data_ptr(),
                                       out.data_ptr()))

    def minres_recur(self, k, h3, st, shift=1, off=0):
        check(self.lib.kry_minres_recur(self.h, int(k), h3.data_ptr(), st.data_ptr(), int(shift), int(off)))

    @realviews
    def minres_update(self, v, w0, w1, yk, st):
        check(self.lib.kry_minres_update(self.h, code(v), v.numel(), v.data_ptr(), w0.data_ptr(),
                                         w1.data_ptr(), yk.data_ptr(), st.data_ptr()))

    @realviews
    def cg_update(self, Ap, p, yk, r, z, dinv, rho, pAp, off=0):
        check(self.lib.kry_cg_update(self.h, code(p), p.numel(), Ap.data_ptr(), p.data_ptr(), yk.data_ptr(),
                                     r.data_ptr(), _p(z), _p(dinv), float(rho), pAp.data_ptr(), int(off)))

    @realviews
    def cg_update_dev(self, Ap, p, yk, r, z, dinv, st):
        """kry_cg_update with the scalars in device memory (st: see include/krypy_b200.h)"""
        check(self.lib.kry_cg_update_dev(self.h, code(p), p.numel(), Ap.data_ptr(), p.data_ptr(), yk.data_ptr(),
                                         r.data_ptr(), _p(z), _p(dinv), st.data_ptr()))

    def cg_scalars(self, st, off=0):
        """new rho (global on row-partitioned runs), shift, beta; publishes to the mailbox"""
        c = self.comm
        if c is not None:
            check(self.lib.kry_cg_scalars(self.h, st.data_ptr(), int(off), c.world, c.rank, c.epoch_dev.data_ptr(),
                                          c.slots.peer_table.data_ptr(), c.flags.peer_table.data_ptr()))
        else:
            check(self.lib.kry_cg_scalars(self.h, st.data_ptr(), int(off), 1, 0, None, None, None))

    @realviews
    def xpby_dev(self, x, beta, y, out):
        """out = x + beta[0] * y with beta in device memory"""
        check(self.lib.kry_xpby_dev(self.h, code(x), x.numel(), x.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                    out.data_ptr()))

    # ---- block kernels (projector set-up) -----------------------------------------
    @staticmethod
    def gram_fits(kx, ky, same):
        return ((kx if same else kx + ky) <= 64) and ((kx + 3) // 4) * ((ky + 3) // 4) <= 32

    @realviews
    def gram(self, X, kx, Y, ky, out):
        """out[i*ky + j] = <X_i, Y_j> in one pass over both blocks (kry_gram)"""
        check(self.lib.kry_gram(self.h, code(X), X.shape[1], X.data_ptr(), X.stride(0), int(kx), Y.data_ptr(),
                                Y.stride(0), int(ky), out.data_ptr()))

    @realviews
    def block_trsm(self, X, d, R, Q):
        """Q = X R^-1 (R upper triangular d x d device matrix, row-major); Q may be X"""
        check(self.lib.kry_block_trsm(self.h, code(X), X.shape[1], X.data_ptr(), X.stride(0), int(d), R.data_ptr(),
                                      Q.data_ptr(), Q.stride(0)))
