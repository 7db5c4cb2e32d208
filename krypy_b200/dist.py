"""Row-partitioned multi-GPU execution: one process per
GPU (torchrun), contiguous row blocks of A, b, x and of every basis vector.

The reference has no distributed path; this module adds one *behind the same
solver classes*: after ``dist.init()`` every inner product / norm computed by the
device layer becomes a global sum and a ``DistCsrOperator`` applies the row block
of A to a distributed vector.  Exchange steps run in our own kernels over NVLink
peer memory (csrc/kry_peer.cu):

  * SpMV exchange  : ``kry_halo_gather`` -- P2P loads of exactly the remote entries
    of x that the local rows reference, straight out of the peers' basis rows (CUDA
    IPC mapped), written behind the local segment so the SpMV kernel sees one
    contiguous extended vector.  For a banded matrix this moves 2*bandwidth values
    instead of the all-gather's N.
  * reductions     : ``kry_peer_allreduce`` -- P2P stores of the partial sums into
    every peer + release/acquire flags, fixed rank-order sum (bitwise identical on
    all ranks, so the replicated Givens/convergence logic cannot diverge).

``torch.distributed`` is used only at set-up (exchange of IPC handles) and for the
optional NCCL all-gather comparison path (``KRY_DIST_EXCHANGE=allgather``).

Host-side planning (RowPartition, HaloPlan) is plain numpy and is covered by the
world_size-2 gloo tests on CPU.
"""
import ctypes
import os
import weakref

import numpy as np

from . import _device, linsys, utils
from ._lib import check

__all__ = ["RowPartition", "HaloPlan", "PeerComm", "DistCsrOperator", "DistLinearSystem",
           "DistGmres", "DistCg", "DistMinres", "init", "shutdown", "local_rows"]


def _roundup(x, a):
    return (x + a - 1) // a * a


class RowPartition(object):
    """Contiguous row blocks of equal size ``block`` (a multiple of 32; the last block may be
    shorter or empty).  rank r owns rows [r*block, min((r+1)*block, N))."""

    def __init__(self, N, world, rank):
        self.N, self.world, self.rank = int(N), int(world), int(rank)
        self.block = _roundup((self.N + self.world - 1) // self.world, 32)
        self.lo = min(self.rank * self.block, self.N)
        self.hi = min(self.lo + self.block, self.N)
        self.nloc = self.hi - self.lo

    def owner(self, cols):
        return np.asarray(cols) // self.block

    def bounds(self, r):
        lo = min(r * self.block, self.N)
        return lo, min(lo + self.block, self.N)


def local_rows(A, part):
    """rows [lo, hi) of a global scipy sparse matrix, global column indices"""
    return A.tocsr()[part.lo:part.hi, :]


class HaloPlan(object):
    """What a rank needs from its peers to apply its row block (pure numpy).

    ``A_rows``: scipy CSR of shape (nloc, N) with GLOBAL column indices.
    Produces the local matrix with columns renumbered into the extended vector
    ``[ x_local (block entries) | halo (nhalo entries) ]`` -- entry order inside each
    row is unchanged, so the row sums are bitwise those of the global matrix --
    and, per halo entry, the owning rank and the offset inside that rank's block.
    """

    def __init__(self, A_rows, part):
        import scipy.sparse as sp
        A_rows = sp.csr_matrix(A_rows)
        assert A_rows.shape == (part.nloc, part.N), (A_rows.shape, part.nloc, part.N)
        cols = A_rows.indices.astype(np.int64)
        remote = (cols < part.lo) | (cols >= part.hi)
        self.halo_cols = np.unique(cols[remote])
        self.nhalo = int(self.halo_cols.shape[0])
        owner = part.owner(self.halo_cols)
        self.halo_peer = owner.astype(np.int32)
        self.halo_off = (self.halo_cols - owner * part.block).astype(np.int32)
        self.nloc, self.block = part.nloc, part.block
        self.ext = part.block + self.nhalo                      # length of the extended vector
        new_cols = np.where(remote, part.block + np.searchsorted(self.halo_cols, cols), cols - part.lo)
        self.indptr = A_rows.indptr.astype(np.int32)
        self.indices = new_cols.astype(np.int32)
        self.data = A_rows.data

    def local_matrix(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=(self.nloc, self.ext))


def extend_vector(plan, part, x_blocks):
    """numpy reference of the halo exchange (used by the CPU tests)."""
    xe = np.zeros(plan.ext, dtype=np.result_type(*[b.dtype for b in x_blocks]))
    xe[: part.nloc] = x_blocks[part.rank]
    for i in range(plan.nhalo):
        xe[part.block + i] = x_blocks[plan.halo_peer[i]][plan.halo_off[i]]
    return xe


# ---------------------------------------------------------------------------------------
# peer memory
# ---------------------------------------------------------------------------------------
class _RawCuda(object):
    """__cuda_array_interface__ holder so torch can view memory we cudaMalloc'ed."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False),
                                         "version": 2}


class SharedRegion(object):
    """A cudaMalloc'ed buffer every rank has mapped: base (mine), peer pointer table (device)."""

    def __init__(self, comm, nbytes):
        t = _device.torch()
        ctx = comm.ctx
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        check(ctx.lib.kry_peer_alloc(ctx.h, self.nbytes, ctypes.byref(p)))
        self.base = int(p.value)
        buf = ctypes.create_string_buffer(64)
        check(ctx.lib.kry_ipc_export(ctx.h, ctypes.c_void_p(self.base), buf))
        handles = comm.all_gather_object(bytes(buf.raw))
        self.peer = []
        self._opened = []
        for r, h in enumerate(handles):
            if r == comm.rank:
                self.peer.append(self.base)
            else:
                q = ctypes.c_void_p()
                check(ctx.lib.kry_ipc_open(ctx.h, ctypes.c_char_p(h), ctypes.byref(q)))
                self.peer.append(int(q.value))
                self._opened.append(int(q.value))
        self.peer_table = t.tensor(self.peer, dtype=t.int64, device=ctx.device)
        self.bytes_view = t.as_tensor(_RawCuda(self.base, self.nbytes), device=ctx.device)

    def view(self, dtype):
        return self.bytes_view.view(dtype)

    def contains(self, ptr, nbytes):
        return self.base <= ptr and ptr + nbytes <= self.base + self.nbytes


class PeerComm(object):
    """Global sums and barriers over NVLink peer memory + the shared-region pool."""

    def __init__(self, ctx, world, rank):
        import torch.distributed as td
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        self._td = td
        self.epoch_dev = _device.torch().zeros(1, dtype=_device.torch().int64, device=ctx.device)
        self.regions = []
        self._free = {}
        self.slots = SharedRegion(self, 2 * self.world * 64 * 8)
        self.flags = SharedRegion(self, max(self.world, 8) * 8)
        self.regions = []
        self.exchange = os.environ.get("KRY_DIST_EXCHANGE", "halo")     # halo | allgather
        self.reduce = os.environ.get("KRY_DIST_REDUCE", "peer")         # peer | nccl
        # split: dot / update / scale kernels with the exchange fused in (measured faster on 2-8 B200:
        # 6603 vs 6301 it/s on C2 at 8 GPUs); coop: ONE cooperative kernel per Gram-Schmidt step
        self.orth_mode = os.environ.get("KRY_DIST_ORTH", "split")
        self.halo_ready = None      # data_ptr of the basis row whose halo the last orth step already gathered
        # gather the halo of v_{k+1} from the peers' un-normalised q (one cross-GPU wait less per
        # Arnoldi step; KRY_DIST_HALO_FROM_Q=0: from their v_{k+1} rows after a second handshake)
        self.halo_from_q = os.environ.get("KRY_DIST_HALO_FROM_Q", "1") not in ("0", "")
        # block-CGS Arnoldi step with ONE cross-GPU wait (kry_dist_dot with <w,w> + kry_dist_update_scale: norm
        # from <w,w> - sum c^2 with an exact-norm guard; update + normalised store + halo + Givens in one kernel);
        # KRY_DIST_FUSED=0: dot / update / scale+halo / Givens as four kernels with two waits
        self.fused_step = os.environ.get("KRY_DIST_FUSED", "1") not in ("0", "")
        self.barrier_sync()

    def all_gather_object(self, obj):
        out = [None] * self.world
        self._td.all_gather_object(out, obj)
        return out

    def barrier_sync(self):
        _device.torch().cuda.synchronize()
        self._td.barrier()

    # -- device-side collectives (stream ordered, no host sync) --
    def allreduce(self, x, n, post=0, acc=None):
        """x[0:n] <- global sum (n may exceed 64: chunked); post/acc as kry_block_dot"""
        ctx = self.ctx
        if self.reduce == "nccl":
            t = _device.torch()
            self._td.all_reduce(x[:n])
            if post == 1:
                x[:n].abs_().sqrt_()
            if acc is not None:
                a = acc if not isinstance(acc, int) else None
                if a is None:
                    raise NotImplementedError("raw accumulator pointer with the NCCL reduction path")
                a[:n].add_(x[:n])
            return
        xp = x if isinstance(x, int) else x.data_ptr()
        ap = None if acc is None else (acc if isinstance(acc, int) else acc.data_ptr())
        i = 0
        while i < n:
            c = min(64, n - i)
            check(ctx.lib.kry_peer_allreduce(ctx.h, self.world, self.rank, self.epoch_dev.data_ptr(), c, xp + 8 * i,
                                             self.slots.peer_table.data_ptr(), self.flags.peer_table.data_ptr(),
                                             int(post), None if ap is None else ap + 8 * i))
            i += c

    def barrier(self):
        ctx = self.ctx
        if self.reduce == "nccl":
            self._td.barrier()
            return
        check(ctx.lib.kry_peer_barrier(ctx.h, self.world, self.rank, self.epoch_dev.data_ptr(),
                                       self.slots.peer_table.data_ptr(), self.flags.peer_table.data_ptr()))

    # -- shared-region pool (collective: every rank must call in the same order) --
    def get_region(self, nbytes):
        nbytes = _roundup(int(nbytes), 512)
        lst = self._free.get(nbytes)
        if lst:
            return lst.pop()
        reg = SharedRegion(self, nbytes)
        self.regions.append(reg)
        return reg

    def release_region(self, reg):
        self._free.setdefault(reg.nbytes, []).append(reg)

    def find_region(self, ptr, nbytes):
        for reg in self.regions:
            if reg.contains(ptr, nbytes):
                return reg
        return None

    def shared_basis(self, rows, ld, dtype):
        """(rows, ld) tensor in a peer-mapped region; the region returns to the pool when the
        tensor is garbage collected."""
        es = 8 if dtype == _device.torch().float64 else 4
        reg = self.get_region(rows * ld * es)
        ten = reg.view(dtype)[: rows * ld].view(rows, ld)
        weakref.finalize(ten, self.release_region, reg)
        return ten


_COMM = None


def init(world=None, rank=None):
    """Attach a PeerComm to this process's device context: from now on every reduction of the
    device layer is a global sum.  Requires an initialised torch.distributed process group."""
    global _COMM
    import torch.distributed as td
    if not td.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    ctx = _device.Context.get()
    world = td.get_world_size() if world is None else world
    rank = td.get_rank() if rank is None else rank
    if _COMM is None:
        _COMM = PeerComm(ctx, world, rank)
        ctx.comm = _COMM
    return _COMM


def shutdown():
    global _COMM
    if _COMM is not None:
        _COMM.ctx.comm = None
        _COMM = None


# ---------------------------------------------------------------------------------------
# distributed operator
# ---------------------------------------------------------------------------------------
class DistCsrOperator(utils._DeviceOperator):
    """Row block of a global sparse matrix acting on row-partitioned vectors.

    ``shape`` is reported as ``(nloc, nloc)`` so the (local) LinearSystem machinery accepts it;
    ``N_global`` holds the true dimension."""

    def __init__(self, A_rows, part, comm=None, plan=None):
        """``plan``: the HaloPlan of an earlier operator over the SAME sparsity pattern and partition
        (host-side symbolic work: which remote entries each row block needs); the numeric values are
        taken from ``A_rows``."""
        self.part = part
        self.comm = comm if comm is not None else _COMM
        if self.comm is None:
            raise RuntimeError("call krypy_b200.dist.init() first")
        if plan is not None:
            import copy
            import scipy.sparse as sp
            A_rows = sp.csr_matrix(A_rows)
            if A_rows.shape != (part.nloc, part.N) or A_rows.nnz != plan.indices.shape[0]:
                raise utils.ArgumentError("plan does not belong to this matrix / partition")
            plan = copy.copy(plan)
            plan.data = A_rows.data
        self.plan = plan if plan is not None else HaloPlan(A_rows, part)
        self.N_global = part.N
        super(DistCsrOperator, self).__init__((part.nloc, part.nloc), A_rows.dtype)
        # peers address each other's basis rows with ONE element offset (row index * leading
        # dimension), so the leading dimension must be identical on every rank: size the halo part
        # of a row for the largest halo of any rank
        ext = getattr(self.plan, "ext_len_all", None)
        if ext is None:
            ext = self.plan.ext_len_all = part.block + max(self.comm.all_gather_object(int(self.plan.nhalo)))
        self._ext_len = ext
        self._devcache = {}
        self._xbuf = {}
        self._napply = 0

    def _dev(self, td):
        obj = self._devcache.get(td)
        if obj is None:
            ctx = self.comm.ctx
            t = _device.torch()
            npdt = _device.torch_to_np_dtype(td)
            pl = self.plan
            A = _device.CsrDev(t.from_numpy(pl.indptr).to(ctx.device), t.from_numpy(pl.indices).to(ctx.device),
                               t.from_numpy(np.ascontiguousarray(pl.data, dtype=npdt)).to(ctx.device),
                               (pl.nloc, pl.ext))
            hp = t.from_numpy(pl.halo_peer).to(ctx.device)
            ho = t.from_numpy(pl.halo_off).to(ctx.device)
            obj = (A, hp, ho)
            self._devcache[td] = obj
        return obj

    def _exchange_buffer(self, td):
        """two alternating peer-mapped staging vectors for inputs that do not live in a shared basis"""
        key = (td, self._napply & 1)
        buf = self._xbuf.get(key)
        if buf is None:
            buf = self.comm.shared_basis(1, _roundup(self._ext_len, 32), td)
            self._xbuf[key] = buf
        return buf

    def _halo_args(self, row):
        """arguments for a kernel that gathers the halo of the basis row ``row`` (1-D view) in place,
        or None when the row does not live in a peer-mapped region"""
        comm, pl = self.comm, self.plan
        es = row.element_size()
        reg = comm.find_region(row.data_ptr(), pl.ext * es)
        if reg is None or comm.reduce != "peer":
            return None
        A, hp, ho = self._dev(row.dtype)
        off = (row.data_ptr() - reg.base) // es
        return (reg.peer_table.data_ptr(), off, hp.data_ptr(), ho.data_ptr(), pl.nhalo,
                row.data_ptr() + pl.block * es)

    def _alloc_vec(self, td):
        """a (1, nloc) block whose storage is a peer-mapped extended vector: applying the operator
        to it needs no staging copy (its halo is gathered in place)"""
        buf = self.comm.shared_basis(1, _roundup(self._ext_len, 32), td)
        v = buf[:, : self.plan.nloc]
        v._keep = buf        # the region returns to the pool when `buf` is collected
        return v

    def _apply_dot_dev(self, p, Ap, pAp):
        """Ap = A p and pAp[0] = <p, Ap> (global) with the dot in the SpMV epilogue (CG, linsys.py:631-634)"""
        self._apply_dev(p, out=Ap, dot_out=pAp)

    def _halo_src_args(self, vec):
        """(peer pointer table, element offset) of a vector that lives in a peer-mapped region (the
        gather SOURCE of a halo exchange), or None"""
        comm, pl = self.comm, self.plan
        es = vec.element_size()
        reg = comm.find_region(vec.data_ptr(), pl.nloc * es)
        if reg is None or comm.reduce != "peer":
            return None
        return reg.peer_table.data_ptr(), (vec.data_ptr() - reg.base) // es

    def _extended(self, x, hp, ho):
        """the extended vector [x | halo] of a local segment ``x`` (1-D view) with its halo in place: ``x`` is
        staged into a peer-mapped buffer if it does not live in one, and the halo is gathered unless the
        Gram-Schmidt step that produced ``x`` did that already"""
        comm, ctx, pl = self.comm, self.comm.ctx, self.plan
        es = x.element_size()
        td = x.dtype
        reg = comm.find_region(x.data_ptr(), pl.ext * es)
        if reg is None:
            comm.halo_ready = None
            buf = self._exchange_buffer(td)
            self._napply += 1
            ctx.axpby(1.0, x, 0.0, None, buf[0][: pl.nloc])
            x = buf[0]
            reg = comm.find_region(x.data_ptr(), pl.ext * es)
        off = (x.data_ptr() - reg.base) // es
        xext = reg.view(td)[off: off + pl.ext]
        if comm.halo_ready is not None and comm.halo_ready == x.data_ptr():
            comm.halo_ready = None           # the Gram-Schmidt step that produced x gathered its halo already
        elif comm.reduce == "peer":
            # one kernel: flag handshake (every rank's segment of this vector is complete) + P2P gather
            check(ctx.lib.kry_dist_halo(ctx.h, _device.code(xext), pl.nhalo, reg.peer_table.data_ptr(), off,
                                        hp.data_ptr(), ho.data_ptr(), xext.data_ptr() + pl.block * es,
                                        comm.world, comm.rank, comm.epoch_dev.data_ptr(),
                                        comm.slots.peer_table.data_ptr(), comm.flags.peer_table.data_ptr()))
        else:
            comm.barrier()
            if pl.nhalo:
                check(ctx.lib.kry_halo_gather(ctx.h, _device.code(xext), pl.nhalo, reg.peer_table.data_ptr(),
                                              off, hp.data_ptr(), ho.data_ptr(), None,
                                              xext.data_ptr() + pl.block * es))
        return xext

    def _apply_dev(self, Xd, out=None, adj=False, dot_out=None):
        if adj:
            raise utils.LinearOperatorError("dot_adj undefined for a row-partitioned operator")
        comm, ctx, pl = self.comm, self.comm.ctx, self.plan
        A, hp, ho = self._dev(Xd.dtype)
        k = Xd.shape[0]
        es = Xd.element_size()
        if out is None:
            out = ctx.empty((k, pl.nloc), Xd.dtype)
        for j in range(k):
            xext = self._extended(Xd[j], hp, ho)
            if dot_out is not None:
                ctx.spmv(A, xext, out[j], w=xext[: pl.nloc], dot_out=dot_out)   # (+ peer all-reduce of the dot)
            else:
                ctx.spmv(A, xext, out[j])
        return out


class DistLinearSystem(linsys.LinearSystem):
    """LinearSystem over a row partition: ``A_rows`` are this rank's rows (scipy CSR, global
    column indices), ``b`` this rank's segment of the right-hand side."""

    def __init__(self, A_rows, b, part, **kwargs):
        comm = init()
        op = A_rows if isinstance(A_rows, DistCsrOperator) else DistCsrOperator(A_rows, part, comm)
        self.part = part
        self.N_global = part.N
        super(DistLinearSystem, self).__init__(op, b, **kwargs)
        if np.dtype(self.dtype).kind == "c":
            raise NotImplementedError("complex row-partitioned systems are not implemented (single GPU only)")


class DistGmres(linsys.Gmres):
    pass


class DistCg(linsys.Cg):
    pass


class DistMinres(linsys.Minres):
    pass
