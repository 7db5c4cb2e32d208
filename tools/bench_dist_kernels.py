"""Per-kernel GPU time of the row-partitioned Arnoldi step at the PER-RANK size of an 8-GPU run,
emulated on ONE GPU (world=1: the peer protocol publishes to itself).  Each kernel is captured in
its own CUDA graph and replayed back to back."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29677")
import numpy as np, torch, torch.distributed as dist
torch.cuda.set_device(0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
import krypy_b200 as kp
from krypy_b200 import dist as kd, problems, _device
from krypy_b200._lib import check, KRY_ORTH_CGS
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
rows = n * n // 8
comm = kd.init(); ctx = _device.Context.get(); lib = ctx.lib
N = n * n
part = kd.RowPartition(N, 8, 3)           # a middle rank of an 8-way partition
A = problems.laplace2d(n, rows=(part.lo, part.hi))
# single process: remap remote columns onto local ones so the halo gather reads valid (own) memory
pl = kd.HaloPlan(A, part)
part1 = kd.RowPartition(N, 8, 3); part1.world = 1
op = kd.DistCsrOperator.__new__(kd.DistCsrOperator)
kp.utils._DeviceOperator.__init__(op, (part.nloc, part.nloc), A.dtype)
op.part, op.comm, op.plan, op.N_global = part, comm, pl, N
pl.halo_peer[:] = 0                       # everything is "owned" by rank 0 = this process
op._ext_len, op._devcache, op._xbuf, op._napply = pl.ext, {}, {}, 0
nloc = part.nloc
m = 30
V = ctx.alloc_basis(m + 1, nloc, torch.float64, op)
Vd = V[:, :nloc]; Vd.normal_(); Vd.mul_(1.0 / np.sqrt(N))
q = torch.randn(1, nloc, dtype=torch.float64, device="cuda")
h = ctx.scalars(64); nrm = ctx.scalars(1)
w, r = 1, 0
ep, sl, fl = comm.epoch_dev.data_ptr(), comm.slots.peer_table.data_ptr(), comm.flags.peer_table.data_ptr()
ld, es = V.stride(0), 8

def graph_time(fn, reps=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ctx.use_current_stream(); fn()
    ctx.use_current_stream()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

print("per-rank rows %d (N=%d / 8), nhalo %d" % (nloc, N, pl.nhalo))
q2 = ctx.alloc_basis(2, nloc, torch.float64, op)
qq = q2[0:1, :nloc]; qq.normal_()
rcol, cs, y = ctx.scalars(64), ctx.scalars(128), ctx.scalars(64)
ks = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(0, 30))
tot = {"spmv": 0.0, "dot": 0.0, "k2": 0.0, "old": 0.0, "ideal": 0.0}
for k in ks:
    nv = k + 1
    vnext = Vd[k + 1]
    hal = op._halo_args(vnext)
    halq = op._halo_src_args(qq[0])
    t_spmv = graph_time(lambda: op._apply_dev(Vd[k:k + 1], out=qq))
    t_dot = graph_time(lambda: ctx.dist_dot_sq(Vd, nv, qq[0]))
    # (K2 alone replays against the flags / slots the last dot left behind: no wait, same arithmetic)
    t_k2 = graph_time(lambda: ctx.dist_update_scale(Vd, nv, qq[0], vnext, h.data_ptr(), h[nv:], hal, halq, pl.block,
                                                     givens=(k, rcol, cs, y, 0)))
    nohalo = (hal[0], hal[1], hal[2], hal[3], 0, hal[5])
    t_k2_plain = graph_time(lambda: ctx.dist_update_scale(Vd, nv, qq[0], vnext, h.data_ptr(), h[nv:], nohalo, halq, pl.block,
                                                           givens=None))
    t_k2_halo = graph_time(lambda: ctx.dist_update_scale(Vd, nv, qq[0], vnext, h.data_ptr(), h[nv:], hal, halq, pl.block,
                                                          givens=None))
    t_k2_giv = graph_time(lambda: ctx.dist_update_scale(Vd, nv, qq[0], vnext, h.data_ptr(), h[nv:], nohalo, halq, pl.block,
                                                         givens=(k, rcol, cs, y, 0)))
    t_upd_old = graph_time(lambda: check(lib.kry_dist_update(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), h.data_ptr(), 1, w, r, ep, sl, fl)))
    print("      K2 variants: plain %5.1f | +halo %5.1f | +givens %5.1f | old update kernel alone %5.1f" % (t_k2_plain, t_k2_halo, t_k2_giv, t_upd_old))
    def old():
        check(lib.kry_dist_dot(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), 0, w, r, ep, sl, fl))
        check(lib.kry_dist_update(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), h.data_ptr(), 1, w, r, ep, sl, fl))
        check(lib.kry_dist_scale(ctx.h, 1, nloc, q.data_ptr(), V.data_ptr() + (k + 1) * ld * es, nrm.data_ptr(), w, r, ep, sl, fl))
    t_old = graph_time(old)
    ideal = ((nv + 1) + (nv + 2)) * 8 * nloc / 6538.9e9 * 1e6
    ideal_spmv = (A.nnz * 12 + 4 * nloc + 16 * nloc) / 6538.9e9 * 1e6
    for key, v in (("spmv", t_spmv), ("dot", t_dot), ("k2", t_k2), ("old", t_old), ("ideal", ideal + ideal_spmv)):
        tot[key] += v
    print("k=%2d: halo+spmv %5.1f (ideal %4.1f) | dot+sq %5.1f | update_scale+halo+givens %5.1f | sum %5.1f (ideal %4.1f) | "
          "old dot+update+scale %5.1f" % (k, t_spmv, ideal_spmv, t_dot, t_k2, t_dot + t_k2, ideal, t_old))
print("sum over k: spmv %.1f  dot %.1f  k2 %.1f  => step total %.1f us (ideal %.1f); old orth kernels %.1f"
      % (tot["spmv"], tot["dot"], tot["k2"], tot["spmv"] + tot["dot"] + tot["k2"], tot["ideal"], tot["old"]))
dist.destroy_process_group()
