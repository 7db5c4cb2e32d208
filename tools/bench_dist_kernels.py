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
for k in (2, 10, 20, 28):
    nv = k + 1
    t_spmv = graph_time(lambda: op._apply_dev(Vd[k:k + 1], out=q))
    t_dot = graph_time(lambda: check(lib.kry_dist_dot(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), w, r, ep, sl, fl)))
    def upd():
        check(lib.kry_dist_dot(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), w, r, ep, sl, fl))
        check(lib.kry_dist_update(ctx.h, 1, nloc, V.data_ptr(), ld, nv, q.data_ptr(), h.data_ptr(), 1, w, r, ep, sl, fl))
    t_du = graph_time(upd)
    def full():
        upd()
        check(lib.kry_dist_scale(ctx.h, 1, nloc, q.data_ptr(), V.data_ptr() + (k + 1) * ld * es, nrm.data_ptr(), w, r, ep, sl, fl))
    t_full = graph_time(full)
    ideal_orth = (2 * nv + 5) * 8 * nloc / 6540.5e9 * 1e6
    ideal_spmv = (A.nnz * 12 + 4 * nloc + 16 * nloc) / 6540.5e9 * 1e6
    print("k=%2d: halo+spmv %.1f us (ideal %.1f) | dot %.1f | dot+update %.1f | dot+update+scale %.1f (ideal %.1f)"
          % (k, t_spmv, ideal_spmv, t_dot, t_du, t_full, ideal_orth))
dist.destroy_process_group()
