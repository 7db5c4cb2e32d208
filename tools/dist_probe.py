"""Probe of the row-partitioned GMRES step: pure GPU time of the CUDA graph of step k (replayed
back to back) vs the end-to-end per-iteration time (host in the loop).  torchrun -n P."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
lr = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(lr)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank = dist.get_rank() if world > 1 else 0
import krypy_b200 as kp
from krypy_b200 import dist as kd, problems, _device
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
N = n * n
if world > 1:
    part = kd.RowPartition(N, world, rank)
    A = problems.laplace2d(n, rows=(part.lo, part.hi)); b = problems.rhs_normal(N)[part.lo:part.hi]
    ls = kd.DistLinearSystem(A, b, part)
else:
    ls = kp.linsys.LinearSystem(problems.laplace2d(n), problems.rhs_normal(N))
ws = kp.utils.SolverWorkspace(graphs="on")
x = None
def cycle(x):
    try: s = kp.linsys.Gmres(ls, x0=x, maxiter=30, tol=1e-12, ortho="cgs", _workspace=ws)
    except kp.utils.ConvergenceError as e: s = e.solver
    return s
for _ in range(3):
    s = cycle(x); x = s.__dict__["_xk_dev"].reshape(-1)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    s = cycle(x); x = s.__dict__["_xk_dev"].reshape(-1)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
if rank == 0: print("n=%d world=%d: cycle %.3f ms -> %.1f us/iteration end to end" % (n, world, dt * 1e3, dt * 1e6 / 30))
for k in (2, 10, 20, 28):
    g = ws.graphs[k]
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): g.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("  graph step k=%2d: %.1f us GPU time per replay" % (k, e0.elapsed_time(e1) * 1e3 / 40))
# host cost of one replay call
t0 = time.perf_counter()
for _ in range(40): ws.graphs[10].replay()
t1 = time.perf_counter(); torch.cuda.synchronize()
if rank == 0: print("  host time per replay() call: %.1f us" % ((t1 - t0) * 1e6 / 40))
if world > 1: dist.barrier(); dist.destroy_process_group()
