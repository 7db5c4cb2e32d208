"""How much of the multi-GPU GMRES(30) step is host pacing / per-step graph launches?  Replays ONE CUDA graph
holding a whole restart cycle (30 Arnoldi steps incl. the Givens updates) of config C2 on the ranks of a
torchrun job and compares the time per step with the solver's own per-step graph replays.  ANALYSIS TOOL.
    python -m torch.distributed.run --nproc-per-node 8 ... tools/probe_cycle_graph.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    import warnings
    warnings.simplefilter("ignore")
    import krypy_b200 as kp
    from krypy_b200 import dist as kd, problems, _device
    n, m = 3162, 30
    N = n * n
    kd.init()
    ctx = _device.Context.get()
    part = kd.RowPartition(N, world, rank)
    A = problems.laplace2d(n, rows=(part.lo, part.hi))
    b = problems.rhs_normal(N)[part.lo:part.hi]
    ls = kd.DistLinearSystem(A, b, part)
    ws = kp.utils.SolverWorkspace()

    def cycle(x0, r0):
        try:
            s = kp.linsys.Gmres(ls, x0=x0, maxiter=m, tol=1e-13, ortho="cgs", _workspace=ws, _x0_residual=r0, _prelaunch=True)
        except kp.utils.ConvergenceError as e:
            s = e.solver
        return s

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    x, r0 = None, None
    for _ in range(3):
        sol = cycle(x, r0)
        x, r0 = sol.__dict__["_xk_dev"].reshape(-1), sol.__dict__.get("_last_residual")
    # the solver's own pace (per-step graphs, look-ahead of one step)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ncyc = 10
    for _ in range(ncyc):
        sol = cycle(x, r0)
        x, r0 = sol.__dict__["_xk_dev"].reshape(-1), sol.__dict__.get("_last_residual")
    e1.record()
    sync_all()
    t_solver = e0.elapsed_time(e1) * 1e3 / (ncyc * m)

    # one graph for the whole cycle over the same buffers
    ar = sol.arnoldi
    bufs = {k[0]: v for k, v in ws.bufs.items()}
    rcol, cs, y = bufs["rcol"], bufs["cs"], bufs["y"]
    offs, o = [], 0
    for k in range(m):
        offs.append(o)
        o += 2 * (k + 2) + 1
    ctx.comm.halo_ready = None
    ar._hcol_store.zero_()
    sync_all()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ctx.use_current_stream()
        for k in range(m):
            if not ar._enqueue(k, givens=(rcol, cs, y, offs[k])):
                ctx.givens_update(k, ar._hcol, rcol, cs, y, offs[k])
    ctx.use_current_stream()
    for _ in range(2):
        g.replay()
    sync_all()
    e0.record()
    for _ in range(ncyc):
        g.replay()
    e1.record()
    sync_all()
    t_graph = e0.elapsed_time(e1) * 1e3 / (ncyc * m)
    t = torch.tensor([t_solver, t_graph], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world %d: solver (per-step graphs, incl. per-cycle overhead) %.1f us/step; whole-cycle graph %.1f us/step"
              % (world, t[0].item(), t[1].item()))
        print("   |y[m]| of the replayed cycle: %.6e (solver's last residual %.6e)"
              % (ctx.mailbox[offs[m - 1]], sol.resnorms[-1] * ls.MMlb_norm))
    kd.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
