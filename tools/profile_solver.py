"""Where does an iteration go?  Runs one BASELINE configuration (single GPU, or one rank's view under
torchrun) under torch.profiler (CUPTI kernel records: per-kernel device time, launch counts) and
cProfile (host time of the Python layer), and prints both next to the wall time per iteration.
    python tools/profile_solver.py c5 [--n 4000] [--maxiter 50]       (c2 | c3 | c5)
ANALYSIS TOOL (numbers under a profiler are not bench values)."""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--grid", dest="n", type=int, default=0)
    ap.add_argument("--maxiter", type=int, default=50)
    ap.add_argument("--host", action="store_true", help="cProfile of the host layer instead of the kernel table")
    a = ap.parse_args()
    warnings.simplefilter("ignore")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    import krypy_b200 as kp
    from krypy_b200 import problems
    kd = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        from krypy_b200 import dist as kd
        kd.init()
        rank = dist.get_rank()
    else:
        rank = 0

    def part_of(N):
        return kd.RowPartition(N, world, rank) if kd else None

    if a.config == "c5":
        n = a.n or 4000
        N = n * n
        pt = part_of(N)
        rows = (pt.lo, pt.hi) if pt else None
        A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32, rows=rows)
        b = problems.rhs_normal(N, dtype=np.float32)
        if pt:
            ls = kd.DistLinearSystem(A, b[pt.lo:pt.hi], pt, ip_B=B[:, pt.lo:pt.hi].tocsr(), self_adjoint=True, dtype=np.float32)
        else:
            ls = kp.linsys.LinearSystem(A, b, ip_B=B, self_adjoint=True, dtype=np.float32)
        run = lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=a.maxiter)
    elif a.config == "c3":
        n = a.n or 256
        N = n ** 3
        pt = part_of(N)
        rows = (pt.lo, pt.hi) if pt else None
        A = problems.poisson3d(n, rows=rows)
        b = problems.rhs_normal(N)
        import scipy.sparse as sp
        nl = A.shape[0]
        M = sp.diags(np.full(nl, 1.0 / 6.0)).tocsr()
        if pt:
            ls = kd.DistLinearSystem(A, b[pt.lo:pt.hi], pt, M=M, self_adjoint=True, positive_definite=True)
        else:
            ls = kp.linsys.LinearSystem(A, b, M=M, self_adjoint=True, positive_definite=True)
        run = lambda: kp.linsys.Cg(ls, tol=1e-30, maxiter=a.maxiter)
    elif a.config == "c4":
        n = a.n or 2000
        N = n * n
        A = problems.convdiff2d(n, c=0.1)
        b = np.ones((N, 1))
        ls = kp.linsys.LinearSystem(A, b)
        xs = torch.arange(1, n + 1, dtype=torch.float64, device="cuda") / (n + 1.0)
        rowsU = [torch.outer(torch.sin(p * np.pi * xs), torch.sin(q * np.pi * xs)).reshape(-1)
                 for p in range(1, 6) for q in range(1, 5)]
        Ublk = kp.utils.DeviceBlock(torch.stack(rowsU))                      # 20 sine modes, resident in HBM
        if a.maxiter == 0:
            run = lambda: kp.deflation.ObliqueProjection(ls, Ublk)            # the set-up alone
        else:
            run = lambda: kp.deflation.DeflatedGmres(ls, U=Ublk, maxiter=a.maxiter, tol=1e-30, ortho="cgs")
    else:
        n = a.n or 3162
        N = n * n
        pt = part_of(N)
        rows = (pt.lo, pt.hi) if pt else None
        A = problems.laplace2d(n, rows=rows)
        b = problems.rhs_normal(N)
        ls = kd.DistLinearSystem(A, b[pt.lo:pt.hi], pt) if pt else kp.linsys.LinearSystem(A, b)
        run = lambda: kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=max(a.maxiter // 30 - 1, 0), tol=1e-30, ortho="cgs")

    def go():
        try:
            return run()
        except kp.utils.ConvergenceError as e:
            return e.solver

    go()
    torch.cuda.synchronize()
    t = time.perf_counter()
    s = go()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    its = max(len(getattr(s, "resnorms", [0, 0])) - 1, 1)
    if rank == 0:
        print("%s world=%d: %d iterations, %.1f us/iteration wall (unprofiled)" % (a.config, world, its, 1e6 * dt / its))
    if a.host:
        pr = cProfile.Profile()
        pr.enable()
        go()
        torch.cuda.synchronize()
        pr.disable()
        if rank == 0:
            st = io.StringIO()
            pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
            print(st.getvalue()[:6000])
    else:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            go()
            torch.cuda.synchronize()
        if rank == 0:
            rows = {}
            for ev in prof.events():
                if ev.device_type == torch.autograd.DeviceType.CUDA:
                    r = rows.setdefault(ev.name[:70], [0, 0.0])
                    r[0] += 1
                    r[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            tot = sum(v[1] for v in rows.values())
            print("device busy %.1f us/iteration (%d kernel records)" % (tot / its, sum(v[0] for v in rows.values())))
            for nm, (c, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:16]:
                print("  %7.1f us/it  %5d x %8.1f us  %s" % (us / its, c, us / c, nm))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        kd.shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
