"""N-rank worker (NCCL for set-up, NVLink peer kernels for the exchange): row-partitioned
solves against the oracle run on the undistributed system."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if torch.cuda.device_count() >= ws:
        torch.cuda.set_device(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    else:
        # fewer GPUs than ranks: all ranks share the available device(s) (CUDA IPC works between
        # processes on one GPU; the driver time-slices the spinning kernels) -- slow, but it
        # exercises the >2-rank protocol (middle ranks with two-sided halos) on a 1-GPU box
        torch.cuda.set_device(lr % torch.cuda.device_count())
        dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import krypy_b200 as kp
    from krypy_b200 import dist as kd, problems, _device
    from oracle import krylov_oracle as ko
    warnings.simplefilter("ignore")
    comm = kd.init()
    ctx = _device.Context.get()

    # ---- peer all-reduce / barrier ---------------------------------------------------------
    for n in (1, 5, 64, 100):
        x = torch.arange(n, dtype=torch.float64, device="cuda") * (rank + 1) + 0.25 * rank
        acc = torch.ones(n, dtype=torch.float64, device="cuda")
        comm.allreduce(x, n, post=0, acc=acc)
        ref = sum(np.arange(n) * (r + 1) + 0.25 * r for r in range(world))
        assert np.allclose(x.cpu().numpy(), ref, rtol=1e-15), (n, rank)
        assert np.allclose(acc.cpu().numpy(), 1 + ref, rtol=1e-15)
    for _ in range(50):
        comm.barrier()
    torch.cuda.synchronize()

    # ---- distributed SpMV (halo gather over peer memory) --------------------------------------
    for A in (problems.laplace2d(37), problems.convdiff2d(23), problems.poisson3d(9)):
        N = A.shape[0]
        part = kd.RowPartition(N, world, rank)
        op = kd.DistCsrOperator(kd.local_rows(A, part), part)
        x = np.random.default_rng(3).standard_normal(N)
        xd = ctx.to_block(x[part.lo:part.hi], torch.float64)
        for rep in range(3):                       # staging buffers alternate
            yd = op._apply_dev(xd)
            assert np.array_equal(ctx.to_numpy(yd)[:, 0], (A @ x)[part.lo:part.hi]), rank

    # ---- solvers: histories vs the oracle on the global system ---------------------------------
    def check(got, ref, rtol=1e-10):
        got, ref = np.array(got), np.array(ref)
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert np.all(np.abs(got - ref) <= rtol * np.abs(ref) + 1e-13), np.max(np.abs(got - ref) / ref)

    n = 40
    A = problems.laplace2d(n)
    N = n * n
    b = problems.rhs_normal(N)
    part = kd.RowPartition(N, world, rank)
    ls = kd.DistLinearSystem(kd.local_rows(A, part), b[part.lo:part.hi], part)
    for ortho in ("cgs", "mgs", "cgs2"):
        # (cgs: five cycles -- per-step graphs in the second, the whole-cycle graph from the third on, each
        # launching its successor speculatively)
        nrest = 4 if ortho == "cgs" else 2
        try:
            sol = kp.linsys.RestartedGmres(ls, maxiter=20, max_restarts=nrest, tol=1e-12, ortho=ortho)
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        try:
            ref = ko.restarted_gmres(ko.System(A, b), maxiter=20, max_restarts=nrest, tol=1e-12)
        except ko.OracleConvergenceError as e:
            ref = e.result
        check(sol.resnorms, ref.resnorms)
        assert np.abs(sol.xk[:, 0] - ref.xk[part.lo:part.hi, 0]).max() < 1e-9 * np.abs(ref.xk).max()

    # the one-wait step (kry_dist_dot with <w,w> + kry_dist_update_scale: norm from <w,w> - sum c^2, Givens
    # update in the extra CTA) and the four-kernel sequence give the same history to rounding
    hist = {}
    for fused in (True, False):
        comm.fused_step = fused
        try:
            sol = kp.linsys.Gmres(ls, maxiter=30, tol=1e-14, ortho="cgs")
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        hist[fused] = (np.array(sol.resnorms), sol.xk[:, 0].copy(), sol.arnoldi.H.copy())
    comm.fused_step = True
    check(hist[True][0], hist[False][0])
    assert np.abs(hist[True][1] - hist[False][1]).max() < 1e-10 * np.abs(hist[False][1]).max()
    assert np.abs(hist[True][2] - hist[False][2]).max() < 1e-11 * np.abs(hist[False][2]).max()
    try:
        ref = ko.gmres(ko.System(A, b), maxiter=30, tol=1e-14)
    except ko.OracleConvergenceError as e:
        ref = e.result
    check(hist[True][0], ref.resnorms)

    # cancellation guard of the fused step: three distinct eigenvalues, so A v_2 lies in span(v_0, v_1, v_2)
    # and <w,w> - sum c^2 is pure rounding -- the kernel must take the exact norm (a second exchange inside
    # the kernel), which makes the host see the invariant subspace exactly like the reference does
    import scipy.sparse as sp
    N = 3001
    D = sp.diags(1.0 + (np.arange(N) % 3)).tocsr()
    b = problems.rhs_normal(N)
    part = kd.RowPartition(N, world, rank)
    lsd = kd.DistLinearSystem(kd.local_rows(D, part), b[part.lo:part.hi], part)
    sol = kp.linsys.Gmres(lsd, maxiter=10, tol=1e-12, ortho="cgs")
    ref = ko.gmres(ko.System(D, b), maxiter=10, tol=1e-12)
    check(sol.resnorms, ref.resnorms)
    assert len(sol.resnorms) == 4 and sol.resnorms[-1] < 1e-13
    assert np.abs(sol.xk[:, 0] - ref.xk[part.lo:part.hi, 0]).max() < 1e-12

    A = problems.poisson3d(12)
    N = A.shape[0]
    b = problems.rhs_normal(N)
    part = kd.RowPartition(N, world, rank)
    Mj = problems.jacobi_csr(A)
    ls = kd.DistLinearSystem(kd.local_rows(A, part), b[part.lo:part.hi], part,
                             M=Mj[part.lo:part.hi, part.lo:part.hi], self_adjoint=True, positive_definite=True)
    sol = kp.linsys.Cg(ls, tol=1e-8, maxiter=200)
    ref = ko.cg(ko.System(A, b, M=Mj), tol=1e-8, maxiter=200)
    check(sol.resnorms, ref.resnorms, rtol=1e-9)

    A, B = problems.shifted_laplace_B(48, sigma=0.3, dtype=np.float64)
    N = A.shape[0]
    b = problems.rhs_normal(N)
    part = kd.RowPartition(N, world, rank)
    ls = kd.DistLinearSystem(kd.local_rows(A, part), b[part.lo:part.hi], part,
                             ip_B=B[part.lo:part.hi, part.lo:part.hi], self_adjoint=True)
    try:
        sol = kp.linsys.Minres(ls, tol=1e-9, maxiter=30)
    except kp.utils.ConvergenceError as e:
        sol = e.solver
    try:
        ref = ko.minres(ko.System(A, b, B=B), tol=1e-9, maxiter=30)
    except ko.OracleConvergenceError as e:
        ref = e.result
    check(sol.resnorms, ref.resnorms)

    torch.cuda.synchronize()
    dist.barrier()
    kd.shutdown()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
