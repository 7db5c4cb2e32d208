"""BASELINE.json's multi-GPU configurations at FULL size, one process per GPU (torchrun):

  C3  CG + Jacobi, 3-D 7-point Poisson, N = 64,000,000, fp64           (1 and 8 B200)
  C5  MINRES(lanczos) with ip_B = B, 2-D shifted Laplacian, N = 16,000,000, fp32 storage (4 B200)

Rows are block-partitioned (krypy_b200.dist); each rank builds only its own rows.  The exchange
steps run over NVLink peer memory (halo gather + peer all-reduce kernels), NCCL only at set-up.
Reports iterations/s (device time, max over ranks), algorithmic GB/s per the SURVEY 8d byte model
(whole job) and size-independent parity properties: the history is identical on every rank, the
first entries equal the single-GPU values recorded in profiles/r1_configs_fullsize.json to 1e-10,
and the explicit residual of the assembled solution equals the last history entry.

usage (on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29631 tools/run_configs_dist.py c3 > profiles/r2_c3_8gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29632 tools/run_configs_dist.py c5 > profiles/r2_c5_4gpu.json
"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    which = [a.lower() for a in sys.argv[1:]] or ["c3"]
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    import krypy_b200 as kp
    from krypy_b200 import dist as kd, problems
    warnings.simplefilter("ignore")
    kd.init()
    single = {}
    p = os.path.join(ROOT, "profiles", "r1_configs_fullsize.json")
    if os.path.exists(p):
        with open(p) as f:
            single = json.load(f)
    out = {"n_gpus": world}

    def timed(fn):
        """device time of fn() in seconds, max over ranks; returns (solver, seconds)"""
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            s = fn()
        except kp.utils.ConvergenceError as e:
            s = e.solver
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return s, float(ms.item()) * 1e-3

    def same_on_all_ranks(hist):
        h = torch.tensor(hist, device="cuda", dtype=torch.float64)
        lo, hi = h.clone(), h.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))

    def global_sum(v):
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    def first_vs_single(hist, key):
        ref = single.get(key, {}).get("first4")
        if not ref:
            return None
        m = min(len(ref), len(hist))
        return float(np.max(np.abs(np.array(hist[:m]) - np.array(ref[:m])) / np.array(ref[:m])))

    if "c3" in which:
        n = 400
        N = n ** 3
        part = kd.RowPartition(N, world, rank)
        A = problems.poisson3d(n, rows=(part.lo, part.hi))
        b = problems.rhs_normal(N)[part.lo:part.hi]
        # Jacobi M = diag(A)^-1 as a CSR diagonal of this rank's rows (constant 1/6 here)
        import scipy.sparse as sp
        M = sp.diags(np.full(part.nloc, 1.0 / 6.0)).tocsr()
        ls = kd.DistLinearSystem(A, b, part, M=M, self_adjoint=True, positive_definite=True)
        timed(lambda: kp.linsys.Cg(ls, tol=1e-8, maxiter=5))
        s, dt = timed(lambda: kp.linsys.Cg(ls, tol=1e-8, maxiter=200))
        its = len(s.resnorms) - 1
        hist = list(map(float, s.resnorms))
        # explicit residual of the distributed solution: ||M^(1/2)(b - A x)|| / ||M^(1/2) b|| from local parts
        xk = s.__dict__["_xk_dev"]
        MMlr, Mlr, rn = ls._get_residual_dev(xk, compute_norm=True)
        out["c3"] = {"config": "CG + Jacobi 3-D 7-pt Poisson N=%d fp64, row-partitioned over %d GPUs" % (N, world),
                     "iterations": its, "seconds": dt, "it_per_s": its / dt,
                     "algorithmic_GBs_whole_job": 192.0 * N * its / dt / 1e9,
                     "history_identical_on_all_ranks": same_on_all_ranks(hist),
                     "first4_vs_single_gpu_max_rel": first_vs_single(hist, "c3"),
                     "final_resnorm": hist[-1], "explicit_check": float(rn / ls.MMlb_norm),
                     "single_gpu_it_per_s": single.get("c3", {}).get("it_per_s"), "first4": hist[:4]}
        del ls, s, A

    if "c5" in which:
        n = 4000
        N = n * n
        part = kd.RowPartition(N, world, rank)
        A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32, rows=(part.lo, part.hi))
        Bloc = B[:, part.lo:part.hi].tocsr()                 # this rank's diagonal block of B
        b = problems.rhs_normal(N, dtype=np.float32)[part.lo:part.hi]
        res = {}
        for name, dtp in (("fp32", np.float32), ("fp64", np.float64)):
            ls = kd.DistLinearSystem(A, b, part, ip_B=Bloc, self_adjoint=True, dtype=dtp)
            timed(lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=5))
            s, dt = timed(lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=50))
            its = len(s.resnorms) - 1
            sz = 4 if dtp == np.float32 else 8
            nnz = global_sum(float(A.nnz))
            by = (nnz * (sz + 4) + 4 * N + 2 * N * sz) + 2 * (N * (sz + 4) + 4 * N + 2 * N * sz) + 13 * N * sz
            hist = list(map(float, s.resnorms))
            res[name] = {"iterations": its, "seconds": dt, "it_per_s": its / dt,
                         "algorithmic_GBs_whole_job": by * its / dt / 1e9,
                         "history_identical_on_all_ranks": same_on_all_ranks(hist), "hist": hist}
            del ls, s
        a, r = np.array(res["fp32"]["hist"]), np.array(res["fp64"]["hist"])
        m = min(len(a), len(r))
        out["c5"] = {"config": "MINRES(lanczos) ip_B=diag SPD, A=B^-1(L-0.3I) 2-D N=%d, maxiter=50, row-partitioned "
                               "over %d GPUs" % (N, world),
                     "fp32": {k: v for k, v in res["fp32"].items() if k != "hist"},
                     "fp64": {k: v for k, v in res["fp64"].items() if k != "hist"},
                     "fp32_vs_fp64_history_max_rel": float(np.max(np.abs(a[:m] - r[:m]) / r[:m])),
                     "final_resnorm_fp32": float(a[-1]), "final_resnorm_fp64": float(r[-1]),
                     "single_gpu": {k: single.get("c5", {}).get(k, {}).get("it_per_s") for k in ("fp32", "fp64")}}

    if rank == 0:
        sys.stdout.write(json.dumps(out, indent=1) + "\n")
        sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    kd.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
