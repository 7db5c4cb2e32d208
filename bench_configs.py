"""BASELINE.json's other named configurations (C3, C4, C5) at FULL size, shared by bench.py (extra keys
of the bench line, short reference truncations) and tools/run_configs_parity.py (BASELINE.md section 5
truncations, written to profiles/).

  C3  Cg, 3-D 7-point Poisson n=400 (N=64,000,000), Jacobi M = csr diag, fp64          1 and 8 GPUs
  C4  DeflatedGmres, 2-D convection-diffusion n=2000 (N=4,000,000), d=20 Ritz vectors, fp64  1 GPU
  C5  Minres(ortho='lanczos'), ip_B = B SPD diagonal, shifted Laplacian n=4000 (N=16,000,000),
      fp32 storage                                                                     1 and 4 GPUs

For every configuration: iterations/s (CUDA events around the solver call, max over ranks),
algorithmic GB/s by the byte model of SURVEY.md section 8(d), fraction of the measured HBM peak, and
parity of the residual history against the UNMODIFIED reference (baseline/_ref; the oracle port if
that install is missing) run on the same inputs on the host, truncated to a bounded number of steps.
"""
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def bytes_per_iteration(name, N, nnz, sz, iters=None):
    """algorithmic bytes of one iteration (whole job)."""
    spmv = nnz * (sz + 4) + 4 * (N + 1) + 2 * N * sz
    if name == "c3":        # B_spmv + 3Ns (p update) + 8Ns (x, r, z update) = 192 N
        return spmv + 11 * N * sz
    if name == "c5":        # B_spmv(A) + 2 B_spmv(B) + 13 Ns = 144 N (fp32)
        return spmv + 2 * (N * (sz + 4) + 4 * (N + 1) + 2 * N * sz) + 13 * N * sz
    if name == "c2z":       # complex twin of C2, native complex kernels: B_spmv + mean_k (2(k+1)+5) Ns over a
        m = 30              # 30-step cycle + the twin row of v_{k+1} (2 Ns), s = 16
        return spmv + (2 * (m + 1) / 2.0 + 5) * N * sz + 2 * N * sz
    if name == "c4":        # GMRES step k: B_spmv + (2k+9) Ns, mean over the steps done, + 672 N projector
        m = max(int(iters or 1), 1)
        return spmv + (m - 1 + 9) * N * sz + (4 * 20 + 4) * N * sz
    raise ValueError(name)


# ------------------------------------------------------------------------------------------
# problems (host, scipy CSR; rows=(lo, hi): this rank's block with global column indices)
# ------------------------------------------------------------------------------------------
def problem(name, n=None, rows=None):
    from krypy_b200 import problems
    import scipy.sparse as sp
    if name == "c3":
        n = n or 400
        N = n ** 3
        A = problems.poisson3d(n, rows=rows)
        lo, hi = rows if rows else (0, N)
        # Jacobi M = diag(A)^-1 as a CSR diagonal of these rows (constant 1/6 for this stencil)
        M = sp.csr_matrix((np.full(hi - lo, 1.0 / 6.0), np.arange(hi - lo, dtype=np.int32),
                           np.arange(hi - lo + 1, dtype=np.int32)), shape=(hi - lo, hi - lo))
        b = problems.rhs_normal(N)[lo:hi]
        return dict(A=A, b=b, M=M, N=N, nnz_global=7 * N - 6 * n * n, sz=8,
                    ls=dict(M=M, self_adjoint=True, positive_definite=True), solver="cg", kw=dict(tol=1e-8),
                    label="C3: Cg + Jacobi (csr diag), 3-D 7-point Poisson n=%d, N=%d, fp64" % (n, N))
    if name == "c5":
        n = n or 4000
        N = n * n
        A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32, rows=rows)
        lo, hi = rows if rows else (0, N)
        Bloc = B[:, lo:hi].tocsr()
        b = problems.rhs_normal(N, dtype=np.float32)[lo:hi]
        return dict(A=A, b=b, B=Bloc, N=N, nnz_global=5 * N - 4 * n, sz=4,
                    ls=dict(ip_B=Bloc, self_adjoint=True), solver="minres", kw=dict(tol=1e-5),
                    label="C5: Minres(lanczos), ip_B = diag(linspace(1,2,N)), A = B^-1(L - 0.3 I), 2-D n=%d, "
                          "N=%d, fp32 storage" % (n, N))
    if name == "c2z":
        # not a BASELINE configuration: the complex128 twin of C2 (half of the reference's own test matrix is
        # complex, test/test_linsys.py:118-141), 80 MB per vector like C2
        n = n or 2236
        N = n * n
        A = (problems.laplace2d(n).astype(np.complex128) - (0.02 + 0.01j) * sp.identity(N, dtype=np.complex128)).tocsr()
        rng = np.random.default_rng(0)
        b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        return dict(A=A, b=b, N=N, nnz_global=int(A.nnz), sz=16, ls=dict(), solver="gmres", kw=dict(tol=1e-12),
                    label="C2z: complex128 twin of C2, GMRES(30) (ortho=cgs), 2-D 5-point Laplacian - (0.02+0.01i) I, "
                          "n=%d, N=%d, native complex kernels (kry_orth_fused_z, kry_spmv_csr_z)" % (n, N))
    if name == "c4":
        n = n or 2000
        N = n * n
        A = problems.convdiff2d(n, c=0.1, rows=rows)
        lo, hi = rows if rows else (0, N)
        b = np.ones((N, 1))[lo:hi]
        return dict(A=A, b=b, N=N, nnz_global=5 * N - 4 * n, sz=8, ls=dict(), solver="gmres", kw=dict(tol=1e-10),
                    label="C4: DeflatedGmres, 2-D convection-diffusion (cell Peclet 0.1) n=%d, N=%d, d=20 Ritz "
                          "vectors of solve 1, fp64" % (n, N))
    raise ValueError(name)


# ------------------------------------------------------------------------------------------
# host reference (unmodified krypy from baseline/_ref, else the oracle port)
# ------------------------------------------------------------------------------------------
def reference_history(name, P, steps, U=None, krypy=None):
    """residual history of `steps` iterations of the reference on the host: (resnorms, seconds, kind)"""
    import bench
    if krypy is None:
        krypy = bench.load_reference()
    t = time.perf_counter()
    if krypy is not None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if name == "c5":
                ls = krypy.linsys.LinearSystem(P["A"], P["b"], ip_B=P["B"], self_adjoint=True)
                cls, kw = krypy.linsys.Minres, {}
            elif name == "c3":
                ls = krypy.linsys.LinearSystem(P["A"], P["b"], M=P["M"], self_adjoint=True, positive_definite=True)
                cls, kw = krypy.linsys.Cg, {}
            elif name == "c2z":
                ls = krypy.linsys.LinearSystem(P["A"], P["b"])
                cls, kw = krypy.linsys.Gmres, {}
            else:
                ls = krypy.linsys.LinearSystem(P["A"], P["b"])
                cls, kw = krypy.deflation.DeflatedGmres, dict(U=U)
            try:
                sol = cls(ls, maxiter=steps, tol=P["kw"]["tol"], **kw)
            except krypy.utils.ConvergenceError as e:
                sol = e.solver
        return list(map(float, sol.resnorms)), time.perf_counter() - t, "reference"
    from oracle import krylov_oracle as ko
    if name == "c5":
        run, sysm, kw = ko.minres, ko.System(P["A"], P["b"], B=P["B"]), {}
    elif name == "c3":
        run, sysm, kw = ko.cg, ko.System(P["A"], P["b"], M=P["M"]), {}
    elif name == "c2z":
        run, sysm, kw = ko.gmres, ko.System(P["A"], P["b"]), {}
    else:
        run, sysm, kw = ko.gmres, ko.System(P["A"], P["b"]), dict(U=U)
    try:
        sol = run(sysm, maxiter=steps, tol=P["kw"]["tol"], **kw)
    except ko.OracleConvergenceError as e:
        sol = e.result
    return list(map(float, sol.resnorms)), time.perf_counter() - t, "port"


def history_parity(got, ref):
    """max relative difference over the common entries; the LAST entry of a run that ended at maxiter
    is an explicit residual (krypy/linsys.py:450-463) and is reported separately"""
    a, r = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    m = min(len(a), len(r))
    rel = np.abs(a[:m] - r[:m]) / np.maximum(np.abs(r[:m]), 1e-300)
    return {"entries": int(m), "max_rel_updated": float(rel[:m - 1].max()) if m > 1 else 0.0,
            "rel_last_explicit": float(rel[m - 1])}


# ------------------------------------------------------------------------------------------
# device runs
# ------------------------------------------------------------------------------------------
def _timed(torch, dist, fn, kp):
    """device time of fn() in seconds (CUDA events on the current stream, max over ranks)"""
    if dist is not None:
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
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return s, ms * 1e-3


def run_device(name, peak, dist=None, rank=0, world=1, maxiter=None, n=None, ref_steps=0, log=None):
    """One configuration on `world` GPUs (row-partitioned when world > 1).  Returns the result dict
    (rank 0; other ranks return None).  ref_steps > 0: parity of the first ref_steps iterations against
    the host reference, run by rank 0 on the global system."""
    import torch
    import krypy_b200 as kp
    warnings.simplefilter("ignore")
    kd = None
    if world > 1:
        from krypy_b200 import dist as kd
        kd.init()
    say = (lambda *a: None) if (log is None or rank != 0) else log
    t0 = time.time()
    if kd is not None:
        Nglob = problem_size(name, n)
        part = kd.RowPartition(Nglob, world, rank)
        P = problem(name, n, rows=(part.lo, part.hi))
        mk_ls = lambda dtype=None: kd.DistLinearSystem(P["A"], P["b"], part, **dict(P["ls"], **({"dtype": dtype} if dtype else {})))
    else:
        P = problem(name, n)
        mk_ls = lambda dtype=None: kp.linsys.LinearSystem(P["A"], P["b"], **dict(P["ls"], **({"dtype": dtype} if dtype else {})))
    build_s = time.time() - t0
    N, sz = P["N"], P["sz"]
    out = {"config": P["label"], "n_gpus": world, "host_build_s": round(build_s, 1)}
    say("  %s: problem built in %.1f s" % (name, build_s))

    if name == "c3":
        maxiter = maxiter or 200
        ls = mk_ls()
        _timed(torch, dist, lambda: kp.linsys.Cg(ls, maxiter=5, **P["kw"]), kp)          # warm-up (no basis: short)
        s, dt = _timed(torch, dist, lambda: kp.linsys.Cg(ls, maxiter=maxiter, **P["kw"]), kp)
        hist = list(map(float, s.resnorms))
        trunc = lambda k: kp.linsys.Cg(ls, maxiter=k, **P["kw"])
    elif name == "c2z":
        if world > 1:
            raise NotImplementedError("complex row-partitioned runs are not implemented")
        maxiter = maxiter or 30
        ls = mk_ls()
        mk = lambda restarts: kp.linsys.RestartedGmres(ls, maxiter=maxiter, max_restarts=restarts, ortho="cgs", **P["kw"])
        _timed(torch, dist, lambda: mk(0), kp)                                               # warm-up: one cycle
        s, dt = _timed(torch, dist, lambda: mk(4), kp)                                       # five cycles
        hist = list(map(float, s.resnorms))
        trunc = lambda k: kp.linsys.Gmres(ls, maxiter=k, ortho="cgs", **P["kw"])
    elif name == "c5":
        maxiter = maxiter or 50
        ls = mk_ls(np.float32)
        _timed(torch, dist, lambda: kp.linsys.Minres(ls, maxiter=maxiter, **P["kw"]), kp)   # warm-up: same basis size
        s, dt = _timed(torch, dist, lambda: kp.linsys.Minres(ls, maxiter=maxiter, **P["kw"]), kp)
        hist = list(map(float, s.resnorms))
        trunc = lambda k: kp.linsys.Minres(ls, maxiter=k, **P["kw"])
    else:
        # C4 work-flow of SURVEY 8(d) on the device: solve 1 (no deflation, store_arnoldi) -> 20 Ritz
        # vectors of smallest magnitude (kept in HBM) -> timed solve 2 with U
        if world > 1:
            raise NotImplementedError("C4 is a single-GPU configuration")
        maxiter = maxiter or 60
        ls = mk_ls()
        fac = kp.recycling.factories.RitzFactorySimple(n_vectors=20, which="sm")
        s1, dt1 = _timed(torch, dist, lambda: kp.deflation.DeflatedGmres(ls, maxiter=maxiter, store_arnoldi=True,
                                                                         ortho="cgs", **P["kw"]), kp)
        torch.cuda.synchronize()
        tf = time.perf_counter()
        Ublk = fac.get(s1)
        torch.cuda.synchronize()
        tfac = time.perf_counter() - tf
        run2 = lambda: kp.deflation.DeflatedGmres(ls, U=Ublk, maxiter=maxiter, ortho="cgs", **P["kw"])
        _timed(torch, dist, run2, kp)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        proj = kp.deflation.ObliqueProjection(ls, Ublk)
        torch.cuda.synchronize()
        tset = time.perf_counter() - tp
        del proj
        s, dt = _timed(torch, dist, run2, kp)
        hist = list(map(float, s.resnorms))
        run0 = lambda: kp.linsys.Gmres(ls, maxiter=maxiter, ortho="cgs", **P["kw"])
        _timed(torch, dist, run0, kp)
        s0, dt0 = _timed(torch, dist, run0, kp)
        out.update({"solve1_it_per_s": (len(s1.resnorms) - 1) / dt1, "ritz_vectors_s": tfac,
                    "projector_setup_s": tset, "undeflated_it_per_s": (len(s0.resnorms) - 1) / dt0,
                    "undeflated_final_resnorm": float(s0.resnorms[-1]),
                    "deflated_step_over_undeflated": ((dt - tset) / max(len(hist) - 1, 1)) / (dt0 / max(len(s0.resnorms) - 1, 1)),
                    "U": "20 Ritz vectors ('sm') of solve 1 (this engine), resident in HBM (utils.DeviceBlock)"})
        trunc = None
    its = len(hist) - 1
    by = bytes_per_iteration(name, N, P["nnz_global"], sz, its)
    gbs = by * its / dt / 1e9
    out.update({"iterations": its, "seconds": dt, "it_per_s": its / dt, "algorithmic_bytes_per_iteration": by,
                "algorithmic_GBs": gbs, "frac_of_measured_peak": gbs / (peak * world),
                "frac_of_8TBs_nominal": gbs / (8000.0 * world), "first4": hist[:4], "final_resnorm": hist[-1]})
    say("  %s: %d iterations, %.1f it/s, %.0f GB/s algorithmic (%.2f of measured peak x %d)"
        % (name, its, its / dt, gbs, out["frac_of_measured_peak"], world))
    if dist is not None:
        h = torch.tensor(hist, device="cuda", dtype=torch.float64)
        lo, hi = h.clone(), h.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["history_identical_on_all_ranks"] = bool(torch.equal(lo, hi))

    # ---- parity against the host reference on a bounded number of steps ----
    if ref_steps and trunc is not None:
        try:
            g = trunc(ref_steps)
        except kp.utils.ConvergenceError as e:
            g = e.solver
        got = list(map(float, g.resnorms))
        if rank == 0:
            Pg = P if world == 1 else problem(name, n)
            ref, tref, kind = reference_history(name, Pg, ref_steps)
            par = history_parity(got, ref)
            par.update({"reference_kind": kind, "reference_seconds": tref, "reference_it_per_s": (len(ref) - 1) / tref,
                        "steps": ref_steps, "tolerance": "1e-10 relative (fp64)" if sz == 8 else "1e-4 relative (fp32 storage)"})
            out["parity_vs_reference"] = par
            out["speedup_vs_reference_same_box"] = out["it_per_s"] / par["reference_it_per_s"]
            say("  %s: parity vs %s over %d steps: max rel %.2e (last, explicit: %.2e); reference %.2f it/s"
                % (name, kind, ref_steps, par["max_rel_updated"], par["rel_last_explicit"], par["reference_it_per_s"]))
        if dist is not None:
            dist.barrier()
    return out if rank == 0 else None


def problem_size(name, n=None):
    if name == "c2z":
        return (n or 2236) ** 2
    if name == "c3":
        return (n or 400) ** 3
    if name == "c5":
        return (n or 4000) ** 2
    return (n or 2000) ** 2
