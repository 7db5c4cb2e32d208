#!/usr/bin/env python
"""bench.py -- GMRES iterations/s and HBM GB/s on BASELINE.json's headline config.

Workload (config.workload): C2 = GMRES(30) on the 2-D 5-point Laplacian in CSR,
n=3162 (N=9,998,244, nnz=49,978,572), fp64, b = default_rng(0).standard_normal.
A "step" is ONE GMRES(30) restart cycle (30 Arnoldi iterations + the cycle's initial and
final explicit residual and the solution update), driven through the public
krypy-compatible API (krypy_b200.linsys.Gmres, x0 = previous cycle's x) exactly as
RestartedGmres does.

  value : iterations/s with A, b resident in HBM (CUDA events around K cycles, max over ranks)
  e2e   : same metric through the public API from HOST (pinned) buffers: every step uploads
          A (rowptr, colidx, vals), b and x0 and downloads x_k inside the timed region
  roofline : dominant kernel = the fused Gram-Schmidt kernel (kry_orth_fused), algorithmic
          bytes per launch (2(k+1)+5)*N*8 / CUDA-event time per launch, live
  cpu_baseline : the oracle (numpy/scipy port of the reference's algorithm) on the host cores

  --impl reference : times the oracle port alone (rank 0 only) on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 3162
RESTART = 30
TOL = 1e-12
METRIC = "gmres_iterations_per_second"
UNIT = "iterations/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=N_GRID, help="grid size n (N = n*n); default = config C2")
    p.add_argument("--ortho", default="cgs", help="cgs (fused block, default) | mgs | dmgs | cgs2")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extra-configs", action="store_true",
                   help="skip the other named configurations (C3/C4/C5 at N=1, C5 at N=4, C3 at N=8)")
    p.add_argument("--extra", default=None, help="comma list overriding the extra configurations, e.g. c3,c5")
    p.add_argument("--no-mgs", action="store_true", help="skip the drop-in default (ortho='mgs') figure")
    p.add_argument("--cpu-iters", type=int, default=4, help="iterations of the bounded CPU sample (cpu_baseline)")
    p.add_argument("--ref-iters", type=int, default=3, help="--impl reference: Arnoldi iterations per step")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------
class Clocks(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.2 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            for ts, line in self.lines[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# CPU reference arm: the UNMODIFIED reference (pip-installed into baseline/_ref, which travels to the
# GPU box) through its own public API; the oracle port only when that install is missing
# ---------------------------------------------------------------------------------------
class _StopReference(Exception):
    """raised from the operator callback to end a reference run after a bounded number of steps"""


def load_reference():
    """the reference package (numpy>=2 / scipy>=1.12 name aliases from oracle/refshim.py; the
    reference's files are untouched), or None"""
    try:
        from oracle import refshim
        if refshim.vendored_available():
            return refshim.import_reference(refshim.VENDORED_ROOT)
    except Exception as exc:           # noqa: BLE001
        sys.stderr.write("reference import failed: %r\n" % (exc,))
    return None


def reference_cycle(krypy, A, b, iters, maxiter=RESTART):
    """`iters` Arnoldi iterations of ONE krypy.linsys.Gmres(maxiter=30) cycle of the unmodified
    reference at full N -- same allocation ((N, 31) basis), same code path as the whole cycle.
    The reference's constructor cannot be stopped, so the run is ended by an exception raised from
    the operator callback when the (iters+1)-th application of A is requested; the residual
    history is read from the solver object found in the traceback.  Returns (iterations, seconds,
    resnorms)."""
    calls = [0]

    def dot(X):
        calls[0] += 1
        if calls[0] > iters:
            raise _StopReference()
        return A.dot(X)

    op = krypy.utils.LinearOperator(A.shape, A.dtype, dot=dot)
    ls = krypy.linsys.LinearSystem(op, b)
    resn = None
    t = time.perf_counter()
    try:
        sol = krypy.linsys.Gmres(ls, maxiter=maxiter, tol=TOL)
        resn = list(map(float, sol.resnorms))
    except _StopReference as e:
        dt = time.perf_counter() - t
        tb = e.__traceback__
        while tb is not None:
            obj = tb.tb_frame.f_locals.get("self")
            if obj is not None and hasattr(obj, "resnorms") and hasattr(obj, "linear_system"):
                resn = list(map(float, obj.resnorms))
            tb = tb.tb_next
    except krypy.utils.ConvergenceError as e:
        resn = list(map(float, e.solver.resnorms))
    dt = time.perf_counter() - t
    return len(resn) - 1, dt, resn


def cpu_cycle(A, b, iters):
    """one truncated GMRES cycle of the oracle PORT at full N: returns (iterations, seconds, resnorms)"""
    from oracle import krylov_oracle as ko
    t = time.perf_counter()
    try:
        r = ko.gmres(ko.System(A, b), maxiter=iters, tol=TOL)
    except ko.OracleConvergenceError as e:
        r = e.result
    dt = time.perf_counter() - t
    return len(r.resnorms) - 1, dt, list(map(float, r.resnorms))


def cpu_sample(A, b, iters):
    """(iterations, seconds, resnorms, kind, description) of the bounded CPU sample"""
    krypy = load_reference()
    N = A.shape[0]
    if krypy is not None:
        it, dt, rn = reference_cycle(krypy, A, b, iters)
        return it, dt, rn, "reference", (
            "UNMODIFIED reference (krypy %s, pip-installed into baseline/_ref) through its public API: "
            "krypy.linsys.Gmres(LinearSystem(A, b), maxiter=30, tol=1e-12) on the full N=%d system, ended after "
            "the first %d of the 30 Arnoldi iterations of the cycle by an exception from the operator callback "
            "(same (N,31) basis allocation and code path as the full cycle; early iterations orthogonalise "
            "against fewer vectors than the cycle average, which favours the CPU; one full cycle of the same "
            "call is recorded in profiles/r2_reference_full_cycle.json)" % (getattr(krypy, "__version__", "?"), N, it))
    it, dt, rn = cpu_cycle(A, b, iters)
    return it, dt, rn, "port", (
        "oracle/krylov_oracle.py (numpy/scipy port; baseline/_ref missing) on the full N=%d system: the first %d "
        "Arnoldi iterations of a GMRES cycle" % (N, it))


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from krypy_b200 import problems
    n = args.n
    A = problems.laplace2d(n)
    b = problems.rhs_normal(n * n)
    iters = args.ref_iters   # bounded sample per step: the first Arnoldi steps of a GMRES(30) cycle
    kind = sample = None
    for _ in range(args.warmup):
        cpu_sample(A, b, iters)
    t_tot, it_tot = 0.0, 0
    for _ in range(args.steps):
        it, dt, _, kind, sample = cpu_sample(A, b, iters)
        t_tot += dt
        it_tot += it
    val = it_tot / t_tot
    cores = cpu_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, world),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "host_cpus": os.cpu_count(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    sys.stdout.write("\n" + json.dumps(line) + "\n")
    sys.stdout.flush()


def config_dict(args, world):
    n = args.n
    return {"workload": "C2: GMRES(30) restart cycles, 2-D 5-point Laplacian CSR, n=%d, N=%d, fp64, "
                        "b=default_rng(0).standard_normal; one step = one 30-iteration cycle" % (n, n * n),
            "N": n * n, "restart": RESTART, "tol": TOL, "ortho": args.ortho,
            "partition": "single GPU" if world == 1 else "row-partitioned over %d GPUs" % world,
            "l2": "inputs larger than L2 (basis 2.5 GB, A 0.64 GB vs 126 MB L2): no flush needed",
            "pacing": ("restart cycles as linsys.RestartedGmres runs them: from the third cycle over one workspace on, "
                       "all 30 steps of a cycle are enqueued ahead of the host (%s), one host synchronisation per "
                       "cycle, and the next cycle is launched on the device-side residual norm before the host has "
                       "read it (every timed step still waits for its own cycle's records, explicit residual and "
                       "x_k; the cycle launched at the end of the last timed step is inside the timed region, the "
                       "one the first timed step starts from was launched by the last warm-up step)"
                       % ("eager launches with live per-kernel CUDA events" if world == 1 else
                          "one CUDA graph per cycle; one cross-GPU wait per Arnoldi step"))}


# ---------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------
def orth_bytes(nq, nv, passes, algo, has_next):
    """algorithmic bytes of one Gram-Schmidt launch (fp64).  kry_orth_fused: Vdot once + Vsub once per pass (nv
    vectors each), q read twice + written once per pass, + the normalised store (read q, write v_next) --
    SURVEY 8d "[2(k+1)+3]Ns + 2Ns".  algo 100 = kry_dist_update_scale (row-partitioned fused step): V once, q
    read once, v_next written once; algo 101 = kry_dist_dot with want_sq, its first kernel."""
    if algo == 101:          # kry_dist_dot with <q,q>: V once, q once
        return (nv + 1) * nq * 8.0
    if algo == 100:
        return (nv + 2) * nq * 8.0
    return (passes * (2 * nv + 3) + (2 if has_next else 0)) * nq * 8.0


def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import warnings
    torch.cuda.set_device(local_rank)
    import krypy_b200 as kp
    from krypy_b200 import _device, problems

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    N = n * n
    ctx = _device.Context.get()
    warnings.simplefilter("ignore")

    # ---- problem (host) ---------------------------------------------------------------
    if world == 1:
        A = problems.laplace2d(n)
        b = problems.rhs_normal(N)
        ls = kp.linsys.LinearSystem(A, b)
        ws = kp.utils.SolverWorkspace()
        make_solver = lambda x0, r0=None: kp.linsys.Gmres(ls, x0=x0, maxiter=RESTART, tol=TOL, ortho=args.ortho,
                                                          _workspace=ws, _x0_residual=r0, _prelaunch=True)
    else:
        from krypy_b200 import dist as kdist
        part = kdist.RowPartition(N, world, rank)
        A = problems.laplace2d(n, rows=(part.lo, part.hi))
        b = problems.rhs_normal(N)[part.lo:part.hi]
        ls = kdist.DistLinearSystem(A, b, part)
        ws = kp.utils.SolverWorkspace()      # persistent buffers + one CUDA graph per Arnoldi step
        make_solver = lambda x0, r0=None: kp.linsys.Gmres(ls, x0=x0, maxiter=RESTART, tol=TOL, ortho=args.ortho,
                                                          _workspace=ws, _x0_residual=r0, _prelaunch=True)

    carry = {"r": None}

    def cycle(x0):
        # exactly what RestartedGmres does per restart (x0 = previous x_k, and the explicit residual the
        # previous cycle computed for it is handed over instead of being recomputed)
        try:
            sol = make_solver(x0, carry["r"] if x0 is not None else None)
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        carry["r"] = sol.__dict__.get("_last_residual")
        return sol

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----------------------------------------------------------------------
    x = None
    for _ in range(max(args.warmup, 1)):
        sol = cycle(x)
        x = sol.__dict__["_xk_dev"].reshape(-1)

    # ---- timed region: value (device-resident inputs) ------------------------------------
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    # N=1: per-kernel CUDA-event brackets live inside the timed region (eager launches).
    # N>1: the timed region replays one CUDA graph per Arnoldi step (no per-kernel events can be
    # recorded inside a graph); the brackets are taken in an extra eager cycle right after it.
    timer = _device.KernelTimer()
    ctx.timer = timer if world == 1 else None
    ctx.reset_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    iters = 0
    hist = []
    for _ in range(args.steps):
        sol = cycle(x)
        x = sol.__dict__["_xk_dev"].reshape(-1)
        iters += len(sol.resnorms) - 1
        hist.append(sol.resnorms)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    ctx.timer = None
    launches = ctx.launch_count()
    l2w = getattr(ctx, "_l2win", None)       # the L2 residency window of w = A v_k, if the product set one (KRY_L2_WINDOW)
    l2_info = ({"enabled": True, "set_aside_bytes": l2w[1][2], "window_bytes": l2w[1][3], "hit_ratio": l2w[1][4],
                "note": "w = A v_k is kept in the L2 set-aside (kry_l2_window): its re-reads are served by L2, so the "
                        "algorithmic GB/s of the Gram-Schmidt kernel can exceed the DRAM copy peak (roofline.frac > 1); "
                        "roofline.traffic was captured with the window off (profiles/r2_traffic.json); same-box A/B: "
                        "profiles/r2_l2window_ab.json"}
               if (l2w and l2w[1]) else {"enabled": False, "note": getattr(ctx, "l2_window_error", None)})
    if world > 1:
        # graph replays are not counted by the library's launch counter: count the kernels of one
        # eagerly launched cycle instead (same sequence the graphs replay) and scale
        ctx.reset_launch_count()
        ctx.timer = timer
        sol = cycle(x)
        torch.cuda.synchronize()
        ctx.timer = None
        launches = ctx.launch_count() * args.steps
    if dist is not None:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clk = clocks.stop(t0, t1) if rank == 0 else None
    value = iters / (ms * 1e-3)

    # ---- per-kernel roofline from the live CUDA-event brackets ------------------------------
    peak, peak_src = peaks()
    summ = timer.summary()
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")     # ncu dram__bytes per launch (committed)
    if os.path.exists(tpath) and n == N_GRID and world == 1:
        with open(tpath) as f:
            traffic = json.load(f)
    roof = None
    extra = {}
    nloc = N if world == 1 else (part.hi - part.lo)
    if "orth" in summ:
        o = summ["orth"]
        # algorithmic bytes per launch: Vdot once + Vsub once per pass (nv vectors each), q read twice +
        # written once per pass, + phase C (read q, write v_next): SURVEY 8d "[2(k+1)+3]Ns + 2Ns"
        by = 0.0
        for (nq, nv, passes, algo, has_next) in o["meta"]:
            by += orth_bytes(nq, nv, passes, algo, has_next)
        fused2 = any(m[3] == 100 for m in o["meta"])
        ach = by / (o["ms_total"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": ("dist_dot_kernel + dist_update_scale_kernel (kry_dist_dot with <w,w>, "
                                           "kry_dist_update_scale: one cross-GPU wait per Arnoldi step)" if fused2 else
                                           "orth_kernel<double,2> (kry_orth_fused, ortho=%s)" % args.ortho),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get("orth", {}).get("traffic_per_launch"),
                "traffic_source": "profiles/r2_traffic.json (ncu dram__bytes_read.sum+dram__bytes_write.sum, "
                                  "mean over the 30 launches of one GMRES(30) cycle)" if traffic else None,
                "peak_source": peak_src, "launches": o["launches"],
                "avg_launch_ms": o["ms_total"] / max(o["launches"], 1),
                "algorithmic_bytes_per_launch_avg": by / max(o["launches"], 1),
                "share_of_step": (o["ms_total"] / ms) if world == 1 else None,
                "timing": "CUDA events in the timed region" if world == 1 else
                          "CUDA events in one eager cycle after the timed region (timed region replays CUDA "
                          "graphs); " + ("SpMV + two Gram-Schmidt kernels per Arnoldi step, one NVLink peer all-reduce" if fused2 else
                                         "split kernels + NVLink peer all-reduces") + ", per-rank bytes"}
        if l2_info.get("enabled"):
            # w = A v_k is resident in the L2 set-aside: part of the ALGORITHMIC bytes never reaches HBM, so the
            # fraction against the DRAM copy peak can exceed 1; `traffic` is the ncu capture with the window off
            roof["l2_resident_vector"] = ("w = A v_k (%d of the mean %.0f vector passes per launch are re-reads / the "
                                          "rewrite of w that the L2 set-aside can serve); frac > 1 is possible"
                                          % (3, by / max(o["launches"], 1) / (8.0 * N / world)))
    if "orth" in summ:
        # per-k profile of the fused Gram-Schmidt kernel: average microseconds and algorithmic GB/s by
        # the number of basis vectors involved (shows the small-k inefficiency; A/B of KRY_ORTH_SMALLK)
        try:
            by_nv = {}
            held = None          # (the one-wait step brackets its two Gram-Schmidt kernels separately: one entry per step)
            for t_ms, (nq, nv, passes, algo, has_next) in zip(summ["orth"]["ms"], summ["orth"]["meta"]):
                ent = (t_ms, orth_bytes(nq, nv, passes, algo, has_next))
                if algo == 101:
                    held = ent
                    continue
                if algo == 100 and held is not None:
                    ent = (ent[0] + held[0], ent[1] + held[1])
                    held = None
                by_nv.setdefault(int(nv), []).append(ent)
            extra["orth_by_nv"] = {str(nv): {"us": round(1e3 * sum(t for t, _ in v) / len(v), 1),
                                             "GBs": round(sum(b for _, b in v) / sum(t for t, _ in v) / 1e6, 0)}
                                   for nv, v in sorted(by_nv.items())}
        except Exception as exc:
            extra["orth_by_nv"] = {"error": repr(exc)}
    if "spmv" in summ:
        s = summ["spmv"]
        # (a third entry: vectors of the multi-dot epilogue, read once each)
        by = sum(m[1] * 12.0 + 4.0 * (m[0] + 1) + 2.0 * m[0] * 8.0 + (m[2] * m[0] * 8.0 if len(m) > 2 else 0.0)
                 for m in s["meta"])
        ach = by / (s["ms_total"] * 1e-3) / 1e9
        extra["roofline_spmv"] = {"bound": "hbm", "kernel": "spmv_staged_kernel<double,8> (kry_spmv_csr)",
                                  "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                  "traffic": traffic.get("spmv", {}).get("traffic_per_launch"),
                                  "algorithmic_bytes_per_launch": by / max(s["launches"], 1),
                                  "launches": s["launches"], "avg_launch_ms": s["ms_total"] / max(s["launches"], 1),
                                  "share_of_step": (s["ms_total"] / ms) if world == 1 else None}
        if roof is not None:
            byo = roof["algorithmic_bytes_per_launch_avg"] * roof["launches"]
            pair = (by + byo) / ((s["ms_total"] + summ["orth"]["ms_total"]) * 1e-3) / 1e9
            extra["roofline_spmv_plus_orth"] = {"achieved": pair, "unit": "GB/s", "frac_of_measured_peak": pair / peak,
                                                "frac_of_8TBs_nominal": pair / 8000.0}
    # whole-iteration algorithmic bytes (383.5*N per iteration at m=30) -> GB/s of the full step
    extra["whole_step_algorithmic_gbs"] = 383.5 * N * value / 1e9

    # ---- e2e: public API from host buffers, copies inside the timed region --------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, kp, torch, A, b, x, dist=dist, part=part if world > 1 else None, ls0=ls)

    # ---- the drop-in default: the same cycle with ortho='mgs' (the reference's exact operation order) ----
    if not args.no_mgs and args.ortho != "mgs":
        try:
            wsm = kp.utils.SolverWorkspace()
            mcarry = {"r": None}
            def mcycle(x0):
                try:
                    sm = kp.linsys.Gmres(ls, x0=x0, maxiter=RESTART, tol=TOL, _workspace=wsm,      # default ortho
                                         _x0_residual=mcarry["r"] if x0 is not None else None)
                except kp.utils.ConvergenceError as e:
                    sm = e.solver
                mcarry["r"] = sm.__dict__.get("_last_residual")
                return sm
            xm = mcycle(None).__dict__["_xk_dev"].reshape(-1)
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            nm, itm = 3, 0
            for _ in range(nm):
                sm_ = mcycle(xm)
                xm = sm_.__dict__["_xk_dev"].reshape(-1)
                itm += len(sm_.resnorms) - 1
            m1.record()
            barrier()
            mms = m0.elapsed_time(m1)
            if dist is not None:
                tm_ = torch.tensor([mms], device="cuda", dtype=torch.float64)
                dist.all_reduce(tm_, op=dist.ReduceOp.MAX)
                mms = float(tm_.item())
            extra["mgs_value"] = {"value": itm / (mms * 1e-3), "unit": UNIT, "steps": nm, "ms_per_step": mms / nm,
                                  "what": "the same C2 cycle with the drop-in default ortho='mgs' (exact modified "
                                          "Gram-Schmidt: k+1 dependent reductions per step, (4(k+1)+4) N s bytes)"}
            del wsm, sm_, xm
        except Exception as exc:       # noqa: BLE001
            extra["mgs_value"] = {"error": repr(exc)}

    # ---- CPU baseline (rank 0): bounded sample of the same workload, and parity against it ----------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            first = kp.linsys.Gmres(ls, maxiter=args.cpu_iters, tol=TOL, ortho=args.ortho)
        except kp.utils.ConvergenceError as e:
            first = e.solver
        if rank == 0:
            Ag, bg = (A, b) if world == 1 else (problems.laplace2d(n), problems.rhs_normal(N))
            it, dt, rn, kind, sample = cpu_sample(Ag, bg, args.cpu_iters)
            cpu = {"value": it / dt, "unit": UNIT, "cores": cpu_threads(), "host_cpus": os.cpu_count(), "kind": kind,
                   "sample": sample + " [%.1f s]" % dt}
            # parity of the first cycle's history against the CPU run, same inputs (x0 = None)
            a, r = np.array(first.resnorms), np.array(rn)
            m = min(len(a), len(r))
            extra["parity_vs_cpu_max_rel"] = float(np.max(np.abs(a[:m] - r[:m]) / r[:m]))
            del Ag, bg
        if dist is not None:
            dist.barrier()

    # ---- the other named configurations (BASELINE.json configs[2..4]) at full size ------------------------
    if not args.no_extra_configs and n == N_GRID:
        import bench_configs
        del sol
        # (c2z: the complex128 twin of C2 on the native complex kernels -- not a BASELINE configuration)
        which = {1: [("c5", 6), ("c4", 0), ("c3", 4), ("c2z", 3)], 4: [("c5", 6)], 8: [("c3", 4)]}.get(world, [])
        if getattr(args, "extra", None):
            which = [(c, {"c5": 6, "c3": 4, "c2z": 3}.get(c, 0)) for c in args.extra.split(",") if c]
        ctx.l2_window(None)                  # C2's window (if any) goes with C2's buffers
        if which:
            # release C2's device memory first
            ws.bufs.clear(); ws.graphs.clear()
            carry["r"] = None
            x = None
            torch.cuda.empty_cache()
        cfgs = {}
        for cname, ref_steps in which:
            try:
                cfgs[cname] = bench_configs.run_device(cname, peak, dist=dist, rank=rank, world=world,
                                                       ref_steps=0 if args.no_cpu_baseline else ref_steps,
                                                       log=lambda *a: sys.stderr.write(" ".join(map(str, a)) + "\n"))
            except Exception as exc:       # noqa: BLE001  (never lose the bench line over an extra)
                cfgs[cname] = {"error": repr(exc)}
            torch.cuda.empty_cache()
        if which:
            extra["configs"] = cfgs

    if rank == 0 and os.environ.get("KRY_TRACE"):
        tr = []
        sys.stderr.write("phase trace (ms since start): " + ", ".join(
            "%s=%.3f" % (nm, 1e3 * (tt - tr[0][1])) for nm, tt in tr) + "\n")
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, world), "iterations": iters, "clocks": clk,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "final_resnorm": float(hist[-1][-1]), "l2_window": l2_info,
        }
        line.update(extra)
        sys.stdout.write("\n" + json.dumps(line) + "\n")
        sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(args, kp, torch, A, b, x_dev, dist=None, part=None, ls0=None):
    """Every step: host (pinned) A, b, x0 -> LinearSystem -> one Gmres(30) cycle -> x_k on the host.
    Row-partitioned runs: every rank uploads ITS row block of A, b, x0 and reads back its segment of
    x_k; the halo plan (host-side symbolic analysis of the sparsity pattern, computed once when the
    first operator over this matrix was built) is reused, the numeric arrays are uploaded every step."""
    import numpy as np
    import scipy.sparse as sp

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    keep = []
    parts = []
    for arr in (A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32)):
        v, t = pin(arr)
        keep.append(t)
        parts.append(v)
    Ah = sp.csr_matrix((parts[0], parts[1], parts[2]), shape=A.shape)
    Ah.has_sorted_indices = True
    bh, tb = pin(b)
    keep.append(tb)
    xh, tx = pin(x_dev.cpu().numpy())
    keep.append(tx)
    h2d = parts[0].nbytes + parts[1].nbytes + parts[2].nbytes + bh.nbytes + xh.nbytes
    d2h = xh.nbytes
    world = 1
    if part is not None:
        from krypy_b200 import dist as kdist
        world = part.world
        plan = ls0.A.plan

    def make_ls():
        if part is None:
            return kp.linsys.LinearSystem(Ah, bh)
        return kdist.DistLinearSystem(kdist.DistCsrOperator(Ah, part, plan=plan), bh, part)

    def step(x0h):
        ls = make_ls()
        try:
            sol = kp.linsys.Gmres(ls, x0=x0h, maxiter=RESTART, tol=TOL, ortho=args.ortho)
        except kp.utils.ConvergenceError as e:
            sol = e.solver
        return sol.xk, len(sol.resnorms) - 1       # .xk: device -> host read of the step's result

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    xk, _ = step(xh)                                 # warm-up
    xk, _ = step(xk)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = max(2, min(args.steps, 5))
    e0.record()
    its = 0
    wall = []
    for _ in range(steps):
        tw = time.perf_counter()
        xk, it = step(xk)                 # (ends with the synchronous read-back of x_k)
        wall.append(round(1e3 * (time.perf_counter() - tw), 2))
        its += it
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tms = torch.tensor([ms, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        mx = tms.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tms, op=dist.ReduceOp.SUM)
        ms, h2d, d2h = float(mx[0].item()), float(tms[1].item()), float(tms[2].item())
    out = {"value": its / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": ms / steps, "step_ms_wall": wall,
           "api": ("krypy_b200.linsys.LinearSystem(A_host, b_host) + Gmres(x0=x_host, maxiter=30) per step"
                   if part is None else
                   "per rank: krypy_b200.dist.DistLinearSystem(DistCsrOperator(A_rows_host, part, plan), b_rows_host) + "
                   "Gmres(x0=x_rows_host, maxiter=30) per step; bytes summed over the %d ranks" % world)}
    if part is not None:
        return out
    # the same from the user's actual call: ONE RestartedGmres(30) solve of 5 cycles per upload of A
    # (informational; the strict per-cycle-upload number above stays the e2e figure)
    try:
        def solve(x0h):
            ls = kp.linsys.LinearSystem(Ah, bh)
            try:
                sol = kp.linsys.RestartedGmres(ls, x0=x0h, maxiter=RESTART, max_restarts=4, tol=TOL, ortho=args.ortho)
            except kp.utils.ConvergenceError as e:
                sol = e.solver
            return sol.xk, len(sol.resnorms) - 1
        solve(xh)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        its2 = 0
        for _ in range(2):
            xk2, it = solve(xh)
            its2 += it
        f1.record()
        torch.cuda.synchronize()
        ms2 = f0.elapsed_time(f1)
        out["per_solve"] = {"value": its2 / (ms2 * 1e-3), "unit": UNIT, "iterations_per_upload": its2 // 2,
                            "ms_per_solve": ms2 / 2, "h2d_bytes_per_solve": int(h2d), "d2h_bytes_per_solve": int(d2h),
                            "api": "krypy_b200.linsys.RestartedGmres(LinearSystem(A_host, b_host), x0=x_host, "
                                   "maxiter=30, max_restarts=4)"}
    except Exception as exc:
        out["per_solve"] = {"error": repr(exc)}
    # untimed diagnostic pass AFTER the measurement: where one e2e step spends its wall time
    # (synchronised phase marks; explains the gap between `e2e` and `value`)
    try:
        out["breakdown_ms"] = e2e_breakdown(kp, torch, Ah, bh, xk, args, keep)
    except Exception as exc:       # diagnostics must never cost the bench line
        out["breakdown_ms"] = {"error": repr(exc)}
    return out


def e2e_breakdown(kp, torch, Ah, bh, xh, args, pinned):
    """One extra e2e step with a device synchronise + wall-clock mark after each phase."""
    def mark():
        torch.cuda.synchronize()
        return time.perf_counter()

    res = {}
    # raw pinned-copy bandwidth of this box (the CSR values array, 0.4 GB), for scale
    vals = pinned[0]
    dst = torch.empty(vals.shape, dtype=vals.dtype, device="cuda")
    t0 = mark()
    dst.copy_(vals, non_blocking=True)
    t1 = mark()
    back = torch.empty(vals.shape, dtype=vals.dtype).pin_memory()
    t2 = mark()
    back.copy_(dst, non_blocking=True)
    t3 = mark()
    res["pinned_h2d_GBps"] = vals.numel() * vals.element_size() / (t1 - t0) / 1e9
    res["pinned_d2h_GBps"] = vals.numel() * vals.element_size() / (t3 - t2) / 1e9
    del dst, back
    t0 = mark()
    ls = kp.linsys.LinearSystem(Ah, bh)
    t1 = mark()
    ls.A._dev(ls._td)                                    # H2D of rowptr/colidx/vals (otherwise lazy)
    t2 = mark()
    try:
        sol = kp.linsys.Gmres(ls, x0=xh, maxiter=RESTART, tol=TOL, ortho=args.ortho)
    except kp.utils.ConvergenceError as e:
        sol = e.solver
    t3 = mark()
    xk = sol.xk
    t4 = mark()
    res.update({"linear_system_init": 1e3 * (t1 - t0), "upload_A": 1e3 * (t2 - t1),
                "solver_cycle_incl_x0_upload": 1e3 * (t3 - t2), "xk_readback": 1e3 * (t4 - t3),
                "total": 1e3 * (t4 - t0)})
    assert xk is not None
    return res


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep NCCL's version banner off stdout (one JSON line only)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
