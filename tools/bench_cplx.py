"""Complex128 twin of config C2 on one B200: restarted GMRES(30) on a shifted 2-D 5-point Laplacian with a complex
shift (Helmholtz-like), N = n^2 complex unknowns (n = 2236: 80 MB per vector, the size of C2's real vectors),
through the public API, with the NATIVE complex kernels (kry_orth_fused_z, kry_spmv_csr_z; default) and with the
real-embedding kernels (KRY_NATIVE_Z=0: twin storage, embedded CSR) on the same inputs.

Reports per path and Gram-Schmidt variant: iterations/s, the two residual histories' agreement, and -- from CUDA
events around every launch (a second, eagerly enqueued run) -- the average duration and algorithmic GB/s of the
SpMV and of the Gram-Schmidt kernel against the measured copy peak.  Byte model per iteration (native, what a
complex kernel needs): SpMV 20 nnz + 32 N; fused block Gram-Schmidt at step k (2(k+1)+5) 16 N; the twin row of
v_{k+1} that the other consumers of the basis still read (kry_rot90) 32 N.

usage: python tools/bench_cplx.py [n] > gpurun_out/cplx.json"""
import json
import os
import sys
import time
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
import torch

import krypy_b200 as kp
from krypy_b200 import _device, problems, utils

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
N = n * n
RESTART, CYCLES = 30, 5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"]
    peak_src = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    PEAK, peak_src = 6500.0, "fallback"
ctx = _device.Context.get()
t0 = time.time()
A = (problems.laplace2d(n).astype(np.complex128) - (0.02 + 0.01j) * sp.identity(N, dtype=np.complex128)).tocsr()
rng = np.random.default_rng(0)
b = rng.standard_normal(N) + 1j * rng.standard_normal(N)
out = {"config": "complex twin of C2: GMRES(%d), 2-D 5-point Laplacian - (0.02+0.01i) I, n=%d, N=%d complex128, "
                 "b = standard normal (re, im), %d cycles" % (RESTART, n, N, CYCLES),
       "nnz": int(A.nnz), "host_build_s": round(time.time() - t0, 1), "peak_GBs": PEAK, "peak_source": peak_src, "runs": {}}


Aop = utils.MatrixLinearOperator(A)      # one operator: its device formats (native CSR, embedded CSR) are uploaded once


def solve(native, ortho, cycles):
    utils._NATIVE_Z = native
    ls = kp.linsys.LinearSystem(Aop, b)
    try:
        return kp.linsys.RestartedGmres(ls, maxiter=RESTART, max_restarts=cycles - 1, tol=1e-14, ortho=ortho)
    except kp.utils.ConvergenceError as e:
        return e.solver


def native_bytes_per_it():
    orth = np.mean([(2 * (k + 1) + 5) * 16.0 * N for k in range(RESTART)])
    return 20.0 * A.nnz + 32.0 * N + orth + 32.0 * N


hist = {}
for ortho in ("cgs", "mgs"):
    for native in (True, False):
        key = "%s_%s" % ("native" if native else "embedding", ortho)
        solve(native, ortho, 1)                                  # warm-up: uploads, occupancy queries
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s = solve(native, ortho, CYCLES)
        e1.record()
        torch.cuda.synchronize()
        its = len(s.resnorms) - 1
        sec = e0.elapsed_time(e1) * 1e-3
        hist[key] = np.array(s.resnorms)
        r = {"iterations": its, "seconds": sec, "it_per_s": its / sec, "final_resnorm": float(s.resnorms[-1]),
             "includes": "LinearSystem set-up and the upload of b per solve; A is resident (uploaded by the warm-up)"}
        if native:
            gbs = native_bytes_per_it() * its / sec / 1e9
            r.update(native_bytes_per_iteration=native_bytes_per_it(), algorithmic_GBs=gbs,
                     frac_of_measured_peak=gbs / PEAK)
        # per-kernel CUDA events (eager enqueue): one more cycle with the timer on
        ctx.timer = _device.KernelTimer()
        try:
            solve(native, ortho, 2)
            summ = ctx.timer.summary()
        finally:
            ctx.timer = None
        kern = {}
        for tag, d in summ.items():
            ms = np.array(d["ms"])
            ent = {"launches": int(d["launches"]), "avg_us": float(ms.mean() * 1e3)}
            if tag == "spmv":
                by = (20.0 * A.nnz + 32.0 * N) if native else (48.0 * A.nnz + 32.0 * N + 8.0 * N)
                ent.update(bytes_per_launch=by, GBs=by / (ms.mean() * 1e-3) / 1e9)
            if tag == "orth":
                # meta: (real length 2N, real rows, passes, algo, has_next); bytes the launch actually moves
                tot = 0.0
                for (nq, nv, passes, algo, has_next) in d["meta"]:
                    rows = nv / 2.0 if native else nv            # native: complex vectors of 16 N bytes = nq * 8
                    tot += (passes * (2 * rows + 3) + (2 if has_next else 0)) * nq * 8.0
                ent.update(bytes_total=tot, GBs=tot / (ms.sum() * 1e-3) / 1e9)
            if "GBs" in ent:
                ent["frac_of_measured_peak"] = ent["GBs"] / PEAK
            kern[tag] = ent
        r["kernels"] = kern
        out["runs"][key] = r
        print("%-16s %8.1f it/s  %s" % (key, r["it_per_s"], {k: (round(v["avg_us"], 1), round(v.get("GBs", 0))) for k, v in kern.items()}),
              file=sys.stderr)
    a, e = hist["native_" + ortho], hist["embedding_" + ortho]
    m = min(len(a), len(e))
    out["runs"]["native_" + ortho]["speedup_over_embedding"] = (
        out["runs"]["native_" + ortho]["it_per_s"] / out["runs"]["embedding_" + ortho]["it_per_s"])
    live = e[:m] > 1e-8          # (relative residuals; below that the two roundings differ relatively, not absolutely)
    out["runs"]["native_" + ortho]["history_max_rel_diff_vs_embedding"] = float(
        np.max(np.abs(a[:m] - e[:m])[live] / e[:m][live]))
    out["runs"]["native_" + ortho]["history_lengths"] = [int(len(a)), int(len(e))]
utils._NATIVE_Z = True
print(json.dumps(out))
