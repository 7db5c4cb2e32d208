"""Run BASELINE.json's configurations C2..C5 at FULL size on one B200 through the public API and
report iterations/s, algorithmic GB/s (SURVEY 8d byte model) and size-independent parity
properties.  usage: python tools/run_configs.py [c2 c3 c4 c4r c5] > profiles/r1_configs.json"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krypy_b200 as kp
from krypy_b200 import problems, utils as u, _device

warnings.simplefilter("ignore")
which = [a.lower() for a in sys.argv[1:]] or ["c2", "c3", "c4", "c5"]
ctx = _device.Context.get()
PEAK = 6540.5
out = {}


def timed(fn):
    torch.cuda.synchronize(); t = time.perf_counter()
    try:
        s = fn()
    except kp.utils.ConvergenceError as e:
        s = e.solver
    torch.cuda.synchronize()
    return s, time.perf_counter() - t


def explicit_relres(A, b, x, M=None):
    r = b.reshape(-1).astype(np.float64) - A.astype(np.float64) @ x.reshape(-1).astype(np.float64)
    if M is not None:
        return np.sqrt(r @ (M @ r)) / np.sqrt(b.reshape(-1) @ (M @ b.reshape(-1)))
    return np.linalg.norm(r) / np.linalg.norm(b)


if "c2" in which:
    n = 3162; N = n * n
    A = problems.laplace2d(n); b = problems.rhs_normal(N)
    ls = kp.linsys.LinearSystem(A, b)
    timed(lambda: kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=1, tol=1e-12, ortho="cgs"))
    s, dt = timed(lambda: kp.linsys.RestartedGmres(ls, maxiter=30, max_restarts=4, tol=1e-12, ortho="cgs"))
    its = len(s.resnorms) - 1
    rn = np.array(s.resnorms)
    out["c2"] = {"config": "GMRES(30) 2-D 5-pt Laplacian N=%d fp64, 5 cycles, ortho=cgs" % N, "iterations": its,
                 "seconds": dt, "it_per_s": its / dt, "algorithmic_GBs": 383.5 * N * its / dt / 1e9,
                 "frac_of_measured_peak": 383.5 * N * its / dt / 1e9 / PEAK,
                 "resnorms_nonincreasing_within_cycles": bool(np.all(np.diff(rn)[np.arange(its) % 30 != 29] <= 1e-12)),
                 "final_resnorm": float(rn[-1]), "explicit_check": float(explicit_relres(A, b, s.xk)),
                 "first4": rn[:4].tolist()}
    del ls, s

if "c3" in which:
    n = 400; N = n ** 3
    t0 = time.time(); A = problems.poisson3d(n); b = problems.rhs_normal(N); M = problems.jacobi_csr(A)
    tb = time.time() - t0
    ls = kp.linsys.LinearSystem(A, b, M=M, self_adjoint=True, positive_definite=True)
    timed(lambda: kp.linsys.Cg(ls, tol=1e-8, maxiter=5))
    s, dt = timed(lambda: kp.linsys.Cg(ls, tol=1e-8, maxiter=200))
    its = len(s.resnorms) - 1
    out["c3"] = {"config": "CG + Jacobi(csr diag) 3-D 7-pt Poisson N=%d nnz=%d fp64" % (N, A.nnz), "iterations": its,
                 "seconds": dt, "it_per_s": its / dt, "algorithmic_GBs": 192.0 * N * its / dt / 1e9,
                 "frac_of_measured_peak": 192.0 * N * its / dt / 1e9 / PEAK, "host_build_s": tb,
                 "final_resnorm": float(s.resnorms[-1]),
                 "explicit_check": float(explicit_relres(A, b, s.xk, M)), "first4": list(map(float, s.resnorms[:4]))}
    del ls, s, A, M

if "c4" in which:
    n = 2000; N = n * n; d = 20
    A = problems.convdiff2d(n, c=0.1); b = np.ones((N, 1))
    # deflation space: the d lowest sine modes of the 2-D Laplacian (analytic, real; SURVEY F10 treats U as input)
    xs = np.arange(1, n + 1) / (n + 1.0)
    modes = sorted(((p * p + q * q, p, q) for p in range(1, 8) for q in range(1, 8)))[:d]
    U = np.stack([np.outer(np.sin(p * np.pi * xs), np.sin(q * np.pi * xs)).reshape(-1) for _, p, q in modes], axis=1)
    ls = kp.linsys.LinearSystem(A, b)
    timed(lambda: kp.deflation.DeflatedGmres(ls, U=U, maxiter=5, tol=1e-10))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    proj = kp.deflation.ObliqueProjection(ls, U); torch.cuda.synchronize(); tset = time.perf_counter() - t0
    s, dt = timed(lambda: kp.deflation.DeflatedGmres(ls, U=U, maxiter=60, tol=1e-10, ortho="cgs"))
    its = len(s.resnorms) - 1
    UtU = u._inner_dev(s.projection._Ud, s.projection._Ud).cpu().numpy()
    E_explicit = u._inner_dev(s.projection._Ud, s.projection._AUd).cpu().numpy()
    per_it = sum((136 + 16 * k) * N for k in range(its)) / its + 672.0 * N
    out["c4"] = {"config": "DeflatedGmres d=20 (sine modes) 2-D conv-diff N=%d fp64, maxiter=60, ortho=cgs" % N,
                 "iterations": its, "seconds_total": dt, "projector_setup_s": tset, "it_per_s": its / dt,
                 "algorithmic_GBs": per_it * its / dt / 1e9, "frac_of_measured_peak": per_it * its / dt / 1e9 / PEAK,
                 "U_orthonormality": float(np.abs(UtU - np.eye(d)).max()),
                 "E_vs_explicit": float(np.abs(s.E - E_explicit).max() / np.abs(E_explicit).max()),
                 "C_shape": list(s.C.shape), "final_resnorm": float(s.resnorms[-1]),
                 "explicit_check": float(np.linalg.norm(b.reshape(-1) - A @ s.xk.reshape(-1)) / np.linalg.norm(b)),
                 "first4": list(map(float, s.resnorms[:4]))}
    # undeflated run for comparison of the history
    s0, dt0 = timed(lambda: kp.linsys.Gmres(ls, maxiter=60, tol=1e-10, ortho="cgs"))
    out["c4"]["undeflated_final_resnorm"] = float(s0.resnorms[-1]); out["c4"]["undeflated_it_per_s"] = 60 / dt0
    del ls, s, s0, proj

if "c4r" in which:
    # SURVEY 8d's C4 workflow end to end on the device: solve 1 (no deflation, store_arnoldi) ->
    # 20 Ritz vectors of smallest magnitude (RitzFactorySimple, realified, kept in HBM) -> timed solve 2
    n = 2000; N = n * n; d = 20
    A = problems.convdiff2d(n, c=0.1); b = np.ones((N, 1))
    ls = kp.linsys.LinearSystem(A, b)
    fac = kp.recycling.factories.RitzFactorySimple(n_vectors=d, which="sm")
    rs = kp.recycling.RecyclingGmres()
    s1, dt1 = timed(lambda: rs.solve(ls, vector_factory=fac, maxiter=60, tol=1e-10, ortho="cgs"))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Ublk = fac.get(s1); torch.cuda.synchronize(); tfac = time.perf_counter() - t0   # (s1 may come from a ConvergenceError)
    s2, dt2 = timed(lambda: kp.deflation.DeflatedGmres(ls, U=Ublk, maxiter=60, tol=1e-10, ortho="cgs",
                                                       store_arnoldi=True))
    out["c4r"] = {"config": "recycled DeflatedGmres: 20 Ritz vectors ('sm') of solve 1, conv-diff N=%d fp64" % N,
                  "solve1_it_per_s": (len(s1.resnorms) - 1) / dt1, "ritz_vectors_s": tfac,
                  "solve2_iterations": len(s2.resnorms) - 1, "solve2_seconds_total": dt2,
                  "solve2_it_per_s": (len(s2.resnorms) - 1) / dt2,
                  "solve1_final_resnorm": float(s1.resnorms[-1]), "solve2_final_resnorm": float(s2.resnorms[-1]),
                  "solve2_explicit_check": float(np.linalg.norm(b.reshape(-1) - A @ s2.xk.reshape(-1)) / np.linalg.norm(b)),
                  "U_shape": list(s2.projection.U.shape)}
    del ls, s1, s2, rs, Ublk

if "c5" in which:
    n = 4000; N = n * n
    A, B = problems.shifted_laplace_B(n, sigma=0.3, dtype=np.float32)
    b = problems.rhs_normal(N, dtype=np.float32)
    res = {}
    for dt_name, dtp in (("fp32", np.float32), ("fp64", np.float64)):
        ls = kp.linsys.LinearSystem(A, b, ip_B=B, self_adjoint=True, dtype=dtp)
        timed(lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=5))
        s, dt = timed(lambda: kp.linsys.Minres(ls, tol=1e-5, maxiter=50))
        its = len(s.resnorms) - 1
        sz = 4 if dtp == np.float32 else 8
        by = (A.nnz * (sz + 4) + 4 * N + 2 * N * sz) + 2 * (N * (sz + 4) + 4 * N + 2 * N * sz) + 13 * N * sz
        res[dt_name] = {"iterations": its, "seconds": dt, "it_per_s": its / dt, "algorithmic_GBs": by * its / dt / 1e9,
                        "frac_of_measured_peak": by * its / dt / 1e9 / PEAK, "resnorms": list(map(float, s.resnorms))}
        del ls, s
    a, r = np.array(res["fp32"]["resnorms"]), np.array(res["fp64"]["resnorms"])
    m = min(len(a), len(r))
    out["c5"] = {"config": "MINRES(lanczos) ip_B=diag SPD CSR, A=B^-1(L-0.3I) 2-D N=%d, maxiter=50" % N,
                 "fp32": {k: v for k, v in res["fp32"].items() if k != "resnorms"},
                 "fp64": {k: v for k, v in res["fp64"].items() if k != "resnorms"},
                 "fp32_vs_fp64_history_max_rel": float(np.max(np.abs(a[:m] - r[:m]) / r[:m])),
                 "final_resnorm_fp32": float(a[-1]), "final_resnorm_fp64": float(r[-1])}

print(json.dumps(out, indent=1))
