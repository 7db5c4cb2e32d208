"""One FULL GMRES(30) cycle of the UNMODIFIED reference (baseline/_ref, krypy.linsys.Gmres through
its public API) on the C2 system at full size, timed on this box's host cores, next to the oracle
port's full cycle -- the like-for-like CPU number behind bench.py's bounded `cpu_baseline` sample.
    python tools/reference_full_cycle.py > profiles/r2_reference_full_cycle.json      (~5 min)"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from krypy_b200 import problems  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_GRID
    warnings.simplefilter("ignore")
    A = problems.laplace2d(n)
    b = problems.rhs_normal(n * n)
    krypy = bench.load_reference()
    out = {"workload": "C2 one full cycle: krypy.linsys.Gmres(LinearSystem(A, b), maxiter=30, tol=1e-12), N=%d" % (n * n),
           "host_cpus": os.cpu_count(), "blas_threads": bench.cpu_threads()}
    try:
        from threadpoolctl import threadpool_info
        out["threadpool_info"] = threadpool_info()
    except Exception as exc:   # noqa: BLE001
        out["threadpool_info"] = repr(exc)
    if krypy is not None:
        it, dt, rn = bench.reference_cycle(krypy, A, b, 10 ** 9)
        out["reference"] = {"kind": "reference (unmodified, baseline/_ref)", "iterations": it, "seconds": dt,
                            "it_per_s": it / dt, "resnorms": rn}
    t = time.perf_counter()
    it, dt, rn = bench.cpu_cycle(A, b, 30)
    out["port"] = {"kind": "oracle/krylov_oracle.py", "iterations": it, "seconds": dt, "it_per_s": it / dt,
                   "resnorms": rn}
    if krypy is not None:
        a, r = np.array(out["reference"]["resnorms"]), np.array(rn)
        out["port_vs_reference_history_max_rel"] = float(np.max(np.abs(a - r) / r))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
