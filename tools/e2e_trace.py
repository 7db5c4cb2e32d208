"""Per-phase wall time of bench.py's e2e step (KRY_TRACE marks inside the solver), to find sporadic
slow steps.  ANALYSIS TOOL.   KRY_TRACE=1 python tools/e2e_trace.py"""
import os, sys, time, gc, warnings
os.environ["KRY_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp, torch
import krypy_b200 as kp
from krypy_b200 import problems
warnings.simplefilter("ignore")
n = 3162
A = problems.laplace2d(n); b = problems.rhs_normal(n * n)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t.numpy(), t
keep = []; parts = []
for arr in (A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32)):
    v, t = pin(arr); keep.append(t); parts.append(v)
Ah = sp.csr_matrix((parts[0], parts[1], parts[2]), shape=A.shape); Ah.has_sorted_indices = True
bh, tb = pin(b); keep.append(tb)
xk = None
for i in range(10):
    if len(sys.argv) > 1 and sys.argv[1] == "nogc" and i == 0:
        gc.disable()
    g0 = gc.get_count()
    t0 = time.perf_counter()
    ls = kp.linsys.LinearSystem(Ah, bh)
    t1 = time.perf_counter()
    try:
        sol = kp.linsys.Gmres(ls, x0=xk, maxiter=30, tol=1e-12, ortho="cgs")
    except kp.utils.ConvergenceError as e:
        sol = e.solver
    t2 = time.perf_counter()
    xk = sol.xk
    t3 = time.perf_counter()
    tr = sol.__dict__.get("_trace", [])
    ph = ", ".join("%s=%.1f" % (nm, 1e3 * (tt - t1)) for nm, tt in tr)
    print("step %d: ls %.1f  gmres %.1f  xk %.1f ms | %s | gc %s mem %.2f GB reserved" % (
        i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), ph, g0, torch.cuda.memory_reserved() / 1e9), flush=True)
    del sol, ls
