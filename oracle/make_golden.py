"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/refshim.py) on the named cases of
tests/cases.py.  Run in the build container:

    python oracle/make_golden.py

The fixtures hold only the reference's outputs (residual histories, solution,
Hessenberg/tridiagonal matrices, deflation matrices); inputs are regenerated
from seeds by tests/cases.py.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refshim  # noqa: E402
import cases  # noqa: E402


def run_reference(krypy, name):
    c = cases.case_inputs(name)
    ls = krypy.linsys.LinearSystem(c["A"], c["b"], **c["ls"])
    kw = dict(c["kw"])
    U = kw.pop("U", None)
    solver = c["solver"]
    if solver == "restarted_gmres":
        cls = krypy.linsys.RestartedGmres
    elif U is not None:
        cls = {"gmres": krypy.deflation.DeflatedGmres, "cg": krypy.deflation.DeflatedCg,
               "minres": krypy.deflation.DeflatedMinres}[solver]
        kw["U"] = U
    else:
        cls = {"gmres": krypy.linsys.Gmres, "cg": krypy.linsys.Cg,
               "minres": krypy.linsys.Minres}[solver]
    converged = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = cls(ls, **kw)
        except krypy.utils.ConvergenceError as e:
            sol = e.solver
            converged = False
    out = dict(resnorms=np.array(sol.resnorms, dtype=np.float64), xk=np.asarray(sol.xk),
               converged=np.array(converged))
    if hasattr(sol, "iter"):
        out["iter"] = np.array(sol.iter)
    if kw.get("store_arnoldi"):
        out["H"] = np.asarray(sol.H)
        out["V_shape"] = np.array(sol.V.shape)
        # orthonormality / Arnoldi relation are checked via properties; keep a
        # cheap signature of V instead of the whole basis
        out["V_colsum_abs"] = np.abs(sol.V).sum(axis=0)
    if U is not None:
        out["C"] = np.asarray(sol.C)
        out["E"] = np.asarray(sol.E)
        out["UMlr"] = np.asarray(sol.UMlr)
    if solver == "cg":
        out["rhos"] = np.array(sol.rhos, dtype=np.float64)
    return out


def givens_table(krypy):
    pts = [(3.0, 4.0), (-3.0, 4.0), (3.0, -4.0), (-4.0, 3.0), (0.0, 2.0), (0.0, -2.0),
           (-2.0, 0.0), (0.0, 0.0), (1e200, 1e200), (1e-200, 1e-200), (1.0, 1e-8),
           (1e8, 1.0), (-1.5, -2.5), (2.5, -1.5)]
    rows = []
    for a, b in pts:
        g = krypy.utils.Givens(np.array([[a], [b]]))
        rows.append([a, b, g.c, g.s, g.r])
    return np.array(rows)


def main():
    if not refshim.available():
        raise SystemExit("reference not present at %s" % refshim.REFERENCE_ROOT)
    krypy = refshim.import_reference()
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    # python oracle/make_golden.py [name ...]: only the named cases (default: all)
    names = sys.argv[1:] or (cases.ALL_CASES + cases.COMPLEX_CASES)
    for name in names:
        out = run_reference(krypy, name)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)
        print("%-28s its=%3d last=%.6e conv=%s" % (
            name, len(out["resnorms"]) - 1, out["resnorms"][-1], bool(out["converged"])))
    if not sys.argv[1:]:
        np.savez_compressed(os.path.join(outdir, "givens_table.npz"), table=givens_table(krypy))
        print("givens table written")


if __name__ == "__main__":
    main()
