"""Fixtures for the section-8f "next" row (deflation.Ritz + recycling.RitzFactorySimple), generated
by the UNMODIFIED reference:  python oracle/make_golden_ritz.py   (TEST INFRASTRUCTURE ONLY)"""
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


def main():
    krypy = refshim.import_reference()
    out = {}
    warnings.simplefilter("ignore")
    # 1. Ritz pairs of deflated solves (values sorted by (real, imag); residual norms in that order)
    for name in ("convdiff_defl_gmres", "lap2d_defl_minres", "lap2d_defl_cg", "c1_gmres_defl"):
        c = cases.case_inputs(name)
        ls = krypy.linsys.LinearSystem(c["A"], c["b"], **c["ls"])
        kw = dict(c["kw"]); kw["store_arnoldi"] = True
        cls = {"gmres": krypy.deflation.DeflatedGmres, "cg": krypy.deflation.DeflatedCg,
               "minres": krypy.deflation.DeflatedMinres}[c["solver"]]
        try:
            sol = cls(ls, **kw)
        except krypy.utils.ConvergenceError as e:
            sol = e.solver
        for mode in ("ritz", "harmonic"):
            r = krypy.deflation.Ritz(sol, mode=mode)
            order = np.lexsort((np.imag(r.values), np.real(r.values)))
            out["%s__%s__values" % (name, mode)] = np.asarray(r.values)[order]
            out["%s__%s__resnorms" % (name, mode)] = np.asarray(r.resnorms)[order]
            if mode == "ritz":
                ex = r.get_explicit_resnorms()
                out["%s__ritz__explicit_resnorms" % name] = np.asarray(ex)[order]
    # 2. recycling scenario of the reference's test/test_recycling.py:8-39
    N = 100
    d = np.linspace(1, 2, N)
    d[:5] = [1e-8, 1e-4, 1e-2, 2e-2, 3e-2]
    for sname, Solver in (("cg", krypy.recycling.RecyclingCg), ("minres", krypy.recycling.RecyclingMinres),
                          ("gmres", krypy.recycling.RecyclingGmres)):
        for which in ("lm", "sm", "lr", "sr", "li", "si", "smallest_res"):
            ls = krypy.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True,
                                           positive_definite=True)
            fac = krypy.recycling.factories.RitzFactorySimple(n_vectors=3, which=which)
            rs = Solver()
            lens, finals = [], []
            for i in range(3):
                s = rs.solve(ls, vector_factory=fac, maxiter=50, tol=1e-5, x0=None)
                lens.append(len(s.resnorms))
                finals.append(s.resnorms[-1])
            out["recycling__%s__%s__lens" % (sname, which)] = np.array(lens)
            out["recycling__%s__%s__finals" % (sname, which)] = np.array(finals)
            print(sname, which, lens)
        # 'smallest_res' after the first recycled solve: the Ritz pairs of the deflated eigenvalues
        # (0.01, 0.02, 0.03) and of 1e-8, 1e-4 are all converged to rounding level (residual norms
        # 1e-15 .. 3e-10 = sqrt of cancellation noise), so WHICH of 0.01/0.02/0.03 becomes the third
        # deflation vector of solve 3 is decided by rounding noise.  Record the reference's iteration
        # count of solve 3 for each possible outcome (selection forced by Ritz value).
        by_third = []
        for third in (0.01, 0.02, 0.03):
            ls = krypy.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True,
                                           positive_definite=True)
            fac = krypy.recycling.factories.RitzFactorySimple(n_vectors=3, which="smallest_res")
            rs = Solver()
            for i in range(2):
                rs.solve(ls, vector_factory=fac, maxiter=50, tol=1e-5, x0=None)
            s = rs.solve(ls, vector_factory=cases.NearestRitzValuesFactory(krypy, (1e-8, 1e-4, third)),
                         maxiter=50, tol=1e-5, x0=None)
            by_third.append(len(s.resnorms))
        out["recycling__%s__smallest_res__solve3_len_by_third" % sname] = np.array(by_third)
        print(sname, "solve-3 length by third deflated value (0.01, 0.02, 0.03):", by_third)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ritz_recycling.npz"), **out)


if __name__ == "__main__":
    main()
