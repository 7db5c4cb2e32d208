"""Diagnostic for the recycled-solve iteration counts of test/test_recycling.py:8-39 (reference):
prints, per solver / criterion / solve, the history length, the tail of the residual history and
the Ritz values + Ritz residual norms of the selected deflation vectors.
    python tools/diag_recycling.py ref   (build container, unmodified reference)
    python tools/diag_recycling.py gpu   (B200, krypy_b200)
TEST/ANALYSIS TOOL ONLY."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(impl, out):
    if impl == "ref":
        from oracle import refshim
        kp = refshim.import_reference()
    else:
        import krypy_b200 as kp
    warnings.simplefilter("ignore")
    N = 100
    d = np.linspace(1, 2, N)
    d[:5] = [1e-8, 1e-4, 1e-2, 2e-2, 3e-2]
    res = {}
    for sname in ("cg", "minres", "gmres"):
        Solver = {"cg": kp.recycling.RecyclingCg, "minres": kp.recycling.RecyclingMinres,
                  "gmres": kp.recycling.RecyclingGmres}[sname]
        for which in ("lm", "sm", "lr", "sr", "li", "si", "smallest_res"):
            ls = kp.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True,
                                        positive_definite=True)
            fac = kp.recycling.factories.RitzFactorySimple(n_vectors=3, which=which)
            rs = Solver()
            rec = []
            for i in range(3):
                s = rs.solve(ls, vector_factory=fac, maxiter=50, tol=1e-5, x0=None)
                r = kp.deflation.Ritz(s, mode="ritz")
                key = {"lm": -np.abs(r.values), "sm": np.abs(r.values), "lr": -np.real(r.values),
                       "sr": np.real(r.values), "li": -np.imag(r.values), "si": np.imag(r.values),
                       "smallest_res": np.asarray(r.resnorms)}[which]
                idx = np.argsort(key)[:3]
                rec.append({"len": len(s.resnorms), "tail": [float(x) for x in s.resnorms[-4:]],
                            "sel_values": [complex(v).real for v in np.asarray(r.values)[idx]],
                            "sel_resnorms": [float(v) for v in np.asarray(r.resnorms)[idx]],
                            "sorted_resnorms_head": [float(v) for v in np.sort(r.resnorms)[:6]]})
            res["%s/%s" % (sname, which)] = rec
            print(sname, which, [x["len"] for x in rec], flush=True)
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "/dev/stdout")
