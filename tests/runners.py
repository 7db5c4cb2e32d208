"""Run a named case (tests/cases.py) through the oracle or through the product
(krypy_b200, CUDA) and return the same result dictionary layout as the golden
fixtures written by oracle/make_golden.py."""
import warnings

import numpy as np

import cases


def load_golden(name):
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")
    with np.load(p) as z:
        return {k: z[k] for k in z.files}


def run_oracle(name):
    from oracle import krylov_oracle as ko

    c = cases.case_inputs(name)
    ls = dict(c["ls"])
    ls.pop("self_adjoint", None), ls.pop("positive_definite", None)
    B = ls.pop("ip_B", None)
    sysm = ko.System(c["A"], c["b"], B=B, **ls)
    kw = dict(c["kw"])
    store = kw.pop("store_arnoldi", False)
    solver = c["solver"]
    converged = True
    try:
        if solver == "restarted_gmres":
            sol = ko.restarted_gmres(sysm, **kw)
        elif solver == "gmres":
            sol = ko.gmres(sysm, **kw)
        elif solver == "minres":
            sol = ko.minres(sysm, **kw)
        else:
            sol = ko.cg(sysm, store_arnoldi=store, **kw)
    except ko.OracleConvergenceError as e:
        sol = e.result
        converged = False
    out = dict(resnorms=np.array(sol.resnorms, dtype=np.float64), xk=np.asarray(sol.xk),
               converged=np.array(converged))
    if hasattr(sol, "iter"):
        out["iter"] = np.array(sol.iter)
    if store:
        out["H"] = np.asarray(sol.H)
        out["V_shape"] = np.array(sol.Vk.shape)
        out["V_colsum_abs"] = np.abs(sol.Vk).sum(axis=0)
    if "U" in kw:
        out["C"] = np.asarray(sol.C)
        out["E"] = np.asarray(sol.E)
        out["UMlr"] = np.asarray(sol.UMlr)
    if solver == "cg":
        out["rhos"] = np.array(sol.rhos, dtype=np.float64)
    return out


def run_product(name, **override):
    import krypy_b200 as kp

    c = cases.case_inputs(name)
    ls = kp.linsys.LinearSystem(c["A"], c["b"], **c["ls"])
    kw = dict(c["kw"])
    kw.update(override)
    U = kw.pop("U", None)
    solver = c["solver"]
    if solver == "restarted_gmres":
        cls = kp.linsys.RestartedGmres
    elif U is not None:
        cls = {"gmres": kp.deflation.DeflatedGmres, "cg": kp.deflation.DeflatedCg,
               "minres": kp.deflation.DeflatedMinres}[solver]
        kw["U"] = U
    else:
        cls = {"gmres": kp.linsys.Gmres, "cg": kp.linsys.Cg, "minres": kp.linsys.Minres}[solver]
    converged = True
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            sol = cls(ls, **kw)
        except kp.utils.ConvergenceError as e:
            sol = e.solver
            converged = False
    out = dict(resnorms=np.array(sol.resnorms, dtype=np.float64), xk=np.asarray(sol.xk),
               converged=np.array(converged), _sol=sol)
    if hasattr(sol, "iter"):
        out["iter"] = np.array(sol.iter)
    if kw.get("store_arnoldi"):
        out["H"] = np.asarray(sol.H)
        V = np.asarray(sol.V)
        out["V_shape"] = np.array(V.shape)
        out["V_colsum_abs"] = np.abs(V).sum(axis=0)
    if U is not None:
        out["C"] = np.asarray(sol.C)
        out["E"] = np.asarray(sol.E)
        out["UMlr"] = np.asarray(sol.UMlr)
    if solver == "cg":
        out["rhos"] = np.array(sol.rhos, dtype=np.float64)
    return out
