"""Shared checks for deflation.Ritz and recycling (used by the CPU tier over the device test
double and by the GPU tier over the real kernels)."""
import warnings

import numpy as np

import cases
import runners


def _deflated(kp, name):
    c = cases.case_inputs(name)
    ls = kp.linsys.LinearSystem(c["A"], c["b"], **c["ls"])
    kw = dict(c["kw"]); kw["store_arnoldi"] = True
    cls = {"gmres": kp.deflation.DeflatedGmres, "cg": kp.deflation.DeflatedCg,
           "minres": kp.deflation.DeflatedMinres}[c["solver"]]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return cls(ls, **kw), ls
        except kp.utils.ConvergenceError as e:
            return e.solver, ls


def _order(values):
    """sort by (real part rounded to 1e-8, imaginary part): stable for conjugate pairs whose real
    parts differ by round-off"""
    v = np.asarray(values)
    return np.lexsort((np.imag(v), np.round(np.real(v), 8)))


def check_ritz_pairs(name):
    """Ritz / harmonic Ritz values and residual norms against the reference (fixtures from
    oracle/make_golden_ritz.py).  Only the part of the spectrum that is insensitive to the
    rounding-level differences of the basis is compared tightly: pairs with a small residual."""
    import krypy_b200 as kp
    gold = runners.load_golden("ritz_recycling")
    sol, ls = _deflated(kp, name)
    for mode in ("ritz", "harmonic"):
        r = kp.deflation.Ritz(sol, mode=mode)
        order = _order(r.values)
        vals, res = np.asarray(r.values)[order], np.asarray(r.resnorms)[order]
        gv, gr = gold["%s__%s__values" % (name, mode)], gold["%s__%s__resnorms" % (name, mode)]
        go = _order(gv)
        gv, gr = gv[go], gr[go]
        assert vals.shape == gv.shape and res.shape == gr.shape
        finite = np.isfinite(gv)
        scale = np.abs(gv[finite]).max()
        # eigenvalues of a (nearly) normal small matrix: perturbation ~ basis noise
        assert np.abs(vals[finite] - gv[finite]).max() <= 1e-6 * scale, (name, mode)
        assert np.abs(res[finite] - gr[finite]).max() <= 1e-6 * max(np.abs(gr[finite]).max(), 1e-300)
        assert r.coeffs.shape[0] == sol.H.shape[1] + sol.projection.U.shape[1]
        np.testing.assert_allclose(np.linalg.norm(r.coeffs, axis=0), 1.0, rtol=1e-12)
    # explicit residual norms agree with the small-matrix formula (reference semantics) and fixtures
    r = kp.deflation.Ritz(sol, mode="ritz")
    if not np.iscomplexobj(r.values) or np.abs(np.imag(r.values)).max() == 0:
        order = _order(r.values)
        ex = r.get_explicit_resnorms()[order]
        ge = gold["%s__ritz__explicit_resnorms" % name][_order(gold["%s__ritz__values" % name])]
        assert np.abs(ex - ge).max() <= 1e-6 * max(np.abs(ge).max(), 1e-300)
        np.testing.assert_allclose(ex, np.asarray(r.resnorms)[order], rtol=1e-4, atol=1e-8 * np.abs(ge).max())
        V = r.get_vectors([0, 1])
        assert V.shape == (ls.N, 2)


def check_recycling(sname, which):
    """reference test/test_recycling.py:8-39 with the iteration counts the reference produces.

    'smallest_res', third solve: after the first recycled solve the Ritz pairs of the deflated
    eigenvalues 0.01, 0.02, 0.03 and of 1e-8, 1e-4 are ALL converged to rounding level (Ritz residual
    norms 5e-15 .. 3e-10, i.e. square roots of cancellation noise -- measured on the reference and
    on the B200, tools/diag_recycling.py, profiles/r2_recycling_diag.json).  Which of 0.01/0.02/0.03
    the criterion ranks third is therefore decided by rounding noise: the reference's arithmetic
    picks 0.01 (14 iterations), the B200's picked 0.03 in the CG run (15 iterations) -- and the
    UNMODIFIED reference also needs 15 when it is handed that selection (fixture
    ``solve3_len_by_third``, oracle/make_golden_ritz.py).  The check is therefore: solves 1 and 2
    exactly as the reference; every selected pair converged to the noise floor; solve 3 exactly the
    reference's count for the selection that was made."""
    import krypy_b200 as kp
    gold = runners.load_golden("ritz_recycling")
    N = 100
    d = np.linspace(1, 2, N)
    d[:5] = [1e-8, 1e-4, 1e-2, 2e-2, 3e-2]
    ls = kp.linsys.LinearSystem(np.diag(d), np.ones((N, 1)), normal=True, self_adjoint=True, positive_definite=True)
    Solver = {"cg": kp.recycling.RecyclingCg, "minres": kp.recycling.RecyclingMinres,
              "gmres": kp.recycling.RecyclingGmres}[sname]
    fac = kp.recycling.factories.RitzFactorySimple(n_vectors=3, which=which)
    rs = Solver()
    lens = []
    want = list(gold["recycling__%s__%s__lens" % (sname, which)])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(3):
            if which == "smallest_res" and i == 2:
                r = kp.deflation.Ritz(rs.last_solver, mode="ritz")
                order = np.argsort(r.resnorms)[:3]
                sel = np.sort(np.real(np.asarray(r.values)[order]))
                assert np.all(np.asarray(r.resnorms)[order] < 1e-8), r.resnorms        # converged to the noise floor
                np.testing.assert_allclose(sel[:2], [1e-8, 1e-4], rtol=1e-4)
                third = int(np.argmin(np.abs(np.array([0.01, 0.02, 0.03]) - sel[2])))
                assert abs(sel[2] - (0.01, 0.02, 0.03)[third]) < 1e-9
                want[2] = int(gold["recycling__%s__smallest_res__solve3_len_by_third" % sname][third])
            s = rs.solve(ls, vector_factory=fac, maxiter=50, tol=1e-5, x0=None)
            lens.append(len(s.resnorms))
            assert s.resnorms[-1] <= 1e-5
            _, _, rn = ls.get_residual(s.xk, compute_norm=True)
            np.testing.assert_almost_equal(s.resnorms[-1], rn / ls.MMlb_norm, decimal=12)
            if i > 0:
                assert lens[-1] <= lens[0]
    assert lens == want, (lens, want, sname, which)
