"""Evaluators that rate a subset of Ritz pairs as deflation space by the estimated time of the next
solve (krypy/recycling/evaluators.py).  Host logic; the Arnoldi relations behind
``RitzApproxKrylov`` come from ``deflation.Arnoldifyer``, whose N-sized parts run on the device."""
import warnings

import numpy

from .. import deflation, utils


class _RitzSubsetEvaluator(object):
    """Abstract base (krypy/recycling/evaluators.py:6-10)."""

    def evaluate(self, ritz, subset):
        raise NotImplementedError("abstract base class cannot be instanciated")


class RitzApriori(_RitzSubsetEvaluator):
    """A-priori bound (``utils.BoundCG`` / ``utils.BoundMinres``) on the Ritz values that are NOT
    deflated -- or on inclusion intervals around them (``strategy='intervals'``) -- turned into a
    time estimate (krypy/recycling/evaluators.py:13-134).  For self-adjoint problems."""

    def __init__(self, Bound, tol=None, strategy="simple", deflweight=1.0):
        self.Bound = Bound
        self.tol = tol
        self.strategy = strategy
        self.deflweight = deflweight

    def evaluate(self, ritz, subset):
        solver = ritz._deflated_solver
        if not solver.linear_system.self_adjoint:
            warnings.warn("RitzApriori is designed for self-adjoint problems but the provided "
                          "LinearSystem is not marked as self-adjoint.")
        tol = solver.tol if self.tol is None else self.tol
        chosen = list(subset)
        others = list(set(range(len(ritz.values))).difference(subset))
        if self.strategy == "simple":
            spectrum = ritz.values[others]
        elif self.strategy == "intervals":
            spectrum = self._estimate_eval_intervals(ritz, chosen, others)
        else:
            raise utils.ArgumentError("Invalid value '%s' for argument 'strategy'. Valid are simple and "
                                      "intervals." % self.strategy)
        nsteps = self.Bound(spectrum).get_step(tol)
        return solver.estimate_time(nsteps, len(subset), deflweight=self.deflweight)

    @staticmethod
    def _estimate_eval_intervals(ritz, indices, indices_remaining, eps_min=0, eps_max=0, eps_res=None):
        """Inclusion intervals for the eigenvalues that remain after deflating the selected Ritz pairs
        (eigenvalue inclusion theorem + heuristic; krypy/recycling/evaluators.py:72-134)."""
        mk = utils.Interval
        if len(indices) == 0:
            return utils.Intervals([mk(mu - r, mu + r) for mu, r in zip(ritz.values, ritz.resnorms)])
        if len(ritz.values) == len(indices):
            raise utils.AssumptionError("selection of all Ritz pairs does not allow estimation.")
        if eps_res is None:
            eps_res = numpy.max(numpy.abs([eps_min, eps_max]))
        res_sel = numpy.linalg.norm(ritz.resnorms[indices], 2)
        res_rest = numpy.linalg.norm(ritz.resnorms[indices_remaining], 2)
        delta = utils.gap(ritz.values[indices], ritz.values[indices_remaining])
        mu_min = utils.Intervals([mk(mu + eps_min, mu + eps_max) for mu in ritz.values[indices]]).min_abs()
        if res_sel + eps_max - eps_min >= delta:
            raise utils.AssumptionError(
                "delta_sel + delta_non_sel + eps_max - eps_min >= delta (%s >= %s)"
                % (res_sel + res_rest + eps_max - eps_min, delta))
        if mu_min == 0:
            raise utils.AssumptionError("mu_min == 0 not allowed")
        eta = (res_sel + eps_res) ** 2 * (1 / (delta - eps_max + eps_min) + 1 / mu_min)
        return utils.Intervals([mk(mu + eps_min - eta, mu + eps_max + eta)
                                for mu in ritz.values[indices_remaining]])


class RitzApproxKrylov(_RitzSubsetEvaluator):
    """Predicts the residual norms of the next solve with an approximate Krylov subspace built from
    the last solve's data (``deflation.bound_pseudo``) and converts the iteration count into a time
    estimate (krypy/recycling/evaluators.py:137-243)."""

    def __init__(self, mode="extrapolate", tol=None, pseudospectra=False, bound_pseudo_kwargs=None,
                 deflweight=1.0):
        self._arnoldifyer = None
        self.mode = mode
        self.tol = tol
        self.pseudospectra = pseudospectra
        self.bound_pseudo_kwargs = bound_pseudo_kwargs or {}
        self.deflweight = deflweight

    def evaluate(self, ritz, subset):
        solver = ritz._deflated_solver
        tol = solver.tol if self.tol is None else self.tol
        if self._arnoldifyer is None or self._arnoldifyer._deflated_solver is not solver:
            self._arnoldifyer = deflation.Arnoldifyer(solver)        # cached per solve
        bound = deflation.bound_pseudo(
            self._arnoldifyer, ritz.coeffs[:, list(subset)], tol=tol,
            pseudo_type="auto" if self.pseudospectra else "omit", **self.bound_pseudo_kwargs)
        if len(bound) <= 1:
            raise utils.AssumptionError("no bound computed")
        if self.mode == "direct":
            if (bound > tol).all():
                raise utils.AssumptionError("tolerance not reached with mode==`direct`.")
            nsteps = (bound > tol).sum()
        elif self.mode == "extrapolate":
            # slowest average residual reduction per step over all prefixes
            steps = numpy.arange(1, len(bound))
            alpha = numpy.max((bound[1:] / bound[0]) ** (1.0 / steps))
            if alpha >= 1 or alpha == 0:
                raise utils.AssumptionError("Cannot compute bound because alpha == %s >= 1" % alpha)
            nsteps = numpy.log(tol / bound[0]) / numpy.log(alpha)
        else:
            raise utils.ArgumentError("Invalid value `%s` for argument `omode`. Valid are `direct` and "
                                      "`extrapolate`." % self.mode)
        return solver.estimate_time(nsteps, len(subset), deflweight=self.deflweight)
