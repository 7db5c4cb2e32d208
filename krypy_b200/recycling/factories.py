"""Deflation-vector factories (krypy/recycling/factories.py)."""
import numpy

from .. import deflation, utils


class _DeflationVectorFactory(object):
    """Abstract base class for selectors (krypy/recycling/factories.py:9-17)."""

    def get(self, solver):
        raise NotImplementedError("abstract base class cannot be instanciated")


class RitzFactory(_DeflationVectorFactory):
    """Greedy selection of Ritz vectors: starting from the empty set, a generator proposes index sets
    to add, an evaluator rates every enlarged set (estimated time of the next solve), the best one is
    kept; the overall best rated set wins (krypy/recycling/factories.py:20-139)."""

    def __init__(self, subset_evaluator, subsets_generator=None, mode="ritz", print_results=None, realify=True):
        from . import generators
        self.subsets_generator = generators.RitzSmall() if subsets_generator is None else subsets_generator
        self.subset_evaluator = subset_evaluator
        self.mode = mode
        self.print_results = print_results
        self.realify = realify      # new, as in RitzFactorySimple: False = the reference's complex vectors

    def get(self, deflated_solver):
        ritz = deflation.Ritz(deflated_solver, mode=self.mode)
        return ritz.get_vectors_dev(self._get_best_subset(ritz), realify=self.realify)

    def _rate(self, ritz, subset, table):
        try:
            table[subset] = self.subset_evaluator.evaluate(ritz, subset)
        except utils.AssumptionError:
            pass                               # this set cannot be rated: skip it

    def _get_best_subset(self, ritz):
        rated = {}
        current = frozenset()
        self._rate(ritz, current, rated)
        everything = set(range(len(ritz.values)))
        while True:
            proposals = self.subsets_generator.generate(ritz, everything.difference(current))
            if len(proposals) == 0:
                break
            round_ = {}
            for add in proposals:
                self._rate(ritz, current.union(add), round_)
            if round_:
                current = min(round_, key=round_.get)
            else:
                # nothing could be rated: extend by the proposal with the smallest residual norms
                sums = [numpy.sum(ritz.resnorms[list(add)]) for add in proposals]
                current = current.union(proposals[int(numpy.argmin(sums))])
            rated.update(round_)
        selection = list(min(rated, key=rated.get)) if rated else []
        self._report(ritz, selection, rated)
        return selection

    def _report(self, ritz, selection, rated):
        how = self.print_results
        if how is None:
            return
        if how == "number":
            print("# of selected deflation vectors: %d" % len(selection))
        elif how == "values":
            print("%d Ritz values corresponding to selected deflation vectors: %s"
                  % (len(selection), ", ".join(str(v) for v in ritz.values[selection])))
        elif how == "timings":
            print("Timings for all successfully evaluated choices of deflation vectors with "
                  "corresponding Ritz values:")
            for subset, time in sorted(rated.items(), key=lambda kv: kv[1]):
                print(" %ss: %s" % (time, ", ".join(str(v) for v in ritz.values[list(subset)])))
        else:
            raise utils.ArgumentError("Invalid value `%s` for argument `print_result`. Valid are `None`, "
                                      "`number`, `values` and `timings`." % how)


class RitzFactorySimple(_DeflationVectorFactory):
    """A fixed number of (harmonic) Ritz vectors chosen by a criterion on the Ritz values
    (krypy/recycling/factories.py:142-194).  The vectors stay in HBM (utils.DeviceBlock).
    ``realify`` (new, default True): complex Ritz vectors of nonsymmetric problems are replaced by a
    real basis of their span, because the device path is real."""

    def __init__(self, mode="ritz", n_vectors=0, which="sm", realify=True):
        self.mode = mode
        self.n_vectors = n_vectors
        self.which = which
        self.realify = realify

    def get(self, solver):
        ritz = deflation.Ritz(solver, mode=self.mode)
        values, k = ritz.values, self.n_vectors
        keys = {
            "lm": (numpy.abs(values), True), "sm": (numpy.abs(values), False),
            "lr": (numpy.real(values), True), "sr": (numpy.real(values), False),
            "li": (numpy.imag(values), True), "si": (numpy.imag(values), False),
            "smallest_res": (ritz.resnorms, False),
        }
        if self.which not in keys:
            raise utils.ArgumentError(
                "Invalid value '%s' for 'which'. Valid are lm, sm, lr, sr, li, si and smallest_res." % self.which)
        key, largest = keys[self.which]
        order = numpy.argsort(key)
        indices = order[-k:] if largest else order[:k]           # factories.py:174-187 (k == 0: see there)
        if k == 0 and largest:
            indices = order[-0:]
        return ritz.get_vectors_dev(indices, realify=self.realify)


class UnionFactory(_DeflationVectorFactory):
    """Concatenate the vectors of several factories (krypy/recycling/factories.py:197-208)."""

    def __init__(self, factories):
        self._factories = factories

    def get(self, solver):
        import torch
        blocks = []
        for f in self._factories:
            v = f.get(solver)
            blocks.append(v.block if isinstance(v, utils.DeviceBlock)
                          else utils._ctx().to_block(numpy.asarray(v), solver._td))
        return utils.DeviceBlock(torch.cat(blocks, dim=0))
