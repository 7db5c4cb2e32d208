"""Recycling solvers (krypy/recycling/linsys.py:7-136)."""
import numpy

from .. import deflation, linsys, utils


def _needs_timings(factory):
    from . import factories
    if isinstance(factory, factories.RitzFactory):
        return True
    return any(_needs_timings(f) for f in getattr(factory, "_factories", ()))


class _RecyclingSolver(object):
    """Base class: keeps the last deflated solver and asks a vector factory for the next
    deflation space (krypy/recycling/linsys.py:7-105)."""

    def __init__(self, DeflatedSolver, vector_factory=None):
        self._DeflatedSolver = DeflatedSolver
        self._vector_factory = vector_factory
        self.timings = utils.Timings()
        self.last_solver = None

    def solve(self, linear_system, vector_factory=None, *args, **kwargs):
        from . import evaluators, factories
        if vector_factory is None:
            vector_factory = self._vector_factory
        shortcuts = {                                            # recycling/linsys.py:76-88
            "RitzApproxKrylov": lambda: evaluators.RitzApproxKrylov(),
            "RitzAprioriCg": lambda: evaluators.RitzApriori(Bound=utils.BoundCG),
            "RitzAprioriMinres": lambda: evaluators.RitzApriori(Bound=utils.BoundMinres),
        }
        if isinstance(vector_factory, str):
            if vector_factory not in shortcuts:
                raise utils.ArgumentError("unknown vector_factory '%s'" % vector_factory)
            vector_factory = factories.RitzFactory(subset_evaluator=shortcuts[vector_factory]())
        # The evaluators rate subsets by estimated TIME and need the operator timings of a
        # TimedLinearSystem (recycling/linsys.py:69-70).  The wrapped operators are the same objects,
        # so nothing is uploaded again.  Factories that only read Ritz pairs keep the system as it is.
        if _needs_timings(vector_factory) and not isinstance(linear_system, linsys.TimedLinearSystem):
            linear_system = linsys.ConvertedTimedLinearSystem(linear_system)
        with self.timings["vector_factory"]:
            if self.last_solver is None or vector_factory is None:
                U = numpy.zeros((linear_system.N, 0))
            else:
                U = vector_factory.get(self.last_solver)
        with self.timings["solve"]:
            self.last_solver = self._DeflatedSolver(linear_system, U=U, store_arnoldi=True, *args, **kwargs)
        return self.last_solver


class RecyclingCg(_RecyclingSolver):
    def __init__(self, *args, **kwargs):
        super(RecyclingCg, self).__init__(deflation.DeflatedCg, *args, **kwargs)


class RecyclingMinres(_RecyclingSolver):
    def __init__(self, *args, **kwargs):
        super(RecyclingMinres, self).__init__(deflation.DeflatedMinres, *args, **kwargs)


class RecyclingGmres(_RecyclingSolver):
    def __init__(self, *args, **kwargs):
        super(RecyclingGmres, self).__init__(deflation.DeflatedGmres, *args, **kwargs)
