"""Recycling solvers (krypy/recycling/linsys.py:7-136): a sequence of solves where each one is
deflated with vectors a factory extracts from the previous one.  The deflation space stays in HBM
between solves (``utils.DeviceBlock``)."""
import numpy

from .. import deflation, linsys, utils


def _wants_timings(factory):
    """does this factory (or a member of a UnionFactory) rate candidates by estimated run time?"""
    from . import factories
    if isinstance(factory, factories.RitzFactory):
        return True
    return any(_wants_timings(f) for f in getattr(factory, "_factories", ()))


def _named_factory(name):
    """the string shortcuts of krypy/recycling/linsys.py:76-88"""
    from . import evaluators, factories
    table = {
        "RitzApproxKrylov": lambda: evaluators.RitzApproxKrylov(),
        "RitzAprioriCg": lambda: evaluators.RitzApriori(Bound=utils.BoundCG),
        "RitzAprioriMinres": lambda: evaluators.RitzApriori(Bound=utils.BoundMinres),
    }
    if name not in table:
        raise utils.ArgumentError("unknown vector_factory '%s'" % name)
    return factories.RitzFactory(subset_evaluator=table[name]())


class _RecyclingSolver(object):
    """Keeps the last deflated solver (``last_solver``) and the wall-clock ``timings`` of the factory
    and solve phases (krypy/recycling/linsys.py:7-105).  Subclasses name the deflated solver class."""

    _Deflated = None

    def __init__(self, vector_factory=None):
        self._DeflatedSolver = self._Deflated
        self._vector_factory = vector_factory
        self.timings = utils.Timings()
        self.last_solver = None

    def solve(self, linear_system, vector_factory=None, *args, **kwargs):
        factory = self._vector_factory if vector_factory is None else vector_factory
        if isinstance(factory, str):
            factory = _named_factory(factory)
        # Evaluator-driven factories rate subsets by estimated TIME and need the operator timings of a
        # TimedLinearSystem (the reference always converts, recycling/linsys.py:69-70).  The wrapped
        # operators are the same objects, so nothing is uploaded again; factories that only read Ritz
        # pairs keep the system as it is.
        if _wants_timings(factory) and not isinstance(linear_system, linsys.TimedLinearSystem):
            linear_system = linsys.ConvertedTimedLinearSystem(linear_system)
        with self.timings["vector_factory"]:
            recycle = self.last_solver is not None and factory is not None
            U = factory.get(self.last_solver) if recycle else numpy.zeros((linear_system.N, 0))
        with self.timings["solve"]:
            self.last_solver = self._DeflatedSolver(linear_system, U=U, store_arnoldi=True, *args, **kwargs)
        return self.last_solver


class RecyclingCg(_RecyclingSolver):
    """krypy/recycling/linsys.py:106-114."""
    _Deflated = deflation.DeflatedCg


class RecyclingMinres(_RecyclingSolver):
    """krypy/recycling/linsys.py:117-125."""
    _Deflated = deflation.DeflatedMinres


class RecyclingGmres(_RecyclingSolver):
    """krypy/recycling/linsys.py:128-136."""
    _Deflated = deflation.DeflatedGmres
