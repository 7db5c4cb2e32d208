"""Recycling solvers (krypy/recycling/linsys.py:7-136)."""
import numpy

from .. import deflation, linsys, utils


class _RecyclingSolver(object):
    """Base class: keeps the last deflated solver and asks a vector factory for the next
    deflation space (krypy/recycling/linsys.py:7-105)."""

    def __init__(self, DeflatedSolver, vector_factory=None):
        self._DeflatedSolver = DeflatedSolver
        self._vector_factory = vector_factory
        self.timings = utils.Timings()
        self.last_solver = None

    def solve(self, linear_system, vector_factory=None, *args, **kwargs):
        # the reference wraps the system in a TimedLinearSystem for its evaluators
        # (recycling/linsys.py:69-70); the factories implemented here never read the timings, so the
        # system is used as it is and its device-resident operators are not re-uploaded
        with self.timings["vector_factory"]:
            if vector_factory is None:
                vector_factory = self._vector_factory
            if isinstance(vector_factory, str):
                raise NotImplementedError(
                    "vector_factory='%s' needs the evaluator-driven RitzFactory (out of scope); pass a "
                    "RitzFactorySimple instance" % vector_factory)
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
