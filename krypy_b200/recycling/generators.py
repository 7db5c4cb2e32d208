"""Generators of candidate subsets of Ritz pairs for deflation (krypy/recycling/generators.py).
Host logic on the n+d Ritz values."""
import numpy


class _RitzSubsetsGenerator(object):
    """Abstract base (krypy/recycling/generators.py:4-10)."""

    def generate(self, ritz, remaining_subset):
        raise NotImplementedError("abstract base class cannot be instanciated")


def _exhausted(ritz, remaining, max_vectors):
    return len(remaining) <= 1 or len(ritz.values) - len(remaining) >= max_vectors


class RitzSmall(_RitzSubsetsGenerator):
    """One candidate per round: the remaining Ritz value of smallest magnitude
    (krypy/recycling/generators.py:13-24)."""

    def __init__(self, max_vectors=numpy.inf):
        self.max_vectors = max_vectors

    def generate(self, ritz, remaining_subset):
        rest = list(remaining_subset)
        if _exhausted(ritz, rest, self.max_vectors):
            return []
        return [{rest[int(numpy.argmin(numpy.abs(ritz.values[rest])))]}]


class RitzExtremal(_RitzSubsetsGenerator):
    """Candidates are the extremal remaining Ritz values: for self-adjoint problems the smallest and
    largest negative and positive ones, otherwise those of smallest and largest magnitude
    (krypy/recycling/generators.py:27-74)."""

    def __init__(self, max_vectors=numpy.inf):
        self.max_vectors = max_vectors

    def generate(self, ritz, remaining_subset):
        rest = numpy.array(list(remaining_subset))
        if _exhausted(ritz, rest, self.max_vectors):
            return []
        vals = ritz.values[rest]

        def ends(idx, key):
            """positions (into rest) of the smallest and largest key among idx"""
            if len(idx) == 0:
                return []
            order = idx[numpy.argsort(key[idx])]
            return [order[0]] if len(order) == 1 else [order[0], order[-1]]

        if ritz._deflated_solver.linear_system.self_adjoint:
            picked = ends(numpy.where(vals < 0)[0], vals) + ends(numpy.where(vals > 0)[0], vals)
        else:
            picked = ends(numpy.arange(len(vals)), numpy.abs(vals))
        return [{int(rest[i])} for i in picked]
