"""Recycling Krylov solvers (krypy/recycling/*): strategy objects on host
scalars that pick deflation vectors from the last solve and run the deflated device solvers.

Implemented: ``Recycling{Cg,Minres,Gmres}`` and the factories that need only Ritz pairs
(``RitzFactorySimple``, ``UnionFactory``).  The evaluator-driven ``RitzFactory`` family depends on
``Arnoldifyer`` / ``bound_pseudo`` (out of scope) and raises."""
from . import factories
from .linsys import RecyclingCg, RecyclingGmres, RecyclingMinres

__all__ = ["RecyclingCg", "RecyclingMinres", "RecyclingGmres", "factories"]
