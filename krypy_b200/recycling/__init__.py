"""Recycling Krylov solvers (krypy/recycling/*): strategy objects on host
scalars that pick deflation vectors from the last solve and run the deflated device solvers.
``Recycling{Cg,Minres,Gmres}``, the factories (``RitzFactorySimple``, ``UnionFactory``, the
evaluator-driven ``RitzFactory``), subset ``generators`` and ``evaluators``."""
from . import evaluators, factories, generators
from .linsys import RecyclingCg, RecyclingGmres, RecyclingMinres

__all__ = ["RecyclingCg", "RecyclingMinres", "RecyclingGmres", "factories", "evaluators", "generators"]
