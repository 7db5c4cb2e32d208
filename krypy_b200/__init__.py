"""krypy_b200 -- B200-native (sm_100a) Krylov solver engine behind the KryPy API.

Drop-in for the hot path of andrenarchy/krypy: ``krypy.linsys.{LinearSystem, Cg,
Minres, Gmres, RestartedGmres}``, ``krypy.deflation.Deflated{Cg, Minres, Gmres}``,
the ``krypy.utils`` operator / inner-product / Arnoldi / Projection surface and
the ``cg / minres / gmres`` convenience functions.  All N-sized arithmetic runs
in hand-written CUDA kernels (krypy_b200/csrc, C ABI in include/krypy_b200.h);
there is no CPU fallback.
"""
from . import deflation, linsys, problems, recycling, utils
from ._convenience import cg, gmres, minres

__version__ = "0.1.0"
__all__ = ["linsys", "deflation", "utils", "recycling", "problems", "cg", "minres", "gmres", "__version__"]
