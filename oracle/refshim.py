"""Import the UNMODIFIED reference (andrenarchy/krypy at /root/reference) in the
build container -- TEST INFRASTRUCTURE ONLY, never imported by krypy_b200.

The reference uses names removed from numpy >= 2 / scipy >= 1.12
(krypy/utils.py:19,122,...); the aliases below restore them before import, the
reference's files stay untouched.  /root/reference does
not exist on the GPU box: only ``oracle/make_golden.py`` and the optional
cross-check tests (skipped when the directory is absent) use this module.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KRYPY_REFERENCE_ROOT", "/root/reference")


VENDORED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "krypy"))


def vendored_available():
    """the unmodified reference installed by pip into baseline/_ref (travels to the GPU box)"""
    return os.path.isdir(os.path.join(VENDORED_ROOT, "krypy"))


def import_reference(root=None):
    import numpy
    import scipy.sparse
    import scipy.sparse._sputils as _sputils

    numpy.complex = complex
    numpy.float = float
    numpy.int = int
    numpy.Inf = numpy.Infinity = numpy.inf

    def _find_common_type(array_types, scalar_types):
        d = [t for t in list(array_types) + list(scalar_types) if t is not None]
        return numpy.result_type(*d) if d else numpy.dtype(None)

    numpy.find_common_type = _find_common_type
    try:
        import scipy.sparse.sputils as sputils
    except Exception:
        sputils = types.ModuleType("scipy.sparse.sputils")
        sys.modules["scipy.sparse.sputils"] = sputils
        scipy.sparse.sputils = sputils
    sputils.isintlike = _sputils.isintlike
    root = REFERENCE_ROOT if root is None else root
    if root not in sys.path:
        sys.path.insert(0, root)
    import krypy

    return krypy
