"""Deflation-vector factories (krypy/recycling/factories.py)."""
import numpy

from .. import deflation, utils


class _DeflationVectorFactory(object):
    """Abstract base class for selectors (krypy/recycling/factories.py:9-17)."""

    def get(self, solver):
        raise NotImplementedError("abstract base class cannot be instanciated")


class RitzFactory(_DeflationVectorFactory):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "RitzFactory needs the subset evaluators built on Arnoldifyer / bound_pseudo, which are "
            "outside the hot-path scope (SURVEY section 2); use RitzFactorySimple")


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
