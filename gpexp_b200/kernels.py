"""Covariance functions with the reference's class names, constructors and hyperParam keys
(gpExp/kernels.py), evaluated by the CUDA library.

`Kernel.evaluate(x1, x2)` keeps the reference's *pairwise* semantics (kernels.py:49-65): two (n,d)
arrays, or one of them (1,d), give an (n,) vector.  The derivative methods of the reference
(kernels.py:125-181, :295-324) belong to the continuous optimisers and are out of scope.
"""
import math

import numpy as np

from . import _lib
from ._lib import check, lib
from .device import Device, ptr


class Kernel(object):
    """Base class (kernels.py:30-70)."""

    nugget = 0.0
    hyperParam = dict({})

    def __init__(self, hyperParam, dimension, *argc):
        self.dimension = dimension
        self.hyperParam = hyperParam
        super(Kernel, self).__init__()

    def updateHyperParameters(self, hyperParamNew):
        for key in hyperParamNew.keys():
            assert key in self.hyperParam.keys(), (key, " is not a valid hyperParameter")
        self.hyperParam = hyperParamNew

    # ---- device side -----------------------------------------------------------------------------
    def _gpx_spec(self):
        """(family, d, flat parameter vector) in the order gpx_set_kernel expects."""
        raise AttributeError("this kernel has no CUDA implementation")

    def _bind(self, dev=None) -> Device:
        dev = dev or Device.get()
        fam, d, params = self._gpx_spec()
        dev.set_kernel(fam, d, params)
        return dev

    def _require_derivative(self):
        """Only KernelSquaredExponential has an N-D `derivative` in the reference."""
        raise AttributeError("derivative of %s not implemented" % type(self).__name__)

    def evaluate(self, x1, x2):
        assert len(x2.shape) > 1 and len(x1.shape) > 1, "Must supply nd arrays to evaluation function"
        nPointsx1 = x1.shape[0]
        nPointsx2 = x2.shape[0]
        assert x1.shape[1] == self.dimension and x2.shape[1] == self.dimension, \
            (" Incorrect dimension of input points fed to kernel ", x1.shape, x2.shape)
        return self.evaluateF(x1, x2)

    def evaluateF(self, x1, x2):
        n1, n2 = x1.shape[0], x2.shape[0]
        # equal lengths, or a single point on one side (what np.tile makes of it in the reference);
        # any other mix fails the reference's shape assert as well
        assert n1 == n2 or n1 == 1 or n2 == 1, "__evaluate() received non-equal shaped point sets"
        dev = self._bind()
        a, b = dev.points(x1), dev.points(x2)
        n = max(n1, n2) if min(n1, n2) > 0 else 0
        out = dev.zeros(max(n, 1))
        check(lib.gpx_kernel_pairwise(dev.h, ptr(a.X), n1, a.ld, ptr(b.X), n2, b.ld, ptr(out), dev.stream),
              "gpx_kernel_pairwise")
        return out[:n].cpu().numpy()


class KernelIsoMatern(Kernel):
    """Isotropic Matern, nu = 3/2 (kernels.py:72-98)."""

    def __init__(self, rho, signalSize, dimension, nu=3.0 / 2.0):
        hyperParam = dict({'rho': rho, 'signalSize': signalSize})
        self.nu = nu
        super(KernelIsoMatern, self).__init__(hyperParam, dimension)

    def _gpx_spec(self):
        if np.abs(1.5 - self.nu) >= 1e-10:  # the reference leaves `out` unbound here (kernels.py:85-91)
            raise UnboundLocalError("KernelIsoMatern is implemented for nu = 3/2 only")
        return _lib.MATERN32, self.dimension, [self.hyperParam['rho'], self.hyperParam['signalSize']]

    def derivativeWrtHypParams(self, x1, x2):
        raise AttributeError("derivativeWrtHypParams not implemented for KernelIsoMatern")


class KernelSquaredExponential(Kernel):
    """exp(-(x-x')^2 / (2 l^2)), isotropic (one length) or ARD (d lengths) (kernels.py:100-123)."""

    def __init__(self, correlationLength, signalSize, dimension):
        hyperParam = dict({})
        if len(correlationLength) == 1:
            correlationLength = np.tile(correlationLength, (dimension))
        for ii in range(len(correlationLength)):
            hyperParam['cl' + str(ii)] = correlationLength[ii]
        hyperParam['signalSize'] = signalSize
        super(KernelSquaredExponential, self).__init__(hyperParam, dimension)

    def _gpx_spec(self):
        cl = [self.hyperParam['cl' + str(ii)] for ii in range(self.dimension)]
        return _lib.SE, self.dimension, cl + [self.hyperParam['signalSize']]

    def _require_derivative(self):
        return True

    def derivative(self, x1, x2, version=0):
        """out[jj, ii] = dK(x1[jj], x2)/dx1[jj, ii] as the reference computes it (kernels.py:146-181):
        -signalSize * (x1 - x2)/cl^2 * evaluate(x1, x2).  x2 is a single (1, d) point."""
        assert len(x2.shape) > 1 and len(x1.shape) > 1, "Must supply nd arrays to evaluation function"
        assert x2.shape[0] == 1 and x2.shape[1] == self.dimension, "x2 not in correct shape"
        assert x1.shape[0] > 0 and x1.shape[1] == self.dimension, "x1 not in correct shape"
        dev = self._bind()
        a, b = dev.points(x1), dev.points(x2)
        n, d = x1.shape
        ld = max(n * d, 1)
        out = dev.zeros(1, ld)
        check(lib.gpx_se_dgram(dev.h, ptr(b.X), 1, b.ld, ptr(a.X), n, a.ld, ptr(out), ld, dev.stream), "gpx_se_dgram")
        return out[0, : n * d].cpu().numpy().reshape(n, d)


class KernelMehlerND(Kernel):
    """Product of 1-D Mehler kernels (kernels.py:183-228)."""

    def __init__(self, tIn, dimension):
        hyperParam = dict({})
        self.oneDKern = []
        for ii in range(dimension):
            hyperParam[ii] = tIn[ii]
            self.oneDKern.append(KernelMehler1D(tIn[ii], 1))
        super(KernelMehlerND, self).__init__(hyperParam, dimension)

    def updateHyperParameters(self, params):
        for keys in self.hyperParam.keys():
            self.hyperParam[keys] = params[keys]
        for ii in range(self.dimension):
            self.oneDKern[ii].updateHyperParameters(dict({'t': self.hyperParam[ii]}))

    def _gpx_spec(self):
        return _lib.MEHLER, self.dimension, [self.hyperParam[ii] for ii in range(self.dimension)]

    def derivative(self, x1, x2):
        raise AttributeError("derivative of KernelMehlerND not yet implemented")


class KernelMehler1D(Kernel):
    """1-D Mehler (Hermite) kernel (kernels.py:250-293)."""

    def __init__(self, tIn, dimension):
        assert dimension == 1, "Mehler Hermite Kernel is only one dimensional"
        hyperParam = dict({})
        hyperParam['t'] = tIn
        super(KernelMehler1D, self).__init__(hyperParam, dimension)

    def _gpx_spec(self):
        return _lib.MEHLER, 1, [self.hyperParam['t']]

    def evaluateF(self, x1, x2):
        assert x1.shape[1] == 1 and x2.shape[1] == 1, "Hermite1d kernel only accepts one dimensional points"
        out = super(KernelMehler1D, self).evaluateF(x1, x2)
        if out.size and math.isnan(out[0]):  # kernels.py:288-292
            print("xs ", x1[0, :], x2[0, :])
            print("t", self.hyperParam['t'])
            print('NAN in kernel hermi1d exiting')
            raise SystemExit
        return out
