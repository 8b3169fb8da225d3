"""Zero-mean Gaussian process with the reference's GP interface (gpExp/gp.py:49-259), backed by the
CUDA library: Gram (K1) + Cholesky (K2) replace `calculateCovarianceMatrix` + `np.linalg.pinv`
(gp.py:176-181), the fused Gram+TRSM (K1+K3) and column sums of squares (K4) replace the per-point
`k^T P k` loops (gp.py:246-256, :133-145).

Kept: GP(kernel, noise), train, evaluate(newpt, compvar=0|1|2), addNodesAndComputeCovariance,
evaluateVariance, attributes kernel / noise / pts / coeff / covarianceMatrix / precisionMatrix.
Out of scope: FITC, variance derivatives, sampling, marginal likelihood and
hyper-parameter optimisation.
"""
import copy

import numpy as np

from . import _lib
from ._lib import GpxError, check, lib
from .device import Device, ptr
from .engine import DesignFactor
from .gp_kernel_utilities import _nugget_arg


class GP:
    """GP with zero prior mean."""

    coeff = None
    noise = None
    pts = None
    FITC = None
    fitcnodes = None

    def __init__(self, kernel_in, noiseIn, **kwargs):
        try:
            self.kernel = copy.deepcopy(kernel_in)
        except Exception:
            print("warning ")
            self.kernel = copy.copy(kernel_in)
        self.noise = noiseIn
        if 'FITC' in kwargs and kwargs['FITC'] is not None:
            raise NotImplementedError("FITC sparse GPs are outside the B200 hot path (gp.py:182-208)")
        self._factor_obj = None
        self._pending = None
        self._cov_host = None
        self._prec_host = None

    # The Gram matrix and its Cholesky factor are built on first use (a costFunctionGP_MI over a pool that only the
    # column-sharded engine can hold must not force a dense |V| x |V| factor on one GPU at construction time).
    @property
    def _factor(self):
        if self._factor_obj is None and self._pending is not None:
            nodes, nugget = self._pending
            dev = self.kernel._bind()
            self._factor_obj = DesignFactor(dev, dev.points(nodes), nugget)
            self._pending = None
        return self._factor_obj

    # covarianceMatrix / precisionMatrix are materialised on the host only when somebody reads them
    @property
    def covarianceMatrix(self):
        if self._cov_host is None and self._factor is not None:
            self._cov_host = self._factor.covariance().T.copy()
        return self._cov_host

    @covarianceMatrix.setter
    def covarianceMatrix(self, value):
        self._cov_host = value

    @property
    def precisionMatrix(self):
        if self._prec_host is None and self._factor is not None:
            self._prec_host = self._factor.precision()
        return self._prec_host

    @precisionMatrix.setter
    def precisionMatrix(self, value):
        self._prec_host = value

    def __copy__(self):
        # costFunctionGP_IVAR shallow-copies the GP (experimentalDesign.py:64)
        new = GP.__new__(GP)
        new.__dict__.update(self.__dict__)
        return new

    def gpPriorMean(self, pts):
        return np.zeros((pts.shape[0]))

    def train(self, pts, evalsIn, noiseIn=None):
        """Compute the GP coefficients (gp.py:76-101)."""
        assert len(evalsIn.shape) == 1, "evaluations must be an (N,) array for training GP"
        evals = evalsIn - self.gpPriorMean(pts)
        self.addNodesAndComputeCovariance(pts, noiseIn)
        self.fVals = evals.copy()
        self.coeff = self._factor.solve_vector(evals)

    def addNodesAndComputeCovariance(self, nodes, noiseIn=None):
        """Gram + Cholesky of the design (gp.py:156-211, non-FITC branch)."""
        nugget = _nugget_arg(self.noise if noiseIn is None else noiseIn)
        self.kernel._bind()  # fails loudly here if there is no device / library
        self._factor_obj = None
        self._pending = (nodes.copy(), nugget.copy() if isinstance(nugget, np.ndarray) else nugget)
        self._cov_host = None
        self._prec_host = None
        self.pts = nodes.copy()

    def _require_factor(self):
        assert self.pts is not None
        if self._factor is None:
            raise GpxError("GP has no device factor; call addNodesAndComputeCovariance or train first")
        self.kernel._bind(self._factor.dev)
        return self._factor

    def _variance_device(self, query):
        f = self._require_factor()
        X = f.dev.points(query)
        W, var = f.solve_gram(X)
        return f, X, W, var

    def evaluateVariance(self, newpt, parallel=1):
        """Posterior variance k(x,x) - |U^-T k(D,x)|^2, raw (not abs'd), as gp.py:213-259.
        `parallel` is accepted for compatibility; nothing is forked (the reference's fork path,
        gp.py:257-258, must never run after CUDA initialisation)."""
        assert self.pts is not None
        assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
        _, X, _, var = self._variance_device(newpt)
        return var[: X.n].cpu().numpy()

    def evaluateVarianceDerivative(self, newpt, noiseFunc=None):
        """Posterior-variance derivative with respect to the training points (gp.py:282-341):
        out[k*l, jj] = dC(newpt[jj], newpt[jj]) / d self.pts[k, l], shape (len(pts)*dim, len(newpt)).
        Squared-exponential kernels only (the reference has no ND derivative for the other families)."""
        assert self.pts is not None, "must specify training points before running this"
        assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
        if noiseFunc is not None:
            raise NotImplementedError("the heteroscedastic variance derivative is not on the device path")
        self.kernel._require_derivative()
        f = self._require_factor()
        X = f.dev.points(newpt)
        out = f.variance_gradient(X)
        return out[: f.n * self.kernel.dimension, : X.n].cpu().numpy()

    def computeLogLike(self, pts, evals):
        """Marginal log-likelihood of (pts, evals) under the current hyper-parameters (gp.py:373-392)."""
        return self.loglikeParams(pts, evals)

    def loglikeParams(self, pts, evals, returnDeriv=0, noiseIn=None):
        """-1/2 y^T K^-1 y - 1/2 log|K| - n/2 log 2 pi with K = Gram + noise (gp.py:394-446), through the device
        Cholesky factor.  The hyper-parameter gradient (returnDeriv=1, gp.py:447-468) is not on the device path."""
        if returnDeriv:
            raise NotImplementedError("hyper-parameter gradients of the log-likelihood are outside the B200 path")
        nugget = _nugget_arg(self.noise if noiseIn is None else noiseIn)
        dev = self.kernel._bind()
        f = DesignFactor(dev, dev.points(pts), nugget)
        first = -0.5 * f.whitened_norm2(evals)
        second = -0.5 * f.logdet()
        third = -len(evals) / 2.0 * np.log(2.0 * np.pi)
        return first + second + third

    def getHypParamNames(self):
        return self.kernel.hyperParam.keys()

    def updateKernelParams(self, paramsIn):
        """Set new hyper-parameters; a 'noise' entry updates the GP noise (gp.py:474-497)."""
        params = copy.copy(paramsIn)
        if 'noise' in params.keys():
            self.noise = copy.copy(params['noise'])
            del params['noise']
        self.kernel.updateHyperParameters(params)

    def evaluate(self, newpt, compvar=0):
        """Posterior mean, and variance (compvar=1, abs'd as gp.py:145) or covariance (compvar=2)."""
        assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
        f = self._require_factor()
        dev = f.dev
        X = dev.points(newpt)
        # mean = k(x, D) . coeff : materialise the q x n cross Gram on the device (K1) and contract there
        q, n = X.n, f.n
        Kx = dev.zeros(max(n, 1), X.ld)
        check(lib.gpx_gram(dev.h, ptr(f.design.X), n, f.design.ld, ptr(X.X), q, X.ld, ptr(Kx), X.ld, 0, None, 0.0,
                           dev.stream), "gpx_gram")
        # mean[j] = sum_k coeff[k] Kx[k, j]  as  C -= (-coeff)^T Kx  on the DMMA routine
        neg = np.zeros((max(n, 1), 2))
        neg[:n, 0] = -np.asarray(self.coeff, dtype=np.float64)
        mean = dev.zeros(2, X.ld)
        check(lib.gpx_dgemm_tn_sub(dev.h, ptr(dev.upload(neg)), 2, ptr(Kx), X.ld, ptr(mean), X.ld, 1, q, n, 0, dev.stream),
              "gpx_dgemm_tn_sub")
        out = mean[0, :q].cpu().numpy() + self.gpPriorMean(newpt)
        if compvar == 1:
            _, var = f.solve_gram(X)
            return out, np.abs(var[:q].cpu().numpy())
        elif compvar == 2:
            W, _ = f.solve_gram(X, want_var=False)
            Kqq = dev.zeros(max(q, 1), X.ld)
            check(lib.gpx_gram(dev.h, ptr(X.X), q, X.ld, ptr(X.X), q, X.ld, ptr(Kqq), X.ld, 0, None, 0.0, dev.stream),
                  "gpx_gram")
            # covar[i, j] = k(x_i, x_j) - W[:, i] . W[:, j]  (gp.py:147-152)
            check(lib.gpx_dgemm_tn_sub(dev.h, ptr(W), X.ld, ptr(W), X.ld, ptr(Kqq), X.ld, q, q, n, 0, dev.stream),
                  "gpx_dgemm_tn_sub")
            return out, Kqq[:q, :q].cpu().numpy()
        else:
            return out
