"""Zero-mean Gaussian process behind the reference's GP interface (gpExp/gp.py:49-468), backed by the CUDA library:
Gram (K1) + Cholesky (K2) replace `calculateCovarianceMatrix` + `np.linalg.pinv` (gp.py:176-181), the fused Gram+TRSM
(K1+K3) and column sums of squares (K4) replace the per-point `k^T P k` loops (gp.py:246-256, :133-145).

Two ways in, as for the kernels: the standalone `GP` class below, or -- with the reference importable --
`gpexp_b200.install_as_gpExp()`, which keeps the reference's own `GP` (constructor, hyper-parameter fitting, sampling)
and rebinds the methods listed in `DEVICE_METHODS` / `DEVICE_ATTRS` onto it.

On the device path: train, evaluate(compvar=0|1|2), addNodesAndComputeCovariance (dense and FITC), evaluateVariance,
evaluateVarianceDerivative (homo- and heteroscedastic), computeLogLike / loglikeParams (value and hyper-parameter gradient).
"""
import copy

import numpy as np

from . import _lib
from ._lib import GpxError, check, lib
from .device import ptr
from .engine import DesignFactor, FitcFactor
from .gp_kernel_utilities import _nugget_arg


# ---- lazily built device factor ------------------------------------------------------------------------------------------
# The Gram matrix and its factor are built on first use (a costFunctionGP_MI over a pool that only the column-sharded engine
# can hold must not force a dense |V| x |V| factor on one GPU at construction time).
def _get_factor(self):
    if self._factor_obj is None and self._pending is not None:
        nodes, nugget, inducing = self._pending
        dev = self.kernel._bind()
        if inducing is None:
            self._factor_obj = DesignFactor(dev, dev.points(nodes), nugget)
        else:
            self._factor_obj = FitcFactor(dev, dev.points(nodes), dev.points(inducing), nugget)
        self._pending = None
    return self._factor_obj


# covarianceMatrix / precisionMatrix are materialised on the host only when somebody reads them
def _get_cov(self):
    if self._cov_host is None and self._factor is not None:
        self._cov_host = self._factor.covariance().T.copy()
    return self._cov_host


def _set_cov(self, value):
    self._cov_host = value


def _get_prec(self):
    if self._prec_host is None and self._factor is not None:
        self._prec_host = self._factor.precision()
    return self._prec_host


def _set_prec(self, value):
    self._prec_host = value


def _copy(self):
    # costFunctionGP_IVAR shallow-copies the GP (experimentalDesign.py:64); the device factor is shared, not duplicated
    new = type(self).__new__(type(self))
    new.__dict__.update(self.__dict__)
    return new


def _fitc_inducing(self, nodes):
    """Inducing points of a FITC GP: a random subset of floor(n * FITC) nodes drawn with the global numpy generator on the
    first call and kept afterwards (gp.py:184-192; the reference's `self.fitcnodes == None` test only survives the first
    call under numpy >= 2, the kept subset is what it means)."""
    if self.fitcnodes is None:
        count = int(np.floor(len(nodes) * self.FITC))
        self.fitcnodes = np.array(nodes[np.random.permutation(len(nodes))[0:count]], dtype=float)
    return self.fitcnodes


def train(self, pts, evalsIn, noiseIn=None):
    """Compute the GP coefficients precision . evals (gp.py:76-101)."""
    assert len(evalsIn.shape) == 1, "evaluations must be an (N,) array for training GP"
    evals = evalsIn - self.gpPriorMean(pts)
    self.addNodesAndComputeCovariance(pts, noiseIn)
    self.fVals = evals.copy()
    self.coeff = self._factor.solve_vector(evals)


def addNodesAndComputeCovariance(self, nodes, noiseIn=None):
    """Gram + factor of the design (gp.py:156-211): dense Cholesky, or the FITC Woodbury precision when the GP was built
    with FITC=fraction (per-point noise is "NOT IMPLEMENTED YET" there in the reference as well, gp.py:209-210)."""
    inducing = None
    if self.FITC is not None:
        if noiseIn is not None:
            print("NOT IMPLEMENTED YET")
            self.pts = nodes.copy()
            return
        inducing = _fitc_inducing(self, nodes)
    nugget = _nugget_arg(self.noise if noiseIn is None else noiseIn)
    self.kernel._bind()  # fails loudly here if there is no device / library
    self._factor_obj = None
    self._pending = (nodes.copy(), nugget.copy() if isinstance(nugget, np.ndarray) else nugget, inducing)
    self._cov_host = None
    self._prec_host = None
    self.pts = nodes.copy()


def _require_factor(self):
    assert self.pts is not None
    if self._factor is None:
        raise GpxError("GP has no device factor; call addNodesAndComputeCovariance or train first")
    self.kernel._bind(self._factor.dev)
    return self._factor


def evaluateVariance(self, newpt, parallel=1):
    """Posterior variance k(x,x) - k^T P k, raw (not abs'd), as gp.py:213-259.  `parallel` is accepted for compatibility;
    nothing is forked (the reference's fork path, gp.py:257-258, must never run after CUDA initialisation)."""
    assert self.pts is not None
    assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
    f = _require_factor(self)
    X = f.dev.points(newpt)
    _, var = f.solve_gram(X)
    return var[: X.n].cpu().numpy()


def evaluateVarianceDerivative(self, newpt, noiseFunc=None):
    """Posterior-variance derivative with respect to the training points (gp.py:282-341):
    out[k*l, jj] = dC(newpt[jj], newpt[jj]) / d self.pts[k, l], shape (len(pts)*dim, len(newpt)).  Squared-exponential
    kernels only (the reference has no N-D derivative for the other families).

    noiseFunc (heteroscedastic branch, gp.py:314-318): an object with `.deriv(points) -> (n, d)`; its derivative enters the
    design-Gram derivative at coincident design points.  The reference's second adjustment (:317-319) fires only when EVERY
    evaluation point coincides with a design point (`np.linalg.norm(p - newpt)` is the norm of the whole difference
    matrix); that degenerate query is not supported here."""
    assert self.pts is not None, "must specify training points before running this"
    assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
    self.kernel._require_derivative()
    f = _require_factor(self)
    if not isinstance(f, DesignFactor):
        raise NotImplementedError("variance derivatives of a FITC GP are not on the device path")
    noise_grad = same = None
    if noiseFunc is not None:
        pts = self.pts
        if any(np.linalg.norm(pts[zz:zz + 1] - newpt) < 1e-10 for zz in range(len(pts))):
            raise NotImplementedError("heteroscedastic variance derivative with all evaluation points on a design point")
        noise_grad = np.asarray(noiseFunc.deriv(pts), dtype=np.float64).reshape(len(pts), self.kernel.dimension)
        same = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=2) < 1e-10
    X = f.dev.points(newpt)
    out = f.variance_gradient(X, noise_grad, same)
    return out[: f.n * self.kernel.dimension, : X.n].cpu().numpy()


def computeLogLike(self, pts, evals):
    """Marginal log-likelihood of (pts, evals) under the current hyper-parameters (gp.py:373-392)."""
    return self.loglikeParams(pts, evals)


def loglikeParams(self, pts, evals, returnDeriv=0, noiseIn=None):
    """-1/2 y^T K^-1 y - 1/2 log|K| - n/2 log 2 pi with K = Gram + noise (gp.py:394-446) through the device factor
    (dense Cholesky or FITC), and with returnDeriv=1 the gradient dict over the kernel's hyper-parameters + 'noise'
    (gp.py:447-468):  1/2 tr((alpha alpha^T - K^-1) dK/dtheta), the 'noise' entry scaled by 2*noise as the reference does.

    The gradient needs `kernel.derivativeWrtHypParams`, which only the squared-exponential kernel has (kernels.py:125-144);
    the reference's own version indexes with a float there and raises IndexError under numpy >= 1.12, so parity of the
    gradient is defined against the analytic expression (oracle.fast_loglike_gradient), not against golden vectors."""
    dev = self.kernel._bind()
    if noiseIn is not None and self.FITC is None:
        nugget = _nugget_arg(noiseIn)
    else:
        nugget = _nugget_arg(self.noise)
    if self.FITC is not None and noiseIn is None:
        f = FitcFactor(dev, dev.points(pts), dev.points(_fitc_inducing(self, pts)), nugget)
        first = -0.5 * f.quad_form(evals)
    else:
        f = DesignFactor(dev, dev.points(pts), nugget)
        first = -0.5 * f.whitened_norm2(evals)
    second = -0.5 * f.logdet()
    third = -len(evals) / 2.0 * np.log(2.0 * np.pi)
    out = first + second + third
    if returnDeriv != 1:
        return out
    if not isinstance(f, DesignFactor):
        raise NotImplementedError("log-likelihood gradient of a FITC GP is not on the device path")
    if self.kernel._gpx_spec()[0] != _lib.SE:
        raise AttributeError("derivativeWrtHypParams not implemented for %s" % type(self.kernel).__name__)
    g = f.loglike_gradient(evals)                       # cl_0 .. cl_{d-1}, signalSize, noise
    d = self.kernel.dimension
    by_name = {'cl%d' % i: g[i] for i in range(d)}
    by_name['signalSize'] = g[d]
    outD = {key: float(by_name[key]) for key in self.kernel.hyperParam.keys()}
    outD['noise'] = float(g[d + 1]) * self.noise * 2.0   # gp.py:463-464
    return out, outD


def evaluate(self, newpt, compvar=0):
    """Posterior mean, and variance (compvar=1, abs'd as gp.py:145) or covariance (compvar=2)."""
    assert newpt.shape[1] == self.kernel.dimension, "evaluation points for GP is incorrect shape"
    f = _require_factor(self)
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
        Kqq = dev.zeros(max(q, 1), X.ld)
        check(lib.gpx_gram(dev.h, ptr(X.X), q, X.ld, ptr(X.X), q, X.ld, ptr(Kqq), X.ld, 0, None, 0.0, dev.stream), "gpx_gram")
        if isinstance(f, DesignFactor):
            # covar[i, j] = k(x_i, x_j) - W[:, i] . W[:, j]  (gp.py:147-152)
            W, _ = f.solve_gram(X, want_var=False)
            check(lib.gpx_dgemm_tn_sub(dev.h, ptr(W), X.ld, ptr(W), X.ld, ptr(Kqq), X.ld, q, q, n, 0, dev.stream),
                  "gpx_dgemm_tn_sub")
        else:
            Z = f.apply_precision(Kx, q, X.ld)
            check(lib.gpx_dgemm_tn_sub(dev.h, ptr(Kx), X.ld, ptr(Z), X.ld, ptr(Kqq), X.ld, q, q, n, 0, dev.stream),
                  "gpx_dgemm_tn_sub")
        return out, Kqq[:q, :q].cpu().numpy()
    else:
        return out


# what install_as_gpExp() rebinds on the reference's own GP class: methods, and class-level attributes / properties
DEVICE_METHODS = {
    "train": train, "evaluate": evaluate, "addNodesAndComputeCovariance": addNodesAndComputeCovariance,
    "evaluateVariance": evaluateVariance, "evaluateVarianceDerivative": evaluateVarianceDerivative,
    "computeLogLike": computeLogLike, "loglikeParams": loglikeParams, "__copy__": _copy,
}
DEVICE_ATTRS = {
    "_factor_obj": None, "_pending": None, "_cov_host": None, "_prec_host": None,
    "_factor": property(_get_factor),
    "covarianceMatrix": property(_get_cov, _set_cov),
    "precisionMatrix": property(_get_prec, _set_prec),
}


class GP:
    """GP with zero prior mean: GP(kernel, noise[, FITC=fraction of the nodes used as inducing points])."""

    coeff = None
    noise = None
    pts = None
    FITC = None
    fitcnodes = None

    def __init__(self, kernel_in, noiseIn, **kwargs):
        try:
            self.kernel = copy.deepcopy(kernel_in)   # the GP owns its kernel (gp.py:62-66)
        except Exception:
            print("warning ")
            self.kernel = copy.copy(kernel_in)
        self.noise = noiseIn
        self.FITC = kwargs.get('FITC', None)

    def gpPriorMean(self, pts):
        return np.zeros((pts.shape[0]))

    def getHypParamNames(self):
        return self.kernel.hyperParam.keys()

    def updateKernelParams(self, paramsIn):
        """New hyper-parameters; a 'noise' entry goes to the GP, the rest to the kernel (gp.py:474-497)."""
        params = dict(paramsIn)
        if 'noise' in params:
            self.noise = copy.copy(params.pop('noise'))
        self.kernel.updateHyperParameters(params)


for _name, _obj in {**DEVICE_ATTRS, **DEVICE_METHODS}.items():
    setattr(GP, _name, _obj)
