"""Gram-matrix builders (gpExp/gp_kernel_utilities.py) on the CUDA library:

    calculateCovarianceMatrix          :34-68    fused distance + kernel Gram (K1)
    calculateCovarianceMatrixFITC      :70-104   FITC covariance / precision through two Cholesky factors
    covTimesV                          :107-142  matrix-free Gram x vector (the operator of the Nystrom eigen-solver)
    calculateKernelBasisFunctionsMC    :144-194  eigsh (ARPACK, on the host as in the reference) over the device operator
"""
import numpy as np

from ._lib import check, lib
from .device import ptr


def _nugget_arg(nugget):
    """The reference accepts a float or an ndarray and dies with NameError on anything else
    (gp_kernel_utilities.py:62-67: `dadd` is never bound)."""
    if isinstance(nugget, float) or isinstance(nugget, np.ndarray):
        return nugget
    raise NameError("name 'dadd' is not defined")


def calculateCovarianceMatrix(kernel, points, nugget=0.0):
    """covarianceMatrix[j, i] = kernel(points[i], points[j]) + nugget on the diagonal."""
    nugget = _nugget_arg(nugget)
    size_of_mat, dim = points.shape
    dev = kernel._bind()
    P = dev.points(points)
    ld = P.ld
    out = dev.zeros(max(size_of_mat, 1), ld)
    nug_vec = dev.upload(nugget.astype(np.float64).ravel()) if isinstance(nugget, np.ndarray) else None
    # row j of the reference is kernel.evaluate(points, points[j]) = k(x_i, x_j) over i
    check(lib.gpx_gram(dev.h, ptr(P.X), size_of_mat, ld, ptr(P.X), size_of_mat, ld, ptr(out), ld, 1, ptr(nug_vec),
                       0.0 if nug_vec is not None else float(nugget), dev.stream), "gpx_gram")
    # gpx_gram writes out[i, j] = k(X[i], Y[j]); the reference's [j, i] = k(x_i, x_j) is its transpose
    return out[:size_of_mat, :size_of_mat].t().contiguous().cpu().numpy()


def calculateCovarianceMatrixFITC(kernel, nodes, nugget, fitc, returnCov=False):
    """FITC precision (and covariance) of `nodes` (gp_kernel_utilities.py:70-104).  `fitc` is either the fraction of nodes
    to draw as inducing points (np.random.permutation, as the reference) or the inducing points themselves."""
    from .engine import FitcFactor
    if isinstance(fitc, float):
        count = int(np.floor(len(nodes) * fitc))
        snodes = np.array(nodes[np.random.permutation(len(nodes))[0:count]], dtype=float)
    else:
        snodes = fitc
    dev = kernel._bind()
    f = FitcFactor(dev, dev.points(nodes), dev.points(np.asarray(snodes, dtype=np.float64)), _nugget_arg(nugget))
    precMat = f.precision()
    if returnCov is False:
        return (precMat, snodes) if fitc is not False else precMat
    covmat = f.covariance().T.copy()
    return (covmat, precMat, snodes) if fitc is not False else (covmat, precMat)


class _GramOperator:
    """v -> K(mcPoints, mcPoints) v without the matrix: coordinates stay resident, one fused kernel per product."""

    def __init__(self, kernel, mcPoints):
        self.kernel = kernel
        self.dev = kernel._bind()
        self.P = self.dev.points(mcPoints)
        self.ws = self.dev.zeros(max(int(lib.gpx_gram_matvec_workspace(self.P.n)), 1))
        self.out = self.dev.zeros(self.P.ld)

    def __call__(self, b):
        dev, P = self.dev, self.P
        self.kernel._bind(dev)
        bd = dev.upload(np.ascontiguousarray(b, dtype=np.float64).ravel())
        check(lib.gpx_gram_matvec(dev.h, ptr(P.X), P.n, P.ld, ptr(P.X), P.n, P.ld, ptr(bd), ptr(self.ws), ptr(self.out),
                                  dev.stream), "gpx_gram_matvec")
        return self.out[: P.n].cpu().numpy()


def covTimesV(b, kernel, mcPoints):
    """out[i] = sum_j kernel(mcPoints[j], mcPoints[i]) b[j]  (gp_kernel_utilities.py:107-142; the reference forks one
    process per core and evaluates one kernel row per Python iteration)."""
    return _GramOperator(kernel, mcPoints)(b).reshape(np.shape(b))


def calculateKernelBasisFunctionsMC(kernel, numBasis, mcPoints):
    """Leading eigenpairs of the kernel's integral operator by Monte Carlo + Nystrom (gp_kernel_utilities.py:144-194):
    eigsh over the matrix-free Gram operator, eigenvalues / nMC in descending order, eigenvectors * sqrt(nMC)."""
    from scipy.sparse.linalg import LinearOperator, eigsh
    nMC = mcPoints.shape[0]
    k = int(min(numBasis, nMC))
    op = _GramOperator(kernel, mcPoints)
    A = LinearOperator((nMC, nMC), matvec=op, dtype=float)
    eigv, eigve = eigsh(A, k=k, maxiter=10 * k)
    eigv, eigve = eigv[::-1], eigve[:, ::-1]
    return eigv / float(nMC), eigve * np.sqrt(float(nMC))
