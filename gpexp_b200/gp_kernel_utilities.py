"""Gram-matrix builder (mirrors gpExp/gp_kernel_utilities.py:34-68) on the fused CUDA Gram kernel (K1).
The FITC / Nystrom helpers of the reference file are approximation features off the greedy path."""
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
