"""Cost functions and discrete greedy design drivers with the reference's names and signatures
(gpExp/experimentalDesign.py), running on the device-resident engines of gpexp_b200.engine.

    costFunctionGP_IVAR                     experimentalDesign.py:60-117   (version 1, MC-integrated variance)
    costFunctionGP_MI                       experimentalDesign.py:223-285
    performGreedyVarExperimentalDesign      experimentalDesign.py:787-845
    performGreedyMIExperimentalDesign       experimentalDesign.py:753-785
    performGreedyIVARExperimentalDesign     NEW: the discrete greedy-IVAR driver the north star asks for; the
                                            reference only has the cost function

The continuous optimisers, the eigen-basis IVAR (version 0, dead code in the reference), the clustering
design and the Bayesian-optimisation costs are out of scope.
"""
import copy
import itertools

import numpy as np

from . import _lib
from ._lib import check, lib
from .device import Device, ptr
from .engine import (DesignFactor, GreedyIVAREngine, GreedyMIEngine, GreedyVarEngine, Shard, ShardedMIEngine,
                     prior_scale)
from .gp_kernel_utilities import _nugget_arg

VERBOSE = True  # the reference prints its progress unconditionally (experimentalDesign.py:812-813)


class costFunctionBase(object):

    def __init__(self, nInputs, space):
        self.numInputs = nInputs
        self.space = space


class costFunctionGP_IVAR(costFunctionBase):
    """Integrated posterior variance of a design, Monte-Carlo version (version=1)."""

    def __init__(self, gaussianProcess, nInputs, space, version=1, **kwargs):
        super(costFunctionGP_IVAR, self).__init__(nInputs, space)
        self.gaussianProcess = copy.copy(gaussianProcess)
        self.version = version
        if self.version == 1:
            if 'mcPoints' in kwargs:
                self.mcPoints = kwargs['mcPoints']
                self.nMC = len(self.mcPoints)
            else:
                self.nMC = 10000
                self.mcPoints = space.sample((self.nMC, space.dimension))
        else:
            raise NotImplementedError("IVAR version 0 needs a kernel eigen-basis that no shipped kernel "
                                      "provides (experimentalDesign.py:119-146 is unreachable)")
        self._mc_dev = None

    def _mc_points(self, dev):
        if self._mc_dev is None or self._mc_dev.dev is not dev:
            self._mc_dev = dev.points(self.mcPoints)
        return self._mc_dev

    def evaluate(self, inputPoints):
        """|mean posterior variance over the MC points| for the design `inputPoints` (:79-117)."""
        assert inputPoints.shape == (self.numInputs, self.space.dimension), \
            ("inputPoints are the wrong size: ", inputPoints.shape)
        gp = self.gaussianProcess
        if self.space.noiseFunc is None:
            gp.addNodesAndComputeCovariance(inputPoints)
        else:
            addNoise = self.space.noiseFunc(inputPoints)
            gp.addNodesAndComputeCovariance(inputPoints, addNoise)
        f = gp._factor
        mc = self._mc_points(f.dev)
        _, var = f.solve_gram(mc)
        total = f.dev.zeros(1)
        check(lib.gpx_sum(f.dev.h, ptr(var), mc.n, ptr(total), f.dev.stream), "gpx_sum")
        cost = 1.0 / float(self.nMC) * float(total.item())
        return np.abs(cost)


    def derivative(self, inputPoints):
        """Gradient of the IVAR cost with respect to the design coordinates, shape (nPoints*dimension,)
        (experimentalDesign.py:148-179, version 1): the row mean over the MC points of
        GP.evaluateVarianceDerivative.  Squared-exponential kernels, homoscedastic noise."""
        if self.space.noiseFunc is not None:
            raise NotImplementedError("the heteroscedastic IVAR gradient is not on the device path")
        gp = self.gaussianProcess
        gp.kernel._require_derivative()
        gp.addNodesAndComputeCovariance(inputPoints)
        f = gp._factor
        mc = self._mc_points(f.dev)
        full = f.variance_gradient(mc)
        rows = f.n * gp.kernel.dimension
        out = f.dev.zeros(max(rows, 1))
        check(lib.gpx_rowsum(f.dev.h, ptr(full), rows, mc.n, mc.ld, 1.0 / float(self.nMC), ptr(out), f.dev.stream),
              "gpx_rowsum")
        return out[:rows].cpu().numpy()


class costFunctionGP_MI(costFunctionBase):
    """Krause-Guestrin mutual-information ratio var(y|A) / var(y|V minus A minus y)."""

    def __init__(self, gaussianProcess, nInputs, space, nmc=None, mcpoints=None, square=False):
        super(costFunctionGP_MI, self).__init__(nInputs, space)
        self.gaussianProcess = gaussianProcess  # aliased, not copied (experimentalDesign.py:227)
        if nmc is not None:
            self.nMC = nmc
            self.mcPoints = np.copy(mcpoints)
        else:
            if space.dimension == 2 and square is True:
                x = np.linspace(-1, 1, 10)
                self.nMC = len(x) * len(x)
                self.mcPoints = np.array(list(itertools.product(x, x)))
            else:
                self.nMC = 200
                self.mcPoints = space.sample((self.nMC, space.dimension))
        self.gaussianProcess.addNodesAndComputeCovariance(self.mcPoints)
        self._engine = None
        self._eval_engine = None
        self._eval_prefix = []

    def add_candidates(self, nCandidates, candidates):
        self.nMC = nCandidates
        self.mcPoints = copy.deepcopy(candidates)
        self.gaussianProcess.addNodesAndComputeCovariance(self.mcPoints)
        self._engine = None
        self._eval_engine = None  # the cached factorisation belongs to the old pool
        self._eval_prefix = []

    # the reference stores these at construction and never reads them again (:241-242)
    @property
    def cov(self):
        return self.gaussianProcess.covarianceMatrix

    @property
    def invcov(self):
        return self.gaussianProcess.precisionMatrix

    def _new_engine(self, n_max):
        gp = self.gaussianProcess
        noise = _nugget_arg(gp.noise)
        if isinstance(noise, np.ndarray):
            raise NotImplementedError("MI with per-point noise is not supported on the device path")
        dev = gp.kernel._bind()
        # left-looking blocked set-up (measured faster than the right-looking dense engine on one GPU as well:
        # |V| = 40 000: 2.33 s vs 2.97 s); GreedyMIEngine stays as the independent cross-check
        return ShardedMIEngine(dev, self.mcPoints, n_max, float(noise))

    def evaluate(self, index, indexAdded):
        """MI ratio of candidate `index` given the already chosen `indexAdded` (:252-285); shape (1,).

        The reference pays two pseudo-inverses per call; here the O(|V|^3) factorisation of the pool is built once and
        kept (until add_candidates or a change of kernel / noise), the chosen points are replayed only when
        `indexAdded` stops extending the previous call's list, and the scores of ALL candidates for that list are
        kept, so the reference's loop `for ind in options: evaluate(ind, indKeep)` costs one scoring pass per step."""
        added = [int(i) for i in indexAdded]
        gp = self.gaussianProcess
        key = (gp.kernel._gpx_spec()[0], tuple(np.ravel(gp.kernel._gpx_spec()[2])), float(_nugget_arg(gp.noise)))
        eng = self._eval_engine
        if eng is None or self._eval_key != key:
            eng = self._eval_engine = self._new_engine(max(64, 2 * (len(added) + 1)))
            self._eval_key, self._eval_prefix, self._eval_scores = key, [], None
        else:
            gp.kernel._bind(eng.dev)
        if added[: len(self._eval_prefix)] != self._eval_prefix or len(added) > eng.ncap:
            eng.reset(max(64, 2 * (len(added) + 1)))
            self._eval_prefix, self._eval_scores = [], None
        if len(added) > len(self._eval_prefix) or self._eval_scores is None:
            for i in added[len(self._eval_prefix):]:
                eng.force(i)
            self._eval_prefix = added
            eng.score()
            self._eval_scores = eng.scores[: eng.pool.n].cpu().numpy()
        return self._eval_scores[int(index): int(index) + 1].copy()


class ExperimentalDesign(object):
    """Base of the continuous optimisers (experimentalDesign.py:296-343): holds the cost function and the
    probability-density bound penalty."""
    nMCpoints = 10000

    def __init__(self, costFunction, nPoints, nDims, **kwargs):
        self.costFunction = costFunction
        self.nPoints = nPoints
        self.nDims = nDims
        super(ExperimentalDesign, self).__init__()

    def boundsFunction(self, optPoints):
        """+1 if every point has non-zero density under space.probDensity, else -1 (:310-343)."""
        if len(np.shape(optPoints)) == 1:
            optPoints = np.reshape(optPoints, (int(len(optPoints) / self.nDims), self.nDims))
        out = self.costFunction.space.probDensity(optPoints)
        out[out == 0.0] = -1e0
        if np.min(out) < 0.0:
            return -1e0
        else:
            return 1e0


class ExperimentalDesignDerivative(ExperimentalDesign):
    """SLSQP polish of a design with the analytic IVAR gradient (experimentalDesign.py:345-497).  The objective and
    its gradient are the device cost function (`evaluate`, `derivative`); the optimiser itself is scipy's SLSQP, the
    branch the reference takes when nlopt is not installed (:461-497)."""

    def __init__(self, costFunction, nPoints, nDims):
        self.addObj = lambda x: 0
        self.addGrad = lambda x: 0
        super(ExperimentalDesignDerivative, self).__init__(costFunction, nPoints, nDims)

    def addPenaltyToObjective(self, addObj, addGrad):
        self.addObj = addObj
        self.addGrad = addGrad

    def beginWithVarGreedy(self, nodesKeep=None, lbounds=[], rbounds=[]):
        """Start from the greedy max-variance ("entropy") design over the MC points, then polish (:379-404)."""
        kTemp = copy.copy(self.costFunction.gaussianProcess.kernel)
        if nodesKeep is not None:
            mcPoints = np.concatenate((nodesKeep, self.costFunction.mcPoints), axis=0)
            indKeep = np.arange(len(nodesKeep)).tolist()
        else:
            try:
                mcPoints = self.costFunction.mcPoints[:]
            except AttributeError:
                nMC = 1000
                mcPoints = self.costFunction.space.sample((nMC, self.costFunction.space.dimension))
            indKeep = []
        startVals = performGreedyVarExperimentalDesign(kTemp, mcPoints, self.nPoints, self.nDims, indKeepStart=indKeep)
        endVals = self.begin([startVals], lbounds, rbounds)
        return endVals

    def begin(self, startValues, lbounds=[], rbounds=[]):
        """Minimise the cost from every start value and return the best end design (:406-497, scipy branch)."""
        from scipy.optimize import fmin_slsqp as slsqp

        def func(xIn, *args):
            in0 = np.reshape(xIn, (int(len(xIn) / self.nDims), self.nDims))
            out = self.costFunction.evaluate(in0) - 10.0 * np.min(np.array([self.boundsFunction(in0), 0.0]))
            return out

        def grad(xIn, *args):
            in0 = np.reshape(xIn, (int(len(xIn) / self.nDims), self.nDims))
            return self.costFunction.derivative(in0)

        if len(lbounds) == 0:
            lb = -100.0 * np.ones((len(startValues[0]) * self.nDims))
            ub = 100.0 * np.ones((len(startValues[0]) * self.nDims))
            bounds = list(zip(lb, ub))
        else:
            bounds = list(zip(lbounds, rbounds))
        sol = []
        obj = np.zeros((len(startValues)))
        for ii in range(len(startValues)):
            pts = slsqp(func, startValues[ii].reshape((len(startValues[ii]) * self.nDims)), fprime=grad, bounds=bounds,
                        acc=1e-6, iprint=1 if VERBOSE else 0)
            sol.append(pts)
            obj[ii] = func(pts)
        indBest = np.argmin(obj)
        endVals = np.reshape(sol[indBest], (int(len(sol[indBest]) / self.nDims), self.nDims))
        return endVals


def performGreedyMIExperimentalDesign(costFuncMI, nPoints, start=0, shard=None):
    """Greedy MI design over the cost function's pool (experimentalDesign.py:753-785).
    Returns the chosen POINTS (as the reference does); the indices are left in
    `costFuncMI.lastIndices`.  With `shard` (gpexp_b200.engine.Shard, every rank passing the same pool) the
    |V| x |V| factorisation is column-sharded over the ranks (ShardedMIEngine) -- required above |V| ~ 9e4."""
    if shard is not None:
        gp = costFuncMI.gaussianProcess
        noise = _nugget_arg(gp.noise)
        if isinstance(noise, np.ndarray):
            raise NotImplementedError("MI with per-point noise is not supported on the device path")
        eng = ShardedMIEngine(gp.kernel._bind(), costFuncMI.mcPoints, nPoints, float(noise), shard=shard)
    else:
        eng = costFuncMI._new_engine(nPoints)
    idx = eng.run(nPoints, start=start)
    costFuncMI.lastIndices = idx
    costFuncMI._engine = eng
    return costFuncMI.mcPoints[idx, :]


def performGreedyVarExperimentalDesign(kernel, mcPoints, nPoints, dimension, weights=None, indKeepStart=[], shard=None):
    """Greedy maximum-posterior-variance design (experimentalDesign.py:787-845).

    Same contract as the reference: returns mcPoints[indKeep, :]; `indKeepStart` seeds the design and
    is extended in place (the reference mutates the caller's list, :808); selected points stay in the
    pool; the nugget is 0.0 (:825).  Extension: with `shard` (gpexp_b200.engine.Shard) every rank passes the full
    pool and works on its own contiguous block; all ranks return the same design.
    """
    if indKeepStart == []:
        indKeep = []
    else:
        indKeep = indKeepStart
    dev = kernel._bind()
    lo, hi = 0, mcPoints.shape[0]
    if shard is not None:
        lo, hi = Shard.split(mcPoints.shape[0], shard.world, shard.rank)
    pool = dev.points(mcPoints[lo:hi])
    eng = GreedyVarEngine(dev, pool, max(nPoints, len(indKeep)), weights=None if weights is None else weights[lo:hi],
                          noise=0.0, shard=shard, index_offset=lo)
    for seed in list(indKeep):
        eng.force(int(seed))

    def progress(have):
        if VERBOSE and have % 10 == 0:
            print("Number of points we have ", have)

    idx = eng.run(nPoints, progress=progress)
    for i in idx[len(indKeep):]:
        indKeep.append(int(i))
    return mcPoints[indKeep, :]


def performGreedyIVARExperimentalDesign(costFuncIVAR, candidates, nPoints, returnIndices=False, shard=None, resident=None):
    """Discrete greedy IVAR: at every step score `costFuncIVAR.evaluate(design + [c])` for every candidate
    c and add the arg-min -- what a loop over costFunctionGP_IVAR.evaluate (experimentalDesign.py:79-117)
    computes, restated with the Schur identity and run as one FP64 tensor-core contraction per step.

    candidates : (C, d) array.  With `shard` (a gpexp_b200.engine.Shard) every rank passes the FULL
    candidate array and scores its own contiguous block.
    resident : keep the M x C posterior covariance in HBM and update it by one rank-1 pass per step instead of
    re-contracting (identical picks, 16*M*C bytes per step instead of 2*M*n*C flop).  None = automatic: on when
    the matrix takes less than half of the free device memory.
    """
    gp = costFuncIVAR.gaussianProcess
    if costFuncIVAR.space.noiseFunc is not None:
        raise NotImplementedError("greedy IVAR with a heteroscedastic noise function is not on the device path")
    noise = _nugget_arg(gp.noise)
    if isinstance(noise, np.ndarray):
        raise NotImplementedError("greedy IVAR needs a scalar noise")
    dev = gp.kernel._bind()
    fam, d, params = gp.kernel._gpx_spec()
    lo, hi = 0, candidates.shape[0]
    if shard is not None:
        lo, hi = Shard.split(candidates.shape[0], shard.world, shard.rank)
    cand = dev.points(candidates[lo:hi])
    mc = dev.points(costFuncIVAR.mcPoints)
    if resident is None:
        import torch
        free, _ = torch.cuda.mem_get_info(dev.torch_device)
        resident = 8.0 * mc.n * cand.ld < 0.5 * free
        if shard is not None and shard.world > 1:  # every rank must take the same path
            flag = torch.tensor([1 if resident else 0], device=dev.torch_device)
            shard.dist.all_reduce(flag, op=shard.dist.ReduceOp.MIN, group=shard.group)
            resident = bool(flag.item())
    scale = prior_scale(fam, params)
    eng = GreedyIVAREngine(dev, cand, mc, nPoints, float(noise), scale, shard=shard, index_offset=lo, resident=bool(resident))
    idx = eng.run(nPoints)
    costFuncIVAR.lastIndices = idx
    costFuncIVAR.lastScores = eng.pick_scores[: eng.n].cpu().numpy()
    costFuncIVAR.lastPivots = eng.pivots()
    # first step from which 1e-9 agreement with the reference's pinv arithmetic cannot be expected (None: whole design)
    costFuncIVAR.illConditionedFrom = eng.ill_conditioned_from(scale)
    if returnIndices:
        return idx
    return candidates[idx, :]


def scoreCandidatesIVAR(costFuncIVAR, design, candidates):
    """One stateless IVAR scoring pass from HOST buffers: cost of design + [c] for every candidate c.
    Returns (costs (C,), argmin index).  This is the end-to-end call bench.py times (`e2e`)."""
    gp = costFuncIVAR.gaussianProcess
    noise = _nugget_arg(gp.noise)
    dev = gp.kernel._bind()
    fam, d, params = gp.kernel._gpx_spec()
    cand = dev.points(candidates)
    mc = dev.points(costFuncIVAR.mcPoints)
    n = design.shape[0]
    eng = GreedyIVAREngine(dev, cand, mc, max(n, 1), float(noise), prior_scale(fam, params))
    if n:
        eng.load_design(DesignFactor(dev, dev.points(design), float(noise)))
    eng.score()
    costs = eng.scores[: cand.n].cpu().numpy()
    return costs, int(eng.idx.item())
