"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference is mounted read-only at /root/reference and
does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests of its own (test/test_kernel.py:1-24 is an empty header), so these
files are what pins both the CPU oracle (oracle/gpexp_oracle.py) and the CUDA path.  Every array
below is an output of reference code (gpExp.kernels / gpExp.gp / gpExp.gp_kernel_utilities /
gpExp.experimentalDesign) on seeded numpy inputs; the inputs are stored next to the outputs.
numpy 2.3.5 / scipy 1.18.1 / OpenBLAS 0.3.30 were used.
"""
import io
import os
import sys
import warnings
from contextlib import redirect_stdout

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore", category=DeprecationWarning)

import gpExp.kernels as rk  # noqa: E402
import gpExp.gp as rgp  # noqa: E402
import gpExp.gp_kernel_utilities as rku  # noqa: E402
import gpExp.experimentalDesign as red  # noqa: E402
from gpExp.approximation import Space  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mk_kernel(name):
    if name == "se_iso_1d":
        return rk.KernelSquaredExponential([0.05], 1.0, 1), 1
    if name == "se_ard_2d":
        return rk.KernelSquaredExponential([0.06, 0.09], 1.0, 2), 2
    if name == "se_ard_2d_wide":
        return rk.KernelSquaredExponential([0.3, 0.45], 1.7, 2), 2
    if name == "se_ard_10d":
        return rk.KernelSquaredExponential(list(np.linspace(0.5, 1.5, 10)), 1.0, 10), 10
    if name == "matern_5d":
        return rk.KernelIsoMatern(1.0, 1.0, 5), 5
    if name == "matern_5d_b":
        return rk.KernelIsoMatern(0.5, 2.5, 5), 5
    if name == "mehler_3d":
        return rk.KernelMehlerND([0.9, 0.9, 0.9], 3), 3
    if name == "mehler_3d_b":
        return rk.KernelMehlerND([0.5, 0.7, 0.3], 3), 3
    if name == "mehler_1d":
        return rk.KernelMehler1D(0.6, 1), 1
    raise KeyError(name)


KERNELS = ["se_iso_1d", "se_ard_2d", "se_ard_2d_wide", "se_ard_10d", "matern_5d", "matern_5d_b",
           "mehler_3d", "mehler_3d_b", "mehler_1d"]


def sample(rng, name, n, d):
    if name.startswith("mehler"):
        return rng.standard_normal((n, d))
    return rng.uniform(-1.0, 1.0, (n, d))


def quiet(fn, *a, **k):
    with redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def match_rows(points, pool):
    """Recover indices from the rows the reference returns (it returns points, not indices)."""
    out = []
    for p in points:
        hit = np.where(np.all(pool == p, axis=1))[0]
        out.append(int(hit[0]))
    return np.array(out, dtype=np.int64)


def gen_kernels(out):
    rng = np.random.default_rng(101)
    for name in KERNELS:
        kern, d = mk_kernel(name)
        x1 = sample(rng, name, 57, d)
        x2 = sample(rng, name, 57, d)
        one = sample(rng, name, 1, d)
        out[f"kern/{name}/x1"] = x1
        out[f"kern/{name}/x2"] = x2
        out[f"kern/{name}/one"] = one
        out[f"kern/{name}/pair"] = kern.evaluate(x1, x2)
        out[f"kern/{name}/bcast_right"] = kern.evaluate(x1, one)
        out[f"kern/{name}/bcast_left"] = kern.evaluate(one, x2)
        out[f"kern/{name}/prior"] = kern.evaluate(x1, x1)


def gen_gram(out):
    rng = np.random.default_rng(102)
    for name in ["se_ard_2d", "se_ard_10d", "matern_5d", "mehler_3d_b"]:
        kern, d = mk_kernel(name)
        pts = sample(rng, name, 45, d)
        nug = rng.uniform(1e-4, 1e-2, 45)
        out[f"gram/{name}/pts"] = pts
        out[f"gram/{name}/nugvec"] = nug
        out[f"gram/{name}/K0"] = rku.calculateCovarianceMatrix(kern, pts)
        out[f"gram/{name}/Kscalar"] = rku.calculateCovarianceMatrix(kern, pts, 1e-3)
        out[f"gram/{name}/Kvec"] = rku.calculateCovarianceMatrix(kern, pts, nug)


def gen_gp(out):
    rng = np.random.default_rng(103)
    for name, noise in [("se_ard_2d_wide", 1e-6), ("matern_5d", 0.0), ("mehler_3d", 1e-2), ("se_ard_10d", 1e-6)]:
        kern, d = mk_kernel(name)
        nodes = sample(rng, name, 40, d)
        query = np.vstack([sample(rng, name, 297, d), nodes[:3]])
        fvals = np.sin(nodes.sum(axis=1))
        gp = rgp.GP(kern, noise)
        gp.train(nodes, fvals)
        var = gp.evaluateVariance(query, parallel=0)
        mean, absvar = gp.evaluate(query, compvar=1)
        mean2, cov = gp.evaluate(query[:25], compvar=2)
        out[f"gp/{name}/nodes"] = nodes
        out[f"gp/{name}/query"] = query
        out[f"gp/{name}/fvals"] = fvals
        out[f"gp/{name}/noise"] = np.float64(noise)
        out[f"gp/{name}/cov"] = gp.covarianceMatrix
        out[f"gp/{name}/prec"] = gp.precisionMatrix
        out[f"gp/{name}/coeff"] = gp.coeff
        out[f"gp/{name}/var"] = var
        out[f"gp/{name}/mean"] = mean
        out[f"gp/{name}/absvar"] = absvar
        out[f"gp/{name}/cov25"] = cov
        out[f"gp/{name}/cond"] = np.float64(np.linalg.cond(gp.covarianceMatrix))
        # IVAR cost of this design (experimentalDesign.py:79-117), homoscedastic
        space = Space(d, None, None, noise=None)
        mc = sample(rng, name, 1500, d)
        cf = red.costFunctionGP_IVAR(gp, 40, space, mcPoints=mc)
        out[f"gp/{name}/mc"] = mc
        out[f"gp/{name}/ivar_cost"] = np.float64(cf.evaluate(nodes))
    # heteroscedastic branch (:110-114): per-point nugget from space.noiseFunc
    kern, d = mk_kernel("se_ard_2d_wide")
    nodes = sample(rng, "se", 25, d)
    mc = sample(rng, "se", 1000, d)
    nf = lambda p: 1e-4 + 1e-3 * (p[:, 0] ** 2)  # noqa: E731
    space = Space(d, None, None, noise=nf)
    cf = red.costFunctionGP_IVAR(rgp.GP(kern, 1e-6), 25, space, mcPoints=mc)
    out["gp/hetero/nodes"] = nodes
    out["gp/hetero/mc"] = mc
    out["gp/hetero/ivar_cost"] = np.float64(cf.evaluate(nodes))


def gen_greedy_var(out):
    rng = np.random.default_rng(104)
    cases = [("se_iso_1d", 300, 12, False, []), ("matern_5d", 400, 30, False, []),
             ("mehler_3d_b", 250, 10, True, []), ("se_ard_2d", 350, 25, True, [7, 3]),
             ("se_ard_10d", 300, 20, False, [])]
    for ci, (name, c, n, use_w, seeds) in enumerate(cases):
        kern, d = mk_kernel(name)
        pool = sample(rng, name, c, d)
        w = rng.uniform(0.5, 1.5, c) if use_w else None
        pts = quiet(red.performGreedyVarExperimentalDesign, kern, pool, n, d, weights=w,
                    indKeepStart=list(seeds))
        idx = match_rows(pts, pool)
        # per-step scores from the reference GP (same k(x,x) - k^T pinv(K) k, nugget 0.0)
        scores = np.zeros((n, c))
        for step in range(len(seeds), n):
            if step == 0:
                k = kern.evaluate(pool, pool)
            else:
                gp = rgp.GP(kern, 0.0)
                gp.addNodesAndComputeCovariance(pool[idx[:step]])
                k = gp.evaluateVariance(pool, parallel=0)
            scores[step] = k * w if use_w else k
        out[f"gvar/{ci}/name"] = np.array(name)
        out[f"gvar/{ci}/pool"] = pool
        out[f"gvar/{ci}/weights"] = w if use_w else np.zeros(0)
        out[f"gvar/{ci}/seeds"] = np.array(seeds, dtype=np.int64)
        out[f"gvar/{ci}/idx"] = idx
        out[f"gvar/{ci}/scores"] = scores


def gen_greedy_ivar(out):
    rng = np.random.default_rng(105)
    cases = [("se_iso_1d", 80, 1000, 8, 1e-6), ("se_ard_2d", 100, 1500, 10, 1e-6),
             ("matern_5d", 90, 1200, 8, 1e-4), ("mehler_3d", 70, 1000, 6, 1e-2),
             ("se_ard_2d_wide", 60, 800, 6, 0.0)]
    for ci, (name, c, m, n, noise) in enumerate(cases):
        kern, d = mk_kernel(name)
        cand = sample(rng, name, c, d)
        mc = sample(rng, name, m, d)
        space = Space(d, None, None, noise=None)
        gp = rgp.GP(kern, float(noise))
        idx, costs, conds = [], np.zeros((n, c)), np.zeros(n)
        for step in range(n):
            cf = red.costFunctionGP_IVAR(gp, step + 1, space, mcPoints=mc)
            for j in range(c):
                costs[step, j] = cf.evaluate(np.vstack([cand[idx], cand[j:j + 1]]))
            idx.append(int(np.argmin(costs[step])))
            kd = rku.calculateCovarianceMatrix(kern, cand[idx], float(noise))
            conds[step] = np.linalg.cond(kd)
        out[f"givar/{ci}/name"] = np.array(name)
        out[f"givar/{ci}/cand"] = cand
        out[f"givar/{ci}/mc"] = mc
        out[f"givar/{ci}/noise"] = np.float64(noise)
        out[f"givar/{ci}/idx"] = np.array(idx, dtype=np.int64)
        out[f"givar/{ci}/costs"] = costs
        out[f"givar/{ci}/cond"] = conds


def gen_greedy_mi(out):
    rng = np.random.default_rng(106)
    cases = [("mehler_3d", 80, 8, 1e-2, 0), ("matern_5d", 60, 7, 1e-3, 5), ("se_ard_2d_wide", 50, 6, 1e-2, 0)]
    for ci, (name, v, n, noise, start) in enumerate(cases):
        kern, d = mk_kernel(name)
        pool = sample(rng, name, v, d)
        space = Space(d, None, None, noise=None)
        gp = rgp.GP(kern, float(noise))
        cf = red.costFunctionGP_MI(gp, n, space, nmc=v, mcpoints=pool)
        pts = red.performGreedyMIExperimentalDesign(cf, n, start=start)
        idx = match_rows(pts, pool)
        scores = np.full((n, v), -np.inf)
        for step in range(1, n):
            for j in range(v):
                if j in idx[:step]:
                    continue
                scores[step, j] = cf.evaluate(j, list(idx[:step]))[0]
        out[f"gmi/{ci}/name"] = np.array(name)
        out[f"gmi/{ci}/pool"] = pool
        out[f"gmi/{ci}/noise"] = np.float64(noise)
        out[f"gmi/{ci}/start"] = np.int64(start)
        out[f"gmi/{ci}/idx"] = idx
        out[f"gmi/{ci}/scores"] = scores
        out[f"gmi/{ci}/cov"] = cf.cov
        out[f"gmi/{ci}/invcov"] = cf.invcov


def _cfg1_chunk(args):
    """Worker of gen_cfg1: reference IVAR cost of design + [c] for a slice of the candidates."""
    cl, noise, design, cand, mc, lo, hi = args
    kern = rk.KernelSquaredExponential([cl], 1.0, 1)
    gp = rgp.GP(kern, float(noise))
    cf = red.costFunctionGP_IVAR(gp, design.shape[0] + 1, Space(1, None, None, noise=None), mcPoints=mc)
    return np.array([cf.evaluate(np.vstack([design, cand[j:j + 1]])) for j in range(lo, hi)])


def gen_cfg1(out):
    """BASELINE.json configs[0] at FULL size through the unmodified reference: 1-D isotropic SE,
    greedy IVAR design of 20 points from 1 000 candidates x 10 000 MC points, seed 1 -- (a) cl = 0.05, noise 1e-6 (the
    well-conditioned measurement input) and (b) the demo.py:52-58 stress values cl = 0.3, noise 0.0, whose design Gram
    reaches cond ~1e17.  The 20 x 1000 reference evaluations of a variant are independent within a step, so the
    candidates are spread over worker processes; every number is still produced by
    costFunctionGP_IVAR.evaluate (experimentalDesign.py:79-117)."""
    import multiprocessing as mp
    rng = np.random.default_rng(1)
    cand = rng.uniform(-1.0, 1.0, (1000, 1))
    mc = rng.uniform(-1.0, 1.0, (10000, 1))
    out["cfg1/cand"] = cand
    out["cfg1/mc"] = mc
    nproc = max(1, min(8, (os.cpu_count() or 2)))
    edges = np.linspace(0, cand.shape[0], nproc + 1).astype(int)
    with mp.get_context("fork").Pool(nproc) as pool:
        for tag, cl, noise, n in [("main", 0.05, 1e-6, 20), ("stress", 0.3, 0.0, 20)]:
            kern = rk.KernelSquaredExponential([cl], 1.0, 1)
            idx, costs, conds = [], np.zeros((n, cand.shape[0])), np.zeros(n)
            for step in range(n):
                design = cand[idx]
                parts = pool.map(_cfg1_chunk, [(cl, noise, design, cand, mc, int(a), int(b))
                                               for a, b in zip(edges[:-1], edges[1:])])
                costs[step] = np.concatenate(parts)
                idx.append(int(np.argmin(costs[step])))
                conds[step] = np.linalg.cond(rku.calculateCovarianceMatrix(kern, cand[idx], float(noise)))
                print(tag, "step", step, "pick", idx[-1], "cond %.3g" % conds[step], flush=True)
            out[f"cfg1/{tag}/cl"] = np.float64(cl)
            out[f"cfg1/{tag}/noise"] = np.float64(noise)
            out[f"cfg1/{tag}/idx"] = np.array(idx, dtype=np.int64)
            out[f"cfg1/{tag}/costs"] = costs
            out[f"cfg1/{tag}/cond"] = conds


def main():
    only = sys.argv[1:]
    for fname, gen in [("kernels.npz", gen_kernels), ("gram.npz", gen_gram), ("gp.npz", gen_gp),
                       ("greedy_var.npz", gen_greedy_var), ("greedy_ivar.npz", gen_greedy_ivar),
                       ("greedy_mi.npz", gen_greedy_mi), ("next.npz", gen_next), ("cfg1.npz", gen_cfg1), ("next2.npz", gen_next2)]:
        if only and fname not in only:
            continue
        out = {}
        gen(out)
        np.savez_compressed(os.path.join(HERE, fname), **out)
        print(fname, len(out), "arrays", os.path.getsize(os.path.join(HERE, fname)) // 1024, "KiB")



def gen_next(out):
    """SURVEY.md 8(f) widening: marginal log-likelihood (gp.py:373-446) and the IVAR gradient with respect to the
    design coordinates (experimentalDesign.py:148-179 -> gp.py:282-341 -> kernels.py:146-181)."""
    rng = np.random.default_rng(107)
    for name, noise in [("se_ard_2d_wide", 1e-6), ("matern_5d", 1e-4), ("mehler_3d", 1e-2), ("se_ard_10d", 1e-6)]:
        kern, d = mk_kernel(name)
        nodes = sample(rng, name, 35, d)
        fvals = np.cos(nodes.sum(axis=1))
        gp = rgp.GP(kern, noise)
        out[f"next/loglike/{name}/nodes"] = nodes
        out[f"next/loglike/{name}/fvals"] = fvals
        out[f"next/loglike/{name}/noise"] = np.float64(noise)
        out[f"next/loglike/{name}/value"] = np.float64(gp.computeLogLike(nodes, fvals))
    for name, noise, n, m in [("se_ard_2d_wide", 1e-6, 9, 400), ("se_ard_10d", 1e-6, 14, 300), ("se_iso_1d", 1e-4, 6, 200)]:
        kern, d = mk_kernel(name)
        design = sample(rng, name, n, d)
        mc = sample(rng, name, m, d)
        one = sample(rng, name, 1, d)
        gp = rgp.GP(kern, noise)
        gp.addNodesAndComputeCovariance(design)
        cf = red.costFunctionGP_IVAR(gp, n, Space(d, None, None, noise=None), mcPoints=mc)
        out[f"next/grad/{name}/design"] = design
        out[f"next/grad/{name}/mc"] = mc
        out[f"next/grad/{name}/one"] = one
        out[f"next/grad/{name}/noise"] = np.float64(noise)
        out[f"next/grad/{name}/kderiv"] = kern.derivative(mc, one)
        out[f"next/grad/{name}/var_deriv"] = gp.evaluateVarianceDerivative(mc[:64])
        out[f"next/grad/{name}/ivar_deriv"] = cf.derivative(design)
        out[f"next/grad/{name}/cond"] = np.float64(np.linalg.cond(gp.covarianceMatrix))
    # continuous polish of a design with SLSQP (ExperimentalDesignDerivative.begin, experimentalDesign.py:406-497;
    # nlopt is absent, so the scipy branch runs), started from the greedy max-variance design (:379-404)
    kern, d = mk_kernel("se_ard_2d_wide")
    mc = sample(rng, "se", 600, d)
    dens = lambda p: np.all(np.abs(p) <= 1.0, axis=1).astype(float)  # noqa: E731
    space = Space(d, lambda s: rng.uniform(-1, 1, s), dens, noise=None)
    gp = rgp.GP(kern, 1e-6)
    cf = red.costFunctionGP_IVAR(gp, 5, space, mcPoints=mc)
    exp = red.ExperimentalDesignDerivative(cf, 5, d)
    start = quiet(red.performGreedyVarExperimentalDesign, kern, mc, 5, d)
    end = quiet(exp.begin, [start], list(-np.ones(10)), list(np.ones(10)))
    end2 = quiet(exp.beginWithVarGreedy, None, list(-np.ones(10)), list(np.ones(10)))
    out["next/slsqp/mc"] = mc
    out["next/slsqp/start"] = start
    out["next/slsqp/end"] = end
    out["next/slsqp/end_greedy"] = end2
    out["next/slsqp/cost_start"] = np.float64(cf.evaluate(start))
    out["next/slsqp/cost_end"] = np.float64(cf.evaluate(end))



from make_golden_shared import QuadNoise  # noqa: E402  (same directory)


def gen_next2(out):
    """SURVEY.md 8(f) rows 2 and 4 from the unmodified reference: heteroscedastic variance derivative (gp.py:282-341,
    noiseFunc branch) and IVAR gradient, the batch-greedy optimiser wrapper (experimentalDesign.py:694-751), FITC
    covariance / precision / prediction (gp.py:182-208, gp_kernel_utilities.py:70-104), covTimesV and the Nystrom
    eigenvalues (gp_kernel_utilities.py:107-194)."""
    rng = np.random.default_rng(108)
    nf = QuadNoise()
    for name, n, m in [("se_ard_2d_wide", 8, 300), ("se_iso_1d", 5, 200)]:
        kern, d = mk_kernel(name)
        design = sample(rng, name, n, d)
        mc = sample(rng, name, m, d)
        gp = rgp.GP(kern, 1e-6)
        gp.addNodesAndComputeCovariance(design, noiseIn=nf(design))
        cf = red.costFunctionGP_IVAR(gp, n, Space(d, None, None, noise=nf), mcPoints=mc)
        out[f"next2/hetero/{name}/design"] = design
        out[f"next2/hetero/{name}/mc"] = mc
        out[f"next2/hetero/{name}/var_deriv"] = gp.evaluateVarianceDerivative(mc[:48], noiseFunc=nf)
        out[f"next2/hetero/{name}/ivar_deriv"] = cf.derivative(design)
        out[f"next2/hetero/{name}/ivar_cost"] = np.float64(cf.evaluate(design))
    # batch-greedy continuous design: 2 + 2 points, each batch = greedy max-variance start + SLSQP polish
    kern, d = mk_kernel("se_ard_2d_wide")
    mc = sample(rng, "se", 500, d)
    srng = np.random.default_rng(7)
    dens = lambda p: np.all(np.abs(p) <= 1.0, axis=1).astype(float)  # noqa: E731
    space = Space(d, lambda s: srng.uniform(-1, 1, s), dens, noise=None)
    cf = red.costFunctionGP_IVAR(rgp.GP(kern, 1e-6), 2, space, mcPoints=mc)
    exp = red.ExperimentalDesignGreedyWithDerivatives(cf, 4, 2, d)
    pts = quiet(exp.begin)
    out["next2/batch/mc"] = mc
    out["next2/batch/points"] = pts
    out["next2/batch/cost"] = np.float64(red.costFunctionGP_IVAR(rgp.GP(kern, 1e-6), 4, space, mcPoints=mc).evaluate(pts))
    # FITC sparse GP
    for name, noise in [("se_ard_2d_wide", 1e-4), ("matern_5d", 1e-3)]:
        kern, d = mk_kernel(name)
        nodes = sample(rng, name, 30, d)
        query = sample(rng, name, 120, d)
        fvals = np.sin(nodes.sum(axis=1))
        np.random.seed(11)
        gp = rgp.GP(kern, noise, FITC=0.5)
        gp.train(nodes, fvals)
        out[f"next2/fitc/{name}/nodes"] = nodes
        out[f"next2/fitc/{name}/query"] = query
        out[f"next2/fitc/{name}/fvals"] = fvals
        out[f"next2/fitc/{name}/noise"] = np.float64(noise)
        out[f"next2/fitc/{name}/inducing"] = gp.fitcnodes
        out[f"next2/fitc/{name}/cov"] = gp.covarianceMatrix
        out[f"next2/fitc/{name}/prec"] = gp.precisionMatrix
        out[f"next2/fitc/{name}/coeff"] = gp.coeff
        out[f"next2/fitc/{name}/var"] = gp.evaluateVariance(query, parallel=0)
        mean, absvar = gp.evaluate(query, compvar=1)
        out[f"next2/fitc/{name}/mean"] = mean
        out[f"next2/fitc/{name}/absvar"] = absvar
        np.random.seed(11)
        out[f"next2/fitc/{name}/loglike"] = np.float64(rgp.GP(kern, noise, FITC=0.5).computeLogLike(nodes, fvals))
        covmat, precmat, sn = rku.calculateCovarianceMatrixFITC(kern, nodes, noise, gp.fitcnodes, returnCov=True)
        out[f"next2/fitc/{name}/util_cov"] = covmat
        out[f"next2/fitc/{name}/util_prec"] = precmat
    # matrix-free Gram x vector and the Nystrom eigenvalues
    for name in ["se_ard_2d_wide", "matern_5d", "mehler_3d_b"]:
        kern, d = mk_kernel(name)
        pts = sample(rng, name, 260, d)
        b = rng.standard_normal(260)
        out[f"next2/matvec/{name}/pts"] = pts
        out[f"next2/matvec/{name}/b"] = b
        out[f"next2/matvec/{name}/Kb"] = np.asarray(quiet(rku.covTimesV, b, kern, pts)).reshape(-1)
    kern, d = mk_kernel("se_ard_2d_wide")
    pts = sample(rng, "se", 200, d)
    eigv, eigve = quiet(rku.calculateKernelBasisFunctionsMC, kern, 6, pts)
    out["next2/nystrom/pts"] = pts
    out["next2/nystrom/eigv"] = eigv
    out["next2/nystrom/eigve"] = eigve


if __name__ == "__main__":
    main()
