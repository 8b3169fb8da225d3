"""Small cases of the mbarrier / TMA kernels for compute-sanitizer (racecheck, memcheck): the IVAR contraction on every
ring (row copies, tensor-map loads, both prologue forms, K tail), the padded update kernel with and without the
lower-triangular skipping, a few steps of the C-side greedy loops, potrf + TRSM.  Results are checked against numpy."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpexp_b200.experimentalDesign as ed  # noqa: E402
from gpexp_b200 import gp, kernels  # noqa: E402
from gpexp_b200._lib import check, lib  # noqa: E402
from gpexp_b200.approximation import Space  # noqa: E402
from gpexp_b200.device import Device, ptr, roundup  # noqa: E402
from gpexp_b200.engine import DesignFactor, GreedyIVAREngine, prior_scale  # noqa: E402
from oracle import gpexp_oracle as orc  # noqa: E402  (checker)

ed.VERBOSE = False
dev = Device.get(0)
rng = np.random.default_rng(1)
cand, mc = rng.uniform(-1, 1, (300, 2)), rng.uniform(-1, 1, (420, 2))
kern = kernels.KernelSquaredExponential([0.2, 0.3], 1.0, 2)
kern._bind(dev)
fam, d, params = kern._gpx_spec()
ks = orc.KernelSpec.se([0.2, 0.3], 1.0, 2)
design = cand[:37]                                                  # K tail: 37 is not a multiple of 4
w_m, var_m = orc.fast_design_state(ks, design, mc, 1e-6)
w_c, var_c = orc.fast_design_state(ks, design, cand, 1e-6)
ref = orc.fast_ivar_scores(ks, cand, mc, w_m, var_m, w_c, var_c, 1e-6)
for ring in (0, 1, 2):
    for diff in (False, True):
        dev.force_diff_form = diff
        check(lib.gpx_set_ivar_ring(dev.h, ring))
        eng = GreedyIVAREngine(dev, dev.points(cand), dev.points(mc), 40, 1e-6, prior_scale(fam, params))
        eng.load_design(DesignFactor(dev, dev.points(design), 1e-6))
        eng.score()
        got = eng.scores[:300].cpu().numpy()
        assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-9, (ring, diff)
        eng.run(40)                                                  # three steps of gpx_ivar_greedy_run
        torch.cuda.synchronize()
        print("ring", ring, "diff" if diff else "expanded", "ok", flush=True)
dev.force_diff_form = False
check(lib.gpx_set_ivar_ring(dev.h, 1))
# padded update kernel, plain and lower-triangular block-cyclic
K, I, B, world = 700, 256, 256, 2
A = rng.standard_normal((K, I))
full = np.tril(rng.standard_normal((1024, 1024)))
for rank in range(world):
    gcols = np.concatenate([np.arange(g * B, (g + 1) * B) for g in range(rank, 4, world)])
    Bl = full[:K, gcols]
    ld = roundup(gcols.size)
    Ad, Bd = dev.zeros(K, 256), dev.zeros(K, ld)
    Ad[:, :I] = dev.upload(A)
    Bd[:, : gcols.size] = dev.upload(Bl)
    for lower in (False, True):
        Cd = dev.zeros(I, ld)
        if lower:
            check(lib.gpx_dgemm_tn_sub_lower(dev.h, ptr(Ad), 256, ptr(Bd), ld, ptr(Cd), ld, I, gcols.size, K, B, world, rank, dev.stream))
        else:
            check(lib.gpx_dgemm_tn_sub_padded(dev.h, ptr(Ad), 256, ptr(Bd), ld, ptr(Cd), ld, I, gcols.size, K, 0, dev.stream))
        assert np.max(np.abs(Cd[:, : gcols.size].cpu().numpy() + A.T @ Bl)) <= 1e-10
print("sub_ws ok", flush=True)
# greedy variance loop, posterior variance (potrf + fused Gram/TRSM), MI engine
mk = kernels.KernelIsoMatern(1.0, 1.0, 3)
pool = rng.uniform(-1, 1, (500, 3))
pts = ed.performGreedyVarExperimentalDesign(mk, pool, 12, 3)
vidx, _ = orc.fast_greedy_var(orc.KernelSpec.matern32(1.0, 1.0, 3), pool, 12)
assert np.array_equal(pts, pool[vidx])
g = gp.GP(kern, 1e-4)
g.addNodesAndComputeCovariance(rng.uniform(-1, 1, (150, 2)))
assert np.all(np.isfinite(g.evaluateVariance(mc)))
hk = kernels.KernelMehlerND([0.9, 0.9, 0.9], 3)
vp = rng.standard_normal((300, 3))
cmi = ed.costFunctionGP_MI(gp.GP(hk, 1e-2), 5, Space(3, None, None), nmc=300, mcpoints=vp)
ed.performGreedyMIExperimentalDesign(cmi, 5, start=0)
midx, _ = orc.fast_greedy_mi(orc.KernelSpec.mehler([0.9, 0.9, 0.9], 3), vp, 1e-2, 5, start=0)
assert [int(i) for i in cmi.lastIndices] == midx
torch.cuda.synchronize()
print("sanitize_case OK", flush=True)
