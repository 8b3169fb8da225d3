"""Where the ~12 ms between the scoring kernel (154 ms) and the end-to-end call (167 ms) of scoreCandidatesIVAR go:
wall-clock of each stage with a device synchronisation after it (cfg-2 sizes, pinned host inputs)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpexp_b200.experimentalDesign as ed  # noqa: E402
from gpexp_b200 import gp, kernels  # noqa: E402
from gpexp_b200.approximation import Space  # noqa: E402
from gpexp_b200.device import Device  # noqa: E402
from gpexp_b200.engine import DesignFactor, GreedyIVAREngine, prior_scale  # noqa: E402

rng = np.random.default_rng(2)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
C_LOCAL = int(os.environ.get("E2E_CANDIDATES", "100000"))  # 12500 = one rank's share of cfg-2 on 8 GPUs
cand_h, mc_h = pin(rng.uniform(-1, 1, (C_LOCAL, 2))), pin(rng.uniform(-1, 1, (100_000, 2)))
design_h = pin(cand_h[rng.permutation(C_LOCAL)[:255]])
kern = kernels.KernelSquaredExponential([0.06, 0.09], 1.0, 2)
cf = ed.costFunctionGP_IVAR(gp.GP(kern, 1e-6), 1, Space(2, None, None), mcPoints=mc_h)
dev = Device.get(0)
for rep in range(3):
    ed.scoreCandidatesIVAR(cf, design_h, cand_h)
torch.cuda.synchronize()
t0 = time.perf_counter()
for rep in range(3):
    ed.scoreCandidatesIVAR(cf, design_h, cand_h)
torch.cuda.synchronize()
print("whole call: %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))


def stage(name, fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    print("  %-34s %8.3f ms" % (name, (time.perf_counter() - t) * 1e3), flush=True)
    return r


for rep in range(2):
    print("rep", rep)
    kern._bind(dev)
    fam, d, params = kern._gpx_spec()
    cand = stage("upload + transpose candidates", lambda: dev.points(cand_h))
    mc = stage("upload + transpose MC points", lambda: dev.points(mc_h))
    des = stage("upload + transpose design", lambda: dev.points(design_h))
    eng = stage("engine buffers (W_C, W_M, workspace)", lambda: GreedyIVAREngine(dev, cand, mc, 255, 1e-6, prior_scale(fam, params)))
    fac = stage("Gram + Cholesky of the design", lambda: DesignFactor(dev, des, 1e-6))
    stage("W_C, W_M, variances (TRSM x 2)", lambda: eng.load_design(fac))
    stage("prepared sides + guard", lambda: eng.prologue())
    stage("score (contraction + arg-min)", lambda: eng.score())
    stage("D2H of the costs", lambda: (eng.scores[: cand.n].cpu().numpy(), int(eng.idx.item())))
