"""configs[0] (1 000 candidates x 10 000 MC points, 20 steps) twice through the C-side loop: for an ncu launch list
(gpu__time_duration per launch) that shows where a ~60 us greedy step goes."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpexp_b200.experimentalDesign as ed  # noqa: E402
from gpexp_b200 import gp, kernels  # noqa: E402
from gpexp_b200.approximation import Space  # noqa: E402

rng = np.random.default_rng(1)
cand, mc = rng.uniform(-1, 1, (1000, 1)), rng.uniform(-1, 1, (10000, 1))
k = kernels.KernelSquaredExponential([0.05], 1.0, 1)
cf = ed.costFunctionGP_IVAR(gp.GP(k, 1e-6), 1, Space(1, None, None), mcPoints=mc)
for resident, one_kernel in ((False, False), (True, False), (True, True)):
    for rep in range(2):
        eng = ed.beginGreedyIVARExperimentalDesign(cf, cand, 20, resident=resident)
        eng.ONE_KERNEL_PAIRS = 32_000_000 if one_kernel else 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.run(20)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(("one-kernel" if one_kernel else "resident") if resident else "contraction", "rep", rep, "issue %.3f ms  total %.3f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3),
              [int(i) for i in eng.indices()[:4]], flush=True)
