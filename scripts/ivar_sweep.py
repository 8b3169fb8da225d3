"""Time the IVAR scoring kernel (gpx_score_ivar) versus design size n on one GPU.
usage: [GPX_IVAR_RING=0|1|2] [GPX_FORCE_DIFF=1] python scripts/ivar_sweep.py [d] [C] [M] [n1,n2,...]
       -> gpurun_out/ivar_sweep_d<d>_ring<r>[_diff].json   (A/B runs of the ring geometry and the prologue form)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpexp_b200 import kernels  # noqa: E402
from gpexp_b200.device import Device  # noqa: E402
from gpexp_b200.engine import DesignFactor, GreedyIVAREngine, prior_scale  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
M = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
ns = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 63, 255, 1023]
dev = Device.get(0)
ring = os.environ.get("GPX_IVAR_RING", "default")  # read by gpx_create
dev.force_diff_form = os.environ.get("GPX_FORCE_DIFF", "0") == "1"
rng = np.random.default_rng(5)
cl = [0.06, 0.09] if d == 2 else list(np.linspace(0.5, 1.5, d))
kern = kernels.KernelSquaredExponential(cl, 1.0, d)
kern._bind(dev)
fam, _, params = kern._gpx_spec()
cand_h, mc_h = rng.uniform(-1, 1, (C, d)), rng.uniform(-1, 1, (M, d))
cand, mc = dev.points(cand_h), dev.points(mc_h)
out = []
for n in ns:
    eng = GreedyIVAREngine(dev, cand, mc, max(n, 1), 1e-6, prior_scale(fam, params))
    if n:
        eng.load_design(DesignFactor(dev, dev.points(cand_h[rng.permutation(C)[:n]]), 1e-6))
    eng.score()
    torch.cuda.synchronize()
    reps = 3 if n < 2000 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.score()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # flops: contraction 2*M*n*C plus the Gram prologue 2*M*dpad*C on the same pipe
    dpad = (d + 3) // 4 * 4
    row = {"d": d, "n": n, "C": C, "M": M, "ms": ms, "cand_per_s": C / ms * 1e3, "ring": ring, "diff_form": dev.force_diff_form,
           "prologue": int(eng.prologue()[0]), "scores_head": [float(v) for v in eng.scores[:8].cpu().numpy()],
           "tflops_contraction": 2.0 * M * n * C / ms / 1e9, "tflops_incl_prologue": 2.0 * M * (n + dpad) * C / ms / 1e9}
    print(json.dumps(row), flush=True)
    out.append(row)
    del eng
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/ivar_sweep_d{d}_ring{ring}{'_diff' if dev.force_diff_form else ''}.json", "w"), indent=1)
