"""Fused Gram/TRSM (gpx_trsm_gram) with its forward-substitution kernel: CUDA-event times at n = 255 and n = 1024 over
1e5 columns, and the residual |U^T W - K(D, X)| on a slice of columns (numpy on the host).  Used in round 2 to choose
the block shape of tri_solve_kernel (numbers in DESIGN.md section 4)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpexp_b200 import kernels  # noqa: E402
from gpexp_b200._lib import check, lib  # noqa: E402
from gpexp_b200.device import Device, ptr  # noqa: E402
from gpexp_b200.engine import DesignFactor  # noqa: E402

dev = Device.get(0)
rng = np.random.default_rng(0)
kern = kernels.KernelSquaredExponential([0.06, 0.09], 1.0, 2)
kern._bind(dev)
C = 100_000
X = dev.points(rng.uniform(-1, 1, (C, 2)))
out = {}
for n in (255, 1024):
    design = rng.uniform(-1, 1, (n, 2))
    f = DesignFactor(dev, dev.points(design), 1e-4)
    W = dev.zeros(n, X.ld)
    f.solve_gram(X, W=W)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        f.solve_gram(X, W=W)
    b.record()
    torch.cuda.synchronize()
    out[f"trsm_gram_n{n}_ms"] = round(a.elapsed_time(b) / 5, 4)
    # check: U^T W = K(D, X) on a slice of columns
    U = f.U[:n, :n].cpu().numpy()
    cols = slice(0, 4000)
    Wd = W[:n, cols].cpu().numpy()
    Xc = X.X[:2, cols].t().cpu().numpy()
    d2 = ((design[:, None, :] - Xc[None, :, :]) ** 2 / np.array([0.06, 0.09]) ** 2).sum(-1)
    Kref = np.exp(-0.5 * d2)
    out[f"resid_n{n}"] = float(np.abs(np.triu(U).T @ Wd - Kref).max())
    del W, f
print(json.dumps(out))
