"""The HBM-bound and factorisation kernels in isolation, for ncu (Gram, fused Gram/TRSM = dmma_core<STORE> + tri_solve,
Cholesky, row append, resident update) -- and, without ncu, their CUDA-event times against the measured ceilings.

    python scripts/profile_kernels.py            -> gpurun_out/kernels_r02.json
    ncu --set full -k regex:'gram_kernel|dmma_core|tri_solve|potrf_diag|append_row|cov_update' ... python scripts/profile_kernels.py --once
"""
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

once = "--once" in sys.argv
dev = Device.get(0)
rng = np.random.default_rng(0)
out = {}


def ev_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    if once:
        return float("nan")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ceilings measured here: device copy (read + write) and pure streaming write
buf = torch.empty(4096 * 100_096, dtype=torch.float64, device="cuda")
src = torch.empty(4096 * 100_096, dtype=torch.float64, device="cuda")
t = ev_ms(lambda: buf.copy_(src), 5)
out["copy_gbs"] = 2 * buf.numel() * 8 / t / 1e6
t = ev_ms(lambda: buf.fill_(1.0), 5)
out["fill_gbs"] = buf.numel() * 8 / t / 1e6
del src

for name, kern, d in [("se_2d", kernels.KernelSquaredExponential([0.06, 0.09], 1.0, 2), 2),
                      ("matern_5d", kernels.KernelIsoMatern(1.0, 1.0, 5), 5),
                      ("se_10d", kernels.KernelSquaredExponential(list(np.linspace(0.5, 1.5, 10)), 1.0, 10), 10)]:
    kern._bind(dev)
    C = 100_000
    X = dev.points(rng.uniform(-1, 1, (C, d)))
    nx = 4096
    G = buf[: nx * X.ld].view(nx, X.ld)
    t = ev_ms(lambda: check(lib.gpx_gram(dev.h, ptr(X.X), nx, X.ld, ptr(X.X), X.n, X.ld, ptr(G), X.ld, 0, None, 0.0, dev.stream)), 5)
    out[f"gram_{name}_gbs"] = 8.0 * nx * C / t / 1e6
    # the same block as a tensor-pipe Gram: DMMA prologue (expanded form) + store epilogue, no contraction (n = 0)
    from gpexp_b200.device import prologue_operands
    R = dev.points(rng.uniform(-1, 1, (nx, d)))
    dev.set_center(X.midrange())
    mode, ra, rb = prologue_operands(R, X)
    t = ev_ms(lambda: check(lib.gpx_cov_from_factors(dev.h, mode, None, R.ld, ptr(ra), nx, None, X.ld, ptr(rb), C, 0, ptr(G), X.ld,
                                                     dev.stream)), 5)
    out[f"gram_dmma_{name}_gbs"] = 8.0 * nx * C / t / 1e6
    out[f"gram_dmma_{name}_mode"] = int(mode)
    if once and name != "se_2d":
        continue
    # fused Gram/TRSM: W = U^-T K(D, X) for a 1024-point design (potrf + dmma_core<STORE> + tri_solve)
    n = 1024
    design = rng.uniform(-1, 1, (n, d))
    f = DesignFactor(dev, dev.points(design), 1e-4)
    t = ev_ms(lambda: DesignFactor(dev, dev.points(design), 1e-4), 3)
    out[f"potrf_{name}_n{n}_ms"] = t
    out[f"potrf_{name}_n{n}_tflops"] = n ** 3 / 3.0 / t / 1e9
    W = dev.zeros(n, X.ld)
    t = ev_ms(lambda: f.solve_gram(X, W=W), 3)
    out[f"trsm_gram_{name}_n{n}_ms"] = t
    out[f"trsm_gram_{name}_n{n}_tflops"] = n * n * C / t / 1e9
    # HBM-bound incremental row append at n = 1023 (reads 8*n*C bytes)
    var = dev.zeros(X.ld)
    rec = dev.zeros(19 + n)
    rec[2] = 1.0
    t = ev_ms(lambda: check(lib.gpx_append_row(dev.h, 0, ptr(rec), None, ptr(X.X), C, X.ld, ptr(W), X.ld, n - 1, ptr(var), dev.stream)), 10)
    out[f"append_{name}_n{n - 1}_gbs"] = 8.0 * (n + 1) * C / t / 1e6
    t = ev_ms(lambda: check(lib.gpx_append_row(dev.h, 0, ptr(rec), None, ptr(X.X), C, X.ld, ptr(W), X.ld, 255, ptr(var), dev.stream)), 20)
    out[f"append_{name}_n255_gbs"] = 8.0 * 257 * C / t / 1e6
    out[f"append_{name}_n255_us"] = t * 1e3
    del W, f
print(json.dumps(out, indent=1))
if not once:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/kernels_r02.json", "w"), indent=1)
