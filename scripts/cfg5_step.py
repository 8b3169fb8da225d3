"""BASELINE cfg-5: one greedy IVAR step, 10-D ARD squared-exponential, n = 4096 design points,
1 000 000 candidates x 100 000 integration points, candidates sharded over the ranks (strong scaling).

    python scripts/cfg5_step.py [--cands 1000000] [--steps 2] [--check]                      # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/cfg5_step.py [...]        # 8 GPUs

Prints one JSON line (rank 0) and writes gpurun_out/cfg5_N<world>.json."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--cands", type=int, default=1_000_000)
ap.add_argument("--mc", type=int, default=100_000)
ap.add_argument("--npts", type=int, default=4096)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--check", action="store_true")
ap.add_argument("--greedy", action="store_true",
                help="grow the n-point design greedily from the candidates with the resident-covariance engine (the real "
                     "cfg-5 design) instead of loading a random design; then cross-check the DMMA step on that state")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from gpexp_b200 import kernels  # noqa: E402
from gpexp_b200.device import Device  # noqa: E402
from gpexp_b200.engine import DesignFactor, GreedyIVAREngine, Shard, prior_scale  # noqa: E402

d, n, C, M, noise = 10, args.npts, args.cands, args.mc, 1e-6
rng = np.random.default_rng(5)
cl = list(np.linspace(0.5, 1.5, d))
cand_h = rng.uniform(-1, 1, (C, d))
mc_h = rng.uniform(-1, 1, (M, d))
design_idx = np.sort(rng.permutation(C)[:n])
design_h = cand_h[design_idx]
dev = Device.get(local)
kern = kernels.KernelSquaredExponential(cl, 1.0, d)
kern._bind(dev)
fam, _, params = kern._gpx_spec()
shard = Shard() if world > 1 else None
lo, hi = Shard.split(C, world, rank)
cand, mc = dev.points(cand_h[lo:hi]), dev.points(mc_h)
eng = GreedyIVAREngine(dev, cand, mc, n + 1, noise, prior_scale(fam, params), shard=shard, index_offset=lo,
                       resident=args.greedy)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


greedy = None
if args.greedy:
    sync()
    t0 = time.perf_counter()
    eng.run(n)
    sync()
    greedy_s = time.perf_counter() - t0
    design_idx = eng.indices()[:n]
    design_h = cand_h[design_idx]
    # next pick by the resident path, then by the DMMA contraction on the very same state
    eng.score()
    res_pick = torch.stack([eng.best.clone(), eng.idx.double() + lo])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.score(contraction=True)
    sync()
    e0.record()
    eng.score(contraction=True)
    e1.record()
    sync()
    dm_ms = e0.elapsed_time(e1)
    dm_pick = torch.stack([eng.best.clone(), eng.idx.double() + lo])
    both = torch.cat([res_pick.flatten(), dm_pick.flatten()])
    if world > 1:
        allp = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(allp, both)
        allp = torch.stack(allp).cpu().numpy()
    else:
        allp = both.cpu().numpy()[None, :]
    r_best = allp[np.lexsort((allp[:, 1], allp[:, 0]))[0]]
    d_best = allp[np.lexsort((allp[:, 3], allp[:, 2]))[0]]
    greedy = {"greedy_design_s": greedy_s, "greedy_ms_per_step": 1e3 * greedy_s / n,
              "resident_hbm_gbs_per_gpu": 16.0 * M * (hi - lo) / (greedy_s / n) / 1e9,
              "next_pick_resident": [float(r_best[0]), int(r_best[1])], "next_pick_dmma": [float(d_best[2]), int(d_best[3])],
              "dmma_step_ms_on_greedy_design": dm_ms, "dmma_tflops_per_gpu": 2.0 * M * n * (hi - lo) / dm_ms / 1e9,
              "first_picks": [int(i) for i in design_idx[:8]]}
    if rank == 0:
        print(json.dumps({"workload": "cfg-5 greedy design grown with the resident covariance", "n_gpus": world, "n": n, "C": C,
                          "M": M, **greedy}), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(greedy, open("gpurun_out/cfg5_greedy_N%d.json" % world, "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)

sync()
t0 = time.perf_counter()
eng.load_design(DesignFactor(dev, dev.points(design_h), noise))
sync()
setup_s = time.perf_counter() - t0
snap = eng.snapshot()
eng.score()
eng.append()
eng.restore(snap)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
e0.record()
for s in range(args.steps):
    ks[s][0].record()
    eng.score()
    ks[s][1].record()
    eng.append()
    if s + 1 < args.steps:
        eng.restore(snap)
e1.record()
sync()
ms = e0.elapsed_time(e1) / args.steps
score_ms = float(np.mean([a.elapsed_time(b) for a, b in ks]))
if world > 1:
    t = torch.tensor([ms, score_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, score_ms = float(t[0]), float(t[1])
winner = int(eng.picks[n].item())
wscore = float(eng.pick_scores[n].item())
if rank == 0:
    flops_gpu = 2.0 * M * n * (hi - lo)
    out = {"workload": "cfg-5: 10-D ARD SE, one greedy IVAR step at n=%d, %d candidates x %d MC points" % (n, C, M),
           "n_gpus": world, "ms_per_step": ms, "score_ms": score_ms, "candidates_per_s": C / ms * 1e3,
           "tflops_per_gpu": flops_gpu / score_ms / 1e9, "frac_of_dmma_peak_37.2": flops_gpu / score_ms / 1e9 / 37.2,
           "setup_from_scratch_s": setup_s, "winner_global_index": winner, "winner_cost": wscore,
           "total_flops_per_step": 2.0 * M * n * C}
    if args.check:
        sys.path.insert(0, ROOT)
        from oracle import gpexp_oracle as orc
        okern = orc.KernelSpec.se(cl, 1.0, d)
        sub = np.unique(np.concatenate([[winner], rng.permutation(C)[:255]]))
        w_m, var_m = orc.fast_design_state(okern, design_h, mc_h, noise)
        w_c, var_c = orc.fast_design_state(okern, design_h, cand_h[sub], noise)
        ref = orc.fast_ivar_scores(okern, cand_h[sub], mc_h, w_m, var_m, w_c, var_c, noise)
        wi = int(np.where(sub == winner)[0][0])
        out["oracle_winner_cost_rel_err"] = abs(ref[wi] - wscore) / abs(ref[wi])
        out["oracle_winner_is_min_of_sample"] = bool(ref[wi] == ref.min())
        local_mask = (sub >= lo) & (sub < hi)
        got = eng.scores[: cand.n].cpu().numpy()[sub[local_mask] - lo]
        out["oracle_max_rel_err_local_sample"] = float(np.max(np.abs(got - ref[local_mask]) / np.abs(ref[local_mask])))
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/cfg5_N%d.json" % world, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
