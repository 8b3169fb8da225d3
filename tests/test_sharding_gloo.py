"""world_size-2 (and 3) gloo test of the candidate-sharding protocol on CPU.

The device engines cannot run without a GPU, so this test drives the *same protocol* with numpy in
place of the CUDA kernels: contiguous blocks from gpexp_b200.engine.Shard.split, one pivot record per
rank with the layout of include/gpexp_b200.h (score, global index, var+noise, x_p[16], W[0:n,p]),
ONE all_gather per step through gpexp_b200.engine.Shard.all_gather, and the selection rule of
gpx_select_pivot (better score, ties to the lowest global index).  The picks must equal the
single-process oracle, including the all-tied first step of a stationary kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gpexp_oracle as orc

HDR = 3 + 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _select(recs, minimize):
    """numpy mirror of select_pivot_kernel (gpexp_b200/csrc/gpx_basic.cu)."""
    bv, bi, bw = 0.0, -1, 0
    for r, rec in enumerate(recs):
        v, i = rec[0], int(rec[1])
        if i < 0:
            continue
        better = bi < 0 or (v < bv if minimize else v > bv) or (v == bv and i < bi)
        if better:
            bv, bi, bw = v, i, r
    return recs[bw]


def _worker(rank, world, port, pool, mc, n_points, noise, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpexp_b200.engine import Shard
    shard = Shard()
    kern = orc.KernelSpec.se([0.3, 0.45], 1.7, 2)
    lo, hi = Shard.split(len(pool), world, rank)
    local = pool[lo:hi]
    reclen = HDR + n_points

    def exchange(score, j, var_plus_noise, col, n, minimize):
        rec = np.zeros(reclen)
        rec[0], rec[1], rec[2] = score, (lo + j if j >= 0 else -1), var_plus_noise
        if j >= 0:
            rec[3:3 + 2] = local[j]
            rec[HDR:HDR + n] = col
        out = torch.zeros(world * reclen, dtype=torch.float64)
        shard.all_gather(out, torch.from_numpy(rec))
        return _select(out.numpy().reshape(world, reclen), minimize)

    # ---- greedy max variance (experimentalDesign.py:787-845 restated incrementally) ----------------
    var = kern.prior(local).copy()
    W = np.zeros((n_points, hi - lo))
    vpicks = []
    for n in range(n_points):
        j = int(np.argmax(var)) if hi > lo else -1
        win = exchange(var[j] if j >= 0 else 0.0, j, var[j] if j >= 0 else 1.0, W[:n, j] if j >= 0 else None, n, False)
        row = (kern.gram(win[3:5][None, :], local)[0] - win[HDR:HDR + n] @ W[:n]) / np.sqrt(win[2])
        W[n] = row
        var -= row * row
        vpicks.append(int(win[1]))

    # ---- greedy IVAR: candidates sharded, integration points replicated ------------------------------
    var_c, var_m = kern.prior(local).copy(), kern.prior(mc).copy()
    Wc, Wm = np.zeros((n_points, hi - lo)), np.zeros((n_points, len(mc)))
    ipicks = []
    for n in range(n_points):
        cost = orc.fast_ivar_scores(kern, local, mc, Wm[:n], var_m, Wc[:n], var_c, noise)
        j = int(np.argmin(cost))
        win = exchange(cost[j], j, var_c[j] + noise, Wc[:n, j], n, True)
        lnn, x, col = np.sqrt(win[2]), win[3:5][None, :], win[HDR:HDR + n]
        rc = (kern.gram(x, local)[0] - col @ Wc[:n]) / lnn
        rm = (kern.gram(x, mc)[0] - col @ Wm[:n]) / lnn
        Wc[n], Wm[n] = rc, rm
        var_c -= rc * rc
        var_m -= rm * rm
        ipicks.append(int(win[1]))
    q.put((rank, vpicks, ipicks))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_matches_single_process_oracle(world):
    rng = np.random.default_rng(17)
    pool, mc = rng.uniform(-1, 1, (101, 2)), rng.uniform(-1, 1, (300, 2))
    pool[60] = pool[10]  # an exact duplicate living on another rank: the tie must go to the lower global index
    n_points, noise = 9, 1e-6
    kern = orc.KernelSpec.se([0.3, 0.45], 1.7, 2)
    vref, _ = orc.fast_greedy_var(kern, pool, n_points)
    iref, _ = orc.fast_greedy_ivar(kern, pool, mc, n_points, noise)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, pool, mc, n_points, noise, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, vp, ip in results:
        assert vp == vref, (rank, vp, vref)
        assert ip == iref, (rank, ip, iref)
    assert vref[0] == 0  # stationary kernel: the first step is an all-way tie and index 0 must win across ranks
