import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gpexp_b200 import kernels
from gpexp_b200.device import Device
from gpexp_b200.engine import GreedyMIEngine, ShardedMIEngine
dev = Device.get(0)
k = kernels.KernelMehlerND([0.9, 0.9, 0.9], 3); k._bind(dev)
for V in (20000, 40000):
    pool = np.random.default_rng(4).standard_normal((V, 3))
    for name, mk in (("dense", lambda: GreedyMIEngine(dev, dev.points(pool), 32, 1e-2)), ("sharded1", lambda: ShardedMIEngine(dev, pool, 32, 1e-2))):
        torch.cuda.synchronize(); t0 = time.perf_counter(); e = mk(); torch.cuda.synchronize(); ts = time.perf_counter() - t0
        t0 = time.perf_counter(); idx = e.run(32); torch.cuda.synchronize(); tr = time.perf_counter() - t0
        print(V, name, "setup %.3f s" % ts, "32 steps %.3f s" % tr, [int(i) for i in idx[:6]], flush=True)
        del e; torch.cuda.empty_cache()
