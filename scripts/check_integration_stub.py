"""Executes the ctypes stub of INTEGRATION.md section 2 verbatim (extracted from the markdown) on cuda:0 and compares its
calculateCovarianceMatrix with the closed form.  Run from the repository root."""
import re, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
src = open("INTEGRATION.md").read()
block = re.search(r"```python\nimport ctypes as C, numpy as np, torch\n(.*?)```", src, re.S).group(0)
code = block.strip("`").replace("python\n", "", 1).replace('C.CDLL("libgpexp_b200.so")', 'C.CDLL(os.path.join(os.getcwd(), "gpexp_b200/lib/libgpexp_b200.so"))')
ns = {"os": os}
exec(code, ns)
class K:  # minimal stand-in with the reference's hyperParam dict
    hyperParam = {"cl0": 0.3, "cl1": 0.7, "signalSize": 1.3}
pts = np.random.default_rng(0).uniform(-1, 1, (300, 2))
got = ns["calculateCovarianceMatrix"](K(), pts, 1e-6)
d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2 / np.array([0.3, 0.7]) ** 2).sum(-1)
ref = 1.3 * np.exp(-0.5 * d2) + 1e-6 * np.eye(300)
print("max abs err", np.abs(got - ref).max())
assert np.abs(got - ref).max() < 1e-12
print("INTEGRATION stub OK")
