"""Named kernel configurations shared by the golden generator, the oracle tests and the GPU tests."""
import numpy as np

from oracle.gpexp_oracle import KernelSpec

_DEFS = {
    "se_iso_1d": ("se", [0.05], 1.0, 1),
    "se_ard_2d": ("se", [0.06, 0.09], 1.0, 2),
    "se_ard_2d_wide": ("se", [0.3, 0.45], 1.7, 2),
    "se_ard_10d": ("se", list(np.linspace(0.5, 1.5, 10)), 1.0, 10),
    "matern_5d": ("matern", 1.0, 1.0, 5),
    "matern_5d_b": ("matern", 0.5, 2.5, 5),
    "mehler_3d": ("mehler", [0.9, 0.9, 0.9], None, 3),
    "mehler_3d_b": ("mehler", [0.5, 0.7, 0.3], None, 3),
    "mehler_1d": ("mehler1d", 0.6, None, 1),
}

KERNEL_NAMES = list(_DEFS)


def spec(name) -> KernelSpec:
    fam, a, b, d = _DEFS[str(name)]
    if fam == "se":
        return KernelSpec.se(a, b, d)
    if fam == "matern":
        return KernelSpec.matern32(a, b, d)
    if fam == "mehler":
        return KernelSpec.mehler(a, d)
    return KernelSpec.mehler([a], 1)


def product_kernel(name):
    """The same configuration as a gpexp_b200.kernels object (the drop-in API)."""
    from gpexp_b200 import kernels as K
    fam, a, b, d = _DEFS[str(name)]
    if fam == "se":
        return K.KernelSquaredExponential(list(a), b, d)
    if fam == "matern":
        return K.KernelIsoMatern(a, b, d)
    if fam == "mehler":
        return K.KernelMehlerND(list(a), d)
    return K.KernelMehler1D(a, 1)
