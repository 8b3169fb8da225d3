"""gpexp_b200 -- B200-native (sm_100a CUDA) implementation of the greedy experimental-design hot path
of goroda/GPEXP behind the reference's own Python API:

    gpexp_b200.kernels              <->  gpExp/kernels.py
    gpexp_b200.gp                   <->  gpExp/gp.py
    gpexp_b200.gp_kernel_utilities  <->  gpExp/gp_kernel_utilities.py  (calculateCovarianceMatrix)
    gpexp_b200.experimentalDesign   <->  gpExp/experimentalDesign.py   (cost functions + greedy drivers)
    gpexp_b200.approximation        <->  gpExp/approximation.py        (Space)

`install_as_gpExp()` registers these modules under the name `gpExp`, so existing scripts that
`import gpExp.kernels` run on the GPU path unchanged.  There is no CPU fallback.
"""
import sys

__version__ = "0.1.0"


def install_as_gpExp():
    """Alias this package as `gpExp` (drop-in for scripts written against the reference)."""
    from . import approximation, experimentalDesign, gp, gp_kernel_utilities, kernels
    pkg = sys.modules[__name__]
    sys.modules["gpExp"] = pkg
    for name, mod in [("kernels", kernels), ("gp", gp), ("gp_kernel_utilities", gp_kernel_utilities),
                      ("experimentalDesign", experimentalDesign), ("approximation", approximation)]:
        sys.modules["gpExp." + name] = mod
    return pkg
