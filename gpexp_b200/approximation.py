"""The `Space` record the cost functions consume (API of gpExp/approximation.py:22-36): input dimension, a sampler
`sample(shape)`, a density `probDensity(points)` and an optional heteroscedastic noise function `noiseFunc(points)`."""
from typing import Callable, Optional


class Space:
    def __init__(self, dimensionIn: int, samplerIn: Optional[Callable], probDensityIn: Optional[Callable],
                 noise: Optional[Callable] = None):
        # positional order and attribute names are the reference's; nothing else lives here
        self.dimension, self.sample, self.probDensity, self.noiseFunc = dimensionIn, samplerIn, probDensityIn, noise
        self.inBoundsBool = None

    def __repr__(self):
        return "Space(dimension=%r, noiseFunc=%s)" % (self.dimension, "set" if self.noiseFunc is not None else "None")
