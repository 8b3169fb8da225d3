"""Space record consumed by the cost functions (mirrors gpExp/approximation.py:22-36)."""


class Space:
    """Describes the input space: dimension, a sampler, a density and an optional noise function."""

    dimension = None
    inBoundsBool = None
    sample = None
    probDensity = None
    noiseFunc = None

    def __init__(self, dimensionIn, samplerIn, probDensityIn, noise=None):
        self.dimension = dimensionIn
        self.sample = samplerIn
        self.probDensity = probDensityIn
        self.noiseFunc = noise
