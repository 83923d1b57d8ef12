"""Alias of lvdm/distributions.py (reference :24)."""
from tooncrafter_b200.diffusion import DiagonalGaussianDistribution  # noqa: F401
