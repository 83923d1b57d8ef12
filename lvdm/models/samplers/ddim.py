"""Alias of lvdm/models/samplers/ddim.py (reference :10)."""
from tooncrafter_b200.sampler import DDIMSampler  # noqa: F401
