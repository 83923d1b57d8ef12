"""Alias of lvdm/models/samplers/ddim_multiplecond.py (reference :10): 3-way guidance sampler."""
from tooncrafter_b200.sampler import DDIMSamplerMultiCond as DDIMSampler  # noqa: F401
