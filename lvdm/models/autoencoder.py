"""Alias of lvdm/models/autoencoder.py (reference :13, :238)."""
from tooncrafter_b200.diffusion import AutoencoderKL, AutoencoderKL_Dualref  # noqa: F401
