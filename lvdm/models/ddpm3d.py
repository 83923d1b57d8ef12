"""Alias of lvdm/models/ddpm3d.py's inference classes (reference :41, :465, :1041, :1243)."""
from tooncrafter_b200.diffusion import (DiffusionWrapper, LatentDiffusion,  # noqa: F401
                                        LatentVisualDiffusion)
