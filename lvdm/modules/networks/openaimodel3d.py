"""Alias of lvdm/modules/networks/openaimodel3d.py (reference :281)."""
from tooncrafter_b200.modules import UNetModel  # noqa: F401
