"""Alias of utils/utils.py:27-42 (config-string instantiation = the reference's plugin seam)."""
from tooncrafter_b200.diffusion import get_obj_from_str, instantiate_from_config  # noqa: F401
