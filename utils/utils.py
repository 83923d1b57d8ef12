"""Alias of the reference's utils/utils.py.  With this repository ahead of the reference on PYTHONPATH this module
shadows the reference's, so it carries the WHOLE public surface of that file, not only the plugin seam:

  instantiate_from_config / get_obj_from_str   utils/utils.py:27-42  (config-string instantiation, the plugin seam)
  count_params                                 :8-12   (imported by lvdm/modules/encoders/condition.py)
  check_istarget                               :15-24
  load_npz_from_dir / load_npz_from_paths      :45-54
  resize_numpy_image                           :57-67  (needs cv2, imported lazily: cv2 is optional on the hot path)
  setup_dist                                   :70-77

Host-side helpers only; no arithmetic of the hot path lives here.
"""
import os

import numpy as np

from tooncrafter_b200.diffusion import get_obj_from_str, instantiate_from_config  # noqa: F401


def count_params(model, verbose=False):
    n = sum(p.numel() for p in model.parameters())
    if verbose:
        print(f"{model.__class__.__name__} has {n * 1.e-6:.2f} M params.")
    return n


def check_istarget(name, para_list):
    """True when any of the partial names in `para_list` occurs in the full parameter name `name`."""
    return any(part in name for part in para_list)


def _load_arr0(paths):
    return np.concatenate([np.load(p)["arr_0"] for p in paths], axis=0)


def load_npz_from_dir(data_dir):
    return _load_arr0([os.path.join(data_dir, n) for n in os.listdir(data_dir)])


def load_npz_from_paths(data_paths):
    return _load_arr0(list(data_paths))


def resize_numpy_image(image, max_resolution=512 * 512, resize_short_edge=None):
    """Resize to a multiple of 64 per side, either to a short-edge length or to a pixel budget (Lanczos)."""
    import cv2
    h, w = image.shape[:2]
    k = resize_short_edge / min(h, w) if resize_short_edge is not None else (max_resolution / (h * w)) ** 0.5
    h, w = int(np.round(h * k / 64)) * 64, int(np.round(w * k / 64)) * 64
    return cv2.resize(image, (w, h), interpolation=cv2.INTER_LANCZOS4)


def setup_dist(args):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    torch.cuda.set_device(args.local_rank)
    dist.init_process_group("nccl", init_method="env://")
