"""Generate tests/golden/resampler_tiny.npz from the UNMODIFIED reference Resampler (/root/reference) — run in the
authoring container:   python tests/golden/make_golden_resampler.py

Same recipe as make_golden.py: the reference class is imported as is, loaded with the seeded synthetic weights of
tooncrafter_b200/synthetic.py and run on CPU in fp32; only its OUTPUT is stored (weights / inputs are regenerated from
seeds by the tests).  Also writes the full-size key manifest of the Resampler of configs/inference_512_v1.0.yaml.
"""
import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))

from tooncrafter_b200 import synthetic  # noqa: E402

TINY = dict(dim=256, depth=2, dim_head=64, heads=4, num_queries=4, embedding_dim=128, output_dim=256, ff_mult=4,
            video_length=4)
FULL = dict(dim=1024, depth=4, dim_head=64, heads=12, num_queries=16, embedding_dim=1280, output_dim=1024, ff_mult=4,
            video_length=16)                      # configs/inference_512_v1.0.yaml: image_proj_stage_config
SEED = 7


def tiny_input():
    return torch.randn(2, 33, TINY["embedding_dim"], generator=synthetic._gen("resampler.x", 123))


def reference_class():
    spec = importlib.util.spec_from_file_location("_ref_resampler", "/root/reference/lvdm/modules/encoders/resampler.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Resampler


def main():
    Ref = reference_class()
    torch.manual_seed(0)
    m = Ref(**TINY).eval()
    synthetic.fill_module_(m, seed=SEED, prefix="image_proj_model.")
    with torch.no_grad():
        y = m(tiny_input())
    np.savez_compressed(HERE / "resampler_tiny.npz", out=y.numpy())
    with torch.device("meta"):
        full = Ref(**FULL)
    manifest = {k: list(v.shape) for k, v in full.state_dict().items()}
    (HERE / "state_dict_manifest_resampler.json").write_text(json.dumps(manifest, indent=0, sort_keys=True))
    print("resampler golden:", tuple(y.shape), float(y.abs().max()), "| full-size keys:", len(manifest))


if __name__ == "__main__":
    main()
