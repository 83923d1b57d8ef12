"""Golden for the reference-glue integration test: the UNMODIFIED reference's scripts/evaluation/funcs.py::
batch_ddim_sampling driving the UNMODIFIED reference model (tiny configuration, CPU fp32) — run in the authoring container:

    python tests/golden/make_golden_glue.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))

from oracle import ref_shims  # noqa: E402
from tiny_config import TINY_CONTEXT_DIM, TINY_LATENT_HW, TINY_T, model_config  # noqa: E402
from tooncrafter_b200 import synthetic  # noqa: E402
import glue_common  # noqa: E402

STRIDE = 5


def main():
    torch.set_num_threads(8)
    model = ref_shims.build_reference_model(model_config()).eval()
    synthetic.fill_module_(model, seed=0)
    model.perframe_ae = True
    model.temporal_length = TINY_T
    # harness shim 3: the reference sampler hard-codes "cuda" in register_buffer (ddim.py:18-22)
    import lvdm.models.samplers.ddim as ref_ddim

    def register_buffer(self, name, attr):
        if isinstance(attr, torch.Tensor):
            attr = attr.to(self.model.device)
        setattr(self, name, attr)
    ref_ddim.DDIMSampler.register_buffer = register_buffer
    funcs = glue_common.load_reference_funcs()
    assert funcs.DDIMSampler is ref_ddim.DDIMSampler
    outs = glue_common.run_glue(funcs, model, TINY_T, *TINY_LATENT_HW, TINY_CONTEXT_DIM)
    save = {}
    for i, o in enumerate(outs):
        save[f"clip{i}_sub"] = o.flatten()[::STRIDE].numpy()
        save[f"clip{i}_shape"] = np.array(o.shape)
        print(i, tuple(o.shape), float(o.abs().max()))
    np.savez_compressed(HERE / "glue_tiny.npz", **save)


if __name__ == "__main__":
    main()
