"""Shared by tests/golden/make_golden_glue.py (drives the UNMODIFIED reference) and tests/glue_driver.py (drives this
repository's alias tree): the reference's own glue function scripts/evaluation/funcs.py::batch_ddim_sampling (:14-93) is
loaded from /root/reference BY FILE and run unchanged on the tiny configuration.

Conditioning stages (CLIP text / image towers, Resampler) are outside the hot path: both sides install the same seeded
test doubles for `get_learned_conditioning`, `embedder` and `image_proj_model`."""
import importlib.util
import sys
import types
from pathlib import Path

import torch

REF_FUNCS = Path("/root/reference/scripts/evaluation/funcs.py")
SEED_RNG = 77
STEPS = 4


def stub_io_modules():
    """decord / cv2 are video-I/O imports at the top of funcs.py (absent here, never called on this path)."""
    for name in ("decord", "cv2"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.VideoReader = m.cpu = None
            sys.modules[name] = m


def load_reference_funcs():
    stub_io_modules()
    spec = importlib.util.spec_from_file_location("reference_eval_funcs", REF_FUNCS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Fn(torch.nn.Module):
    """A parameter-free nn.Module around a callable (the conditioning stages are registered sub-modules)."""

    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def forward(self, *a, **k):
        return self._fn(*a, **k)


def install_conditioning_doubles(model, T, ctx_dim):
    from tooncrafter_b200 import synthetic
    g = lambda n: synthetic._gen(n, 555)
    txt_empty = torch.randn(1, 77, ctx_dim, generator=g("glue.txt.empty"))
    img_zero = torch.randn(1, 16 * T, ctx_dim, generator=g("glue.img.zero"))
    model.get_learned_conditioning = lambda prompts: txt_empty.expand(len(prompts), -1, -1).clone()
    model.embedder = _Fn(lambda img: img.new_zeros(img.shape[0], 4, ctx_dim))
    model.image_proj_model = _Fn(lambda tok: img_zero.expand(tok.shape[0], -1, -1).clone())


def glue_inputs(T, h, w, ctx_dim):
    from tooncrafter_b200 import synthetic
    g = lambda n: synthetic._gen(n, 556)
    frames = torch.rand(2, 3, 8 * h, 8 * w, generator=g("glue.frames")) * 2 - 1
    z = torch.randn(1, 4, T, h, w, generator=g("glue.z")) * 0.18215 * 3
    cc = torch.zeros_like(z)
    cc[:, :, 0], cc[:, :, -1] = z[:, :, 0], z[:, :, -1]
    prompts = [torch.randn(1, 77 + 16 * T, ctx_dim, generator=g(f"glue.ctx.{i}")) for i in range(2)]
    mask = (torch.rand(1, 1, T, h, w, generator=g("glue.mask")) > 0.5).float()
    x0 = torch.randn(1, 4, T, h, w, generator=g("glue.x0"))
    return dict(frames=frames, c_concat=cc, prompts=prompts, mask=mask, x0=x0, fs=torch.tensor([10]))


def run_glue(funcs, model, T, h, w, ctx_dim):
    """Two prompts back to back (a stale conditioning cache would show in the second), then prompt 0 again with the
    mask / x0 blending kwargs (ddim.py:174-180).  Returns the three decoded clips."""
    gi = glue_inputs(T, h, w, ctx_dim)
    install_conditioning_doubles(model, T, ctx_dim)
    with torch.no_grad():
        post, hidden = model.first_stage_model.encode(gi["frames"], return_hidden_states=True)
        hs = [hh.reshape(1, 2, *hh.shape[1:]).permute(0, 2, 1, 3, 4).contiguous().float() for hh in hidden]
        outs = []
        for k, extra in ((0, {}), (1, {}), (0, dict(mask=gi["mask"], x0=gi["x0"]))):
            cond = {"c_crossattn": [gi["prompts"][k].clone()], "c_concat": [gi["c_concat"]], "fs": gi["fs"]}
            torch.manual_seed(SEED_RNG)
            v = funcs.batch_ddim_sampling(model, cond, [1, 4, T, h, w], n_samples=1, ddim_steps=STEPS, ddim_eta=1.0,
                                          cfg_scale=7.5, hs=hs, **extra)
            outs.append(v.float())
    return outs
