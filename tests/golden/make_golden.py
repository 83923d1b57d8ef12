"""Generate tests/golden/* from the UNMODIFIED reference (/root/reference) — run in the authoring container:

    python tests/golden/make_golden.py

The reference is imported through oracle/ref_shims.py (3 harness shims, nothing edited), its modules are loaded
with the seeded synthetic weights of tooncrafter_b200/synthetic.py (strict=True), and driven on the tiny
configuration of tests/tiny_config.py on CPU in fp32.  Weights and inputs are regenerated from seeds by the
tests, so only the reference's OUTPUTS are stored.  The full-size state-dict manifest (keys + shapes of the real
inference_512_v1.0 model, built on the meta device) pins checkpoint compatibility.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE.parent))

from oracle import ref_shims  # noqa: E402
from tiny_config import (FULL_DDCONFIG, FULL_UNET, TINY_CONTEXT_DIM, TINY_DDCONFIG, TINY_LATENT_HW, TINY_T,  # noqa: E402
                         TINY_UNET, model_config)
from tooncrafter_b200 import synthetic  # noqa: E402

SEED = 0


def golden_inputs():
    """Deterministic inputs shared by make_golden.py and the tests."""
    T = TINY_T
    h, w = TINY_LATENT_HW
    g = lambda name: synthetic._gen(name, 123)
    unet_in = dict(
        x=torch.randn(2, 8, T, h, w, generator=g("unet.x")),
        t=torch.tensor([999, 499]),
        ctx=torch.randn(2, 77 + 16 * T, TINY_CONTEXT_DIM, generator=g("unet.ctx")),
        fs=torch.tensor([10, 24]))
    frames = torch.rand(2, 3, 8 * h, 8 * w, generator=g("vae.frames")) * 2 - 1
    z = torch.randn(1, 4, T, h, w, generator=g("vae.z")) * 0.18215 * 3
    x_T, cond, uncond = synthetic.synthetic_inputs(1, T, h, w, TINY_CONTEXT_DIM, seed=123)
    S = 4
    noises = [torch.randn(x_T.shape, generator=g(f"ddim.noise.{i}")) for i in range(S)]
    return dict(unet=unet_in, frames=frames, z=z, x_T=x_T, cond=cond, uncond=uncond, S=S, noises=noises,
                fs=torch.tensor([10]))


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = ref_shims.build_reference_model(model_config()).eval()
    synthetic.fill_module_(model, seed=SEED)
    model.perframe_ae = True
    model.temporal_length = TINY_T
    gi = golden_inputs()
    out = {}
    with torch.no_grad():
        u = gi["unet"]
        out["unet_y"] = model.model.diffusion_model(u["x"], u["t"], context=u["ctx"], fs=u["fs"]).numpy()

        post, hidden = model.first_stage_model.encode(gi["frames"], return_hidden_states=True)
        out["enc_moments"] = post.parameters.numpy()
        for i, hsm in enumerate(hidden):
            flat = hsm.flatten()
            out[f"enc_hidden{i}_sub"] = flat[::97].numpy()           # strided subsample + moments
            out[f"enc_hidden{i}_stats"] = np.array([flat.mean().item(), flat.std().item(), flat.abs().max().item()])
        ref_ctx = [hh.reshape(1, 2, *hh.shape[1:]).permute(0, 2, 1, 3, 4).contiguous() for hh in hidden]
        out["decode"] = model.decode_first_stage(gi["z"], ref_context=ref_ctx).numpy()

        # DDIM: drive the reference sampler with teacher-forced noise by patching torch.randn via the generator
        # order: the sampler draws exactly one randn(shape) per step (ddim.py:273 -> common.py:31-34)
        sampler = ref_shims.make_sampler(model)
        import lvdm.models.samplers.ddim as ref_ddim
        it = iter(gi["noises"])
        orig = ref_ddim.noise_like
        ref_ddim.noise_like = lambda shape, device, repeat=False: next(it)
        try:
            samples, _ = sampler.sample(S=gi["S"], batch_size=1, shape=list(gi["x_T"].shape[1:]),
                                        conditioning=gi["cond"], unconditional_conditioning=gi["uncond"], eta=1.0,
                                        unconditional_guidance_scale=7.5, x_T=gi["x_T"], fs=gi["fs"],
                                        timestep_spacing="uniform_trailing", guidance_rescale=0.7, verbose=False)
        finally:
            ref_ddim.noise_like = orig
        out["ddim_samples"] = samples.numpy()

        # schedule known answers straight from the reference buffers / tables
        sampler.make_schedule(ddim_num_steps=50, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
        kat = dict(
            ddim_timesteps_50=[int(v) for v in sampler.ddim_timesteps],
            alphas_cumprod_0=float(model.alphas_cumprod[0]), alphas_cumprod_last=float(model.alphas_cumprod[-1]),
            a_prev_last=float(sampler.ddim_alphas_prev[-1]), sigma_last=float(sampler.ddim_sigmas[-1]),
            scale_arr_999=float(model.scale_arr[999]), scale_arr_19=float(model.scale_arr[19]),
            scale_arr_len=int(model.scale_arr.shape[0]))
        sampler.make_schedule(ddim_num_steps=10, ddim_discretize="uniform_trailing", ddim_eta=1.0, verbose=False)
        kat["ddim_timesteps_10"] = [int(v) for v in sampler.ddim_timesteps]
        for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "scale_arr"):
            out["sched_" + k] = getattr(model, k).numpy()
    np.savez_compressed(HERE / "tiny_reference_outputs.npz", **out)
    (HERE / "schedule_kat.json").write_text(json.dumps(kat, indent=1))

    # full-size manifest: UNet and VAE built on the meta device (the wrapper's schedule code needs numpy, so the
    # 13 top-level schedule buffers are taken from the tiny model: their shapes do not depend on the width)
    with torch.device("meta"):
        unet = ref_shims.build_reference_unet(FULL_UNET)
        vae = ref_shims.build_reference_vae(FULL_DDCONFIG)
    manifest = {k: list(v.shape) for k, v in model.state_dict().items()
                if not k.startswith(("model.", "first_stage_model."))}
    manifest.update({"model.diffusion_model." + k: list(v.shape) for k, v in unet.state_dict().items()})
    manifest.update({"first_stage_model." + k: list(v.shape) for k, v in vae.state_dict().items()})
    (HERE / "state_dict_manifest_512.json").write_text(json.dumps(manifest, indent=0))
    tiny_manifest = {k: list(v.shape) for k, v in model.state_dict().items()}
    (HERE / "state_dict_manifest_tiny.json").write_text(json.dumps(tiny_manifest, indent=0))
    print("wrote", [p.name for p in HERE.iterdir()])
    for k, v in out.items():
        print(k, getattr(v, "shape", None), float(np.abs(v).max()))


if __name__ == "__main__":
    main()
