"""Width/size-reduced configurations of the reference model used by CPU tests and golden fixtures.

Same topology as configs/inference_512_v1.0.yaml (4 levels, 2 res blocks, attention at ds 1/2/4, init_attn,
temporal conv, image cross attention, fs conditioning) with model_channels 64 instead of 320, T=4 frames and a
16x16 latent, so the unmodified reference runs on CPU in seconds.
"""

TINY_T = 4
TINY_LATENT_HW = (16, 16)
TINY_CONTEXT_DIM = 128

TINY_UNET = dict(
    in_channels=8, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64, transformer_depth=1,
    context_dim=TINY_CONTEXT_DIM, use_linear=True, use_checkpoint=True, temporal_conv=True,
    temporal_attention=True, temporal_selfatt_only=True, use_relative_position=False,
    use_causal_attention=False, temporal_length=TINY_T, addition_attention=True, image_cross_attention=True,
    default_fs=24, fs_condition=True)

TINY_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=64,
                     ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)

FULL_UNET = dict(
    in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], dropout=0.1, num_head_channels=64, transformer_depth=1, context_dim=1024,
    use_linear=True, use_checkpoint=True, temporal_conv=True, temporal_attention=True,
    temporal_selfatt_only=True, use_relative_position=False, use_causal_attention=False, temporal_length=16,
    addition_attention=True, image_cross_attention=True, default_fs=24, fs_condition=True)

FULL_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                     ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def model_config(unet=None, ddconfig=None):
    """YAML-equivalent dict of configs/inference_512_v1.0.yaml with conditioning stages stubbed out."""
    return {
        "target": "lvdm.models.ddpm3d.LatentVisualDiffusion",
        "params": dict(
            rescale_betas_zero_snr=True, parameterization="v", linear_start=0.00085, linear_end=0.012,
            num_timesteps_cond=1, timesteps=1000, first_stage_key="video", cond_stage_key="caption",
            cond_stage_trainable=False, conditioning_key="hybrid", image_size=[40, 64], channels=4,
            scale_by_std=False, scale_factor=0.18215, use_ema=False, uncond_type="empty_seq",
            use_dynamic_rescale=True, base_scale=0.7, fps_condition_type="fps", perframe_ae=True, loop_video=True,
            unet_config={"target": "lvdm.modules.networks.openaimodel3d.UNetModel",
                         "params": dict(unet or TINY_UNET)},
            first_stage_config={"target": "lvdm.models.autoencoder.AutoencoderKL_Dualref",
                                "params": dict(embed_dim=4, monitor="val/rec_loss",
                                               ddconfig=dict(ddconfig or TINY_DDCONFIG),
                                               lossconfig={"target": "torch.nn.Identity"})},
            cond_stage_config={"target": "torch.nn.Identity"},
            img_cond_stage_config={"target": "torch.nn.Identity"},
            image_proj_stage_config={"target": "torch.nn.Identity"},
        ),
    }
