"""Hyper-parameters of the denoiser hot path.

The reference selects these through OmegaConf YAML
(`DynamiCrafter/configs/inference_512_v1.0.yaml:24-55`, `inference_1024_v1.0.yaml`);
OmegaConf is not a dependency of this package, so the same values are kept here as plain dicts
that can be splatted into `UNetModel(**params)` exactly as `instantiate_from_config` would
(`DynamiCrafter/utils/utils.py:27-34`).
"""

UNET_512 = dict(
    in_channels=8,
    out_channels=4,
    model_channels=320,
    attention_resolutions=[4, 2, 1],
    num_res_blocks=2,
    channel_mult=[1, 2, 4, 4],
    dropout=0.1,
    num_head_channels=64,
    transformer_depth=1,
    context_dim=1024,
    use_linear=True,
    use_checkpoint=True,
    temporal_conv=True,
    temporal_attention=True,
    temporal_selfatt_only=True,
    use_relative_position=False,
    use_causal_attention=False,
    temporal_length=16,
    addition_attention=True,
    image_cross_attention=True,
    default_fs=24,
    fs_condition=True,
)

UNET_1024 = dict(UNET_512, default_fs=10)

# Diffusion-wrapper parameters the sampler reads (`inference_512_v1.0.yaml:3-23`).
DIFFUSION_512 = dict(
    timesteps=1000,
    linear_start=0.00085,
    linear_end=0.012,
    parameterization="v",
    rescale_betas_zero_snr=True,
    use_dynamic_rescale=True,
    base_scale=0.7,
    turning_step=400,
    conditioning_key="hybrid",
)
DIFFUSION_1024 = dict(DIFFUSION_512, base_scale=0.3)

# First-stage (KL-VAE) configuration (`inference_512_v1.0.yaml:57-79`, identical for 1024) and the latent scale
# factor of LatentDiffusion (`inference_512_v1.0.yaml:14`).
VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)
SCALE_FACTOR = 0.18215

# Latent shapes (C, T, H, W) of the two headline resolutions.
LATENT_512 = (4, 16, 40, 64)
LATENT_1024 = (4, 16, 72, 128)

TEXT_CONTEXT_LEN = 77      # `attention.py:45` text_context_len
IMAGE_TOKENS_PER_FRAME = 16  # `openaimodel3d.py:560` hard-coded 77 + t*16


def small_unet(model_channels=64, context_dim=128, **over):
    """A structurally identical U-Net that is cheap enough for CPU oracles and unit tests."""
    p = dict(UNET_512, model_channels=model_channels, context_dim=context_dim)
    p.update(over)
    return p
