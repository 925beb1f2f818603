"""Minimal latent-diffusion wrapper around the denoiser: the part of `LatentDiffusion` the sampler uses.

Mirrors (reference: /root/reference/DynamiCrafter/lvdm/models/ddpm3d.py)
  DDPM.register_schedule                 :119-182   all buffers stored in bf16 (:139)
  LatentDiffusion dynamic-rescale table  :505-510   1400 entries = linspace(1, base, 400) + full(1000, base)
  predict_start/eps_from_z_and_v         :235-247
  q_sample                               :301-305
  LatentDiffusion.apply_model            :724-739
  DiffusionWrapper.forward               :1066-1129 ('hybrid', 'concat', 'crossattn' keys)
The VAE, text/image encoders, EMA, losses and Lightning plumbing of the reference class are outside
the hot path and are not reproduced; `DDIMSampler` only reads what is here, so
either this class or the reference's own `LatentVisualDiffusion` can be handed to the sampler.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from .unet import UNetModel


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule {schedule!r} (the Open-Pandora configs use 'linear')")
    return (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()


def rescale_zero_terminal_snr(betas):
    """utils_diffusion.py:112-144 (arXiv 2305.08891, Algorithm 1)."""
    abar_sqrt = np.sqrt(np.cumprod(1.0 - betas, axis=0))
    first, last = abar_sqrt[0].copy(), abar_sqrt[-1].copy()
    abar_sqrt -= last
    abar_sqrt *= first / (first - last)
    abar = abar_sqrt ** 2
    return 1 - np.concatenate([abar[0:1], abar[1:] / abar[:-1]])


def extract_into_tensor(a, t, x_shape):
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class DiffusionWrapper(nn.Module):
    def __init__(self, diffusion_model, conditioning_key):
        super().__init__()
        self.diffusion_model = diffusion_model
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None, **kwargs):
        key = self.conditioning_key
        if key == "hybrid":
            cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
            if isinstance(self.diffusion_model, UNetModel):
                # the channel concat is folded into the layout kernel (no torch.cat on the data path)
                return self.diffusion_model(x, t, context=cc, x_concat=list(c_concat), **kwargs)
            return self.diffusion_model(torch.cat([x] + list(c_concat), dim=1), t, context=cc, **kwargs)
        if key == "concat":
            return self.diffusion_model(torch.cat([x] + list(c_concat), dim=1), t, **kwargs)
        if key == "crossattn":
            return self.diffusion_model(x, t, context=torch.cat(c_crossattn, 1), **kwargs)
        raise NotImplementedError(f"conditioning_key {key!r}")


class LatentDiffusionLite(nn.Module):
    def __init__(self, unet, timesteps=1000, linear_start=0.00085, linear_end=0.012, parameterization="v",
                 rescale_betas_zero_snr=True, use_dynamic_rescale=True, base_scale=0.7, turning_step=400,
                 conditioning_key="hybrid", beta_schedule="linear"):
        super().__init__()
        assert parameterization in ("eps", "x0", "v")
        self.parameterization = parameterization
        self.model = DiffusionWrapper(unet, conditioning_key)
        self.use_dynamic_rescale = use_dynamic_rescale
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start, linear_end)
        if rescale_betas_zero_snr:
            betas = rescale_zero_terminal_snr(betas)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        self.num_timesteps = int(betas.shape[0])
        bf = partial(torch.tensor, dtype=torch.bfloat16)   # the reference keeps these in bf16
        self.register_buffer("betas", bf(betas))
        self.register_buffer("alphas_cumprod", bf(ac))
        self.register_buffer("alphas_cumprod_prev", bf(np.append(1.0, ac[:-1])))
        self.register_buffer("sqrt_alphas_cumprod", bf(np.sqrt(ac)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", bf(np.sqrt(1.0 - ac)))
        with np.errstate(divide="ignore"):
            self.register_buffer("log_one_minus_alphas_cumprod", bf(np.log(1.0 - ac)))
        if use_dynamic_rescale:
            self.register_buffer("scale_arr", bf(np.concatenate(
                (np.linspace(1.0, base_scale, turning_step), np.full(self.num_timesteps, base_scale)))))

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond, **kwargs):
        if not isinstance(cond, dict):
            cond = cond if isinstance(cond, list) else [cond]
            cond = {"c_concat" if self.model.conditioning_key == "concat" else "c_crossattn": cond}
        out = self.model(x_noisy, t, **cond, **kwargs)
        return out[0] if isinstance(out, tuple) else out

    def predict_start_from_z_and_v(self, x_t, t, v):
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t
                - extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def predict_eps_from_z_and_v(self, x_t, t, v):
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_t.shape) * v
                + extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * x_t)

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)
