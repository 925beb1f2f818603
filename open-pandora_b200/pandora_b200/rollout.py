"""Autoregressive multi-round rollout around the sampler (BASELINE.json config 4), without the VAE.

Open-Pandora chains rounds through pixel space: the last 4 decoded frames of round r are VAE-encoded again and
tiled to 16 frames as the `c_concat` conditioning of round r+1 (`model.py:690-701` `get_latent_z`, `:1179-1187`;
einops ``(repeat t)`` = tile, order f0 f1 f2 f3 f0 ...); round 1 conditions on ONE image latent repeated 16 times.
The VAE and the text / image encoders are outside the hot path, so this helper chains the
rounds in latent space with caller-supplied (synthetic) context tensors: what it exercises is exactly the part of a
rollout that the hot path sees — 2 x steps U-Net forwards per round, fresh conditioning tensors every round (the
context K/V caches and the captured CUDA graph are rebuilt per `sample()` call), fp32 latent carried between steps.
"""
import torch


def chain_c_concat(latent, frames_kept=4):
    """c_concat of the next round from this round's latent (b, c, t, h, w): the last `frames_kept` frames tiled
    along t (f0 f1 f2 f3 f0 f1 ...), as `repeat(z, 'b c t h w -> b c (repeat t) h w')` does in model.py:698."""
    t = latent.shape[2]
    tail = latent[:, :, -frames_kept:]
    reps = (t + frames_kept - 1) // frames_kept
    return tail.repeat(1, 1, reps, 1, 1)[:, :, :t].contiguous()


def first_round_c_concat(image_latent, frames):
    """Round 1: one image latent (b, c, 1, h, w) or (b, c, h, w) repeated over all frames."""
    z = image_latent if image_latent.dim() == 5 else image_latent[:, :, None]
    return z[:, :, :1].repeat(1, 1, frames, 1, 1).contiguous()


@torch.no_grad()
def rollout(sampler, image_latent, contexts, uncond_contexts, shape, steps=50, cfg_scale=4.0, eta=1.0, fs=None,
            guidance_rescale=0.0, timestep_spacing="uniform_trailing", x_T=None, on_round=None):
    """Run len(contexts) rounds; returns the list of per-round latents (b, c, t, h, w).

    contexts / uncond_contexts: per-round cross-attention context tensors (b, L, D) — in Open-Pandora the LLM /
    Q-Former / CLIP output of that round's prompt; x_T: optional list of per-round initial noise (else torch.randn
    is drawn per round exactly as `sample` does); defaults follow ChatWM (`model.py:990-993`: cfg 4, eta 1)."""
    c, t, h, w = shape
    b = image_latent.shape[0]
    z = first_round_c_concat(image_latent, t)
    outs = []
    for r, (ctx, uctx) in enumerate(zip(contexts, uncond_contexts)):
        cond = dict(c_crossattn=[ctx], c_concat=[z])
        uc = dict(c_crossattn=[uctx], c_concat=[z])
        x, _ = sampler.sample(S=steps, batch_size=b, shape=shape, conditioning=cond, unconditional_conditioning=uc,
                              unconditional_guidance_scale=cfg_scale, eta=eta, fs=fs, verbose=False,
                              timestep_spacing=timestep_spacing, guidance_rescale=guidance_rescale,
                              x_T=None if x_T is None else x_T[r], precision=torch.float32)
        outs.append(x)
        if on_round is not None:
            on_round(r, x)
        z = chain_c_concat(x.float())
    return outs
