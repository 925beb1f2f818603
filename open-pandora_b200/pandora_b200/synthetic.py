"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

The reference constructor zero-initialises every residual-branch output layer
(`openaimodel3d.py:182,272-273,385-386,549`, `attention.py:288-290,360-362`), so a freshly
constructed model outputs exactly 0 and exercises nothing.  Every tensor is therefore re-drawn
here.  Each tensor gets its own generator seeded from (seed, crc32(name)), so any component can
regenerate any parameter by name without agreeing on an iteration order.

Recipe: matrices / conv kernels ~ N(0, 1/fan_in); 1-D ``*.weight`` (norm
gains) = 1 + 0.1 N(0,1); other 1-D (biases) = 0.02 N(0,1); everything rounded through bf16 so the
fp32 and bf16 arms hold bit-identical weights.
"""
import math
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(((zlib.crc32(name.encode()) * 2654435761) ^ (seed * 40503 + 17)) & 0x7FFFFFFFFFFF)
    return g


def synth_param(name: str, shape, seed: int = 1234) -> torch.Tensor:
    shape = tuple(int(s) for s in shape)
    g = _gen(name, seed)
    if len(shape) > 1:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        w = torch.randn(shape, generator=g, dtype=torch.float32) / math.sqrt(fan_in)
    elif name.endswith(".weight"):
        w = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
    else:
        w = 0.02 * torch.randn(shape, generator=g, dtype=torch.float32)
    return w.to(torch.bfloat16).to(torch.float32)


def synth_state_dict(shapes, seed: int = 1234):
    """shapes: mapping name -> shape (e.g. from ``{k: v.shape for k, v in sd.items()}``)."""
    return {k: synth_param(k, s, seed) for k, s in shapes.items()}


@torch.no_grad()
def randomize_module_(module: torch.nn.Module, seed: int = 1234) -> torch.nn.Module:
    """Overwrite every parameter of ``module`` in place with the synthetic recipe."""
    for name, p in module.state_dict().items():
        if p.is_floating_point():
            p.copy_(synth_param(name, p.shape, seed).to(p.dtype))
    return module


def synth_inputs(batch=1, latent=(4, 16, 40, 64), context_dim=1024, seed=0, per_frame_image=True):
    """x, c_concat, cond context, uncond context, all N(0,1)."""
    c, t, h, w = latent
    L = 77 + 16 * t if per_frame_image else 77
    out = {}
    for i, (k, shp) in enumerate([
        ("x", (batch, c, t, h, w)),
        ("c_concat", (batch, c, t, h, w)),
        ("context", (batch, L, context_dim)),
        ("uc_context", (batch, L, context_dim)),
    ]):
        out[k] = torch.randn(shp, generator=_gen("input." + k, seed), dtype=torch.float32)
    return out
