"""pandora_b200 — B200-native (sm_100a) implementation of Open-Pandora's video-denoising hot path:
DynamiCrafter's `UNetModel.forward` inside `DDIMSampler`'s classifier-free-guidance loop.

Public surface mirrors the reference's (`UNetModel`, `DDIMSampler`), see INTEGRATION.md:

    import pandora_b200
    pandora_b200.install()      # lvdm.modules.networks.openaimodel3d.UNetModel / ...ddim.DDIMSampler now
                                # resolve to the native classes; config.yaml / model.py stay unchanged
"""
import importlib
import sys

from . import abi, config  # noqa: F401
from .diffusion import DiffusionWrapper, LatentDiffusionLite  # noqa: F401
from .sampler import DDIMSampler, DDIMSampler_multicond  # noqa: F401
from .unet import UNetModel  # noqa: F401
from .vae import AutoencoderKL, swap_vae_decoder  # noqa: F401

__all__ = ["UNetModel", "DDIMSampler", "DDIMSampler_multicond", "LatentDiffusionLite", "DiffusionWrapper", "AutoencoderKL", "install",
           "swap_unet", "swap_vae_decoder", "abi", "config"]

_UNET_MODULES = ("lvdm.modules.networks.openaimodel3d", "DynamiCrafter.lvdm.modules.networks.openaimodel3d")
_DDIM_MODULES = ("lvdm.models.samplers.ddim", "DynamiCrafter.lvdm.models.samplers.ddim")
_DDIM3_MODULES = ("lvdm.models.samplers.ddim_multiplecond", "DynamiCrafter.lvdm.models.samplers.ddim_multiplecond")


def install(strict=False):
    """Make the reference's dotted paths resolve to the native classes.

    `instantiate_from_config` does `getattr(importlib.import_module(module), cls)`
    (DynamiCrafter/utils/utils.py:37-42), so replacing the attribute on the already-importable
    reference modules is enough for `target: lvdm.modules.networks.openaimodel3d.UNetModel`
    (configs/inference_512_v1.0.yaml:25).  `model.py` binds `DDIMSampler` by name at import
    (`model.py:448`) and the three-branch class as `DDIMSampler_multicond` (`model.py:449`), so an
    already-imported `model` module is patched as well.
    Returns the list of patched module names."""
    patched = []
    for names, attr, obj in ((_UNET_MODULES, "UNetModel", UNetModel), (_DDIM_MODULES, "DDIMSampler", DDIMSampler),
                             (_DDIM3_MODULES, "DDIMSampler", DDIMSampler_multicond)):
        for name in names:
            try:
                mod = importlib.import_module(name)
            except Exception:
                if strict:
                    raise
                continue
            setattr(mod, attr, obj)
            patched.append(name)
    for name, mod in list(sys.modules.items()):
        if name.split(".")[-1] == "model" and getattr(mod, "DDIMSampler", None) is not None \
                and mod.DDIMSampler is not DDIMSampler and hasattr(mod, "WorldModel"):
            mod.DDIMSampler = DDIMSampler
            if hasattr(mod, "DDIMSampler_multicond"):
                mod.DDIMSampler_multicond = DDIMSampler_multicond
            patched.append(name)
    return patched


def swap_unet(latent_diffusion, **ctor_kwargs):
    """Replace `latent_diffusion.model.diffusion_model` (a reference UNetModel) by a native one that
    shares its Parameter objects.  `ctor_kwargs` = the `unet_config.params` the reference was built with."""
    ref = latent_diffusion.model.diffusion_model
    if isinstance(ref, UNetModel):
        return ref
    new = UNetModel.from_reference(ref, **ctor_kwargs)
    latent_diffusion.model.diffusion_model = new
    return new
