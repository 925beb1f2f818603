"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF (build container only).

    python oracle/make_golden.py            # writes tests/golden/, prints oracle-vs-reference errors
    python oracle/make_golden.py --full     # the full-width (1.44 B parameter) fixture, ~15 min of CPU

Imports the unmodified reference modules from /root/reference/DynamiCrafter (they need only
torch/einops/numpy), builds structurally identical but narrow models (model_channels 64),
overwrites every parameter with the synthetic recipe (`pandora_b200.synthetic`), runs
`UNetModel.forward` and `DDIMSampler.sample`, and stores inputs-by-seed + outputs.  The same run
checks the oracle restatements (`oracle/unet_oracle.py`, `oracle/ddim_oracle.py`) against the
reference outputs, which is what pins the oracle.  /root/reference does not exist on the GPU box,
so nothing at test/bench time imports this script.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PANDORA_REFERENCE", "/root/reference/DynamiCrafter")
sys.path.insert(0, os.path.join(ROOT, "open-pandora_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from pandora_b200 import config as pcfg  # noqa: E402
from pandora_b200.synthetic import randomize_module_, synth_inputs, synth_state_dict  # noqa: E402
from oracle import ddim_oracle, unet_oracle, vae_oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


class AttrDict(dict):
    """Stand-in for OmegaConf nodes: item access + attribute access (ddpm3d.py:83)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return AttrDict(v) if isinstance(v, dict) else v


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def build_reference_unet(params, seed):
    from lvdm.modules.networks.openaimodel3d import UNetModel
    m = UNetModel(**dict(params, use_checkpoint=False)).eval()
    randomize_module_(m, seed)
    return m


def unet_case(name, params, latent, per_frame_image, seed_w, seed_x, tval, fs):
    ref = build_reference_unet(params, seed_w)
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    # the oracle's own parameter inventory must reproduce the reference's state dict exactly
    shapes = unet_oracle.unet_param_shapes(params)
    assert set(shapes) == set(sd), (set(shapes) ^ set(sd))
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    sd2 = synth_state_dict(shapes, seed_w)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)

    inp = synth_inputs(1, latent, params["context_dim"], seed_x, per_frame_image)
    if not per_frame_image:  # one 16-token image for the whole clip -> `else` branch :565-566
        inp["context"] = torch.cat([inp["context"], synth_inputs(1, latent, params["context_dim"], seed_x + 7)["context"][:, :16]], 1)
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([tval], dtype=torch.long)
    fs_t = None if fs is None else torch.tensor([fs], dtype=torch.long)
    with torch.no_grad():
        y_ref = ref(x, t, context=inp["context"], fs=fs_t, cfg_img=None, round_info=(1, 1))
        y_ref64 = ref.double()(x.double(), t, context=inp["context"].double(), fs=fs_t).float()
    y_orc = unet_oracle.unet_forward(sd2, params, x, t, inp["context"], fs_t)
    e = rel_l2(y_orc, y_ref)
    print(f"[unet:{name}] oracle vs reference fp32 rel-L2 = {e:.3e}; "
          f"reference fp32 vs fp64 = {rel_l2(y_ref, y_ref64):.3e}; out std = {float(y_ref.std()):.3f}")
    assert e < 2e-5, e
    np.savez_compressed(
        os.path.join(GOLD, f"unet_{name}.npz"),
        y=y_ref.numpy(), y64=y_ref64.numpy(),
        meta=json.dumps(dict(params=params, latent=latent, per_frame_image=per_frame_image,
                             seed_w=seed_w, seed_x=seed_x, t=tval, fs=fs,
                             x_sum=float(x.double().sum()), w_sum=float(sum(v.double().sum() for v in sd.values())))))
    return ref


def sampler_case(name, params, dparams, latent, S, cfg_scale, rescale, seed_w, seed_x, spacing, eta,
                 multicond=False, cfg_img=None):
    """`multicond`: the three-branch class (ddim_multiplecond.py).  Its make_schedule calls numpy on the model's
    bf16 schedule buffers (:40, TypeError), so the reference is run the only way it can run: on `model.float()`
    (same buffer VALUES, fp32 containers); the DDIM tables are then identical to ddim.py's."""
    from lvdm.models.ddpm3d import LatentVisualDiffusion
    if multicond:
        from lvdm.models.samplers.ddim_multiplecond import DDIMSampler
    else:
        from lvdm.models.samplers.ddim import DDIMSampler
    ident = {"target": "torch.nn.Identity"}
    unet_config = AttrDict(target="lvdm.modules.networks.openaimodel3d.UNetModel",
                           params=dict(params, use_checkpoint=False))
    model = LatentVisualDiffusion(
        img_cond_stage_config=ident, image_proj_stage_config=ident,
        first_stage_config=AttrDict(ident), cond_stage_config=AttrDict(ident),
        unet_config=unet_config, timesteps=dparams["timesteps"],
        linear_start=dparams["linear_start"], linear_end=dparams["linear_end"],
        parameterization=dparams["parameterization"],
        rescale_betas_zero_snr=dparams["rescale_betas_zero_snr"],
        use_dynamic_rescale=dparams["use_dynamic_rescale"], base_scale=dparams["base_scale"],
        conditioning_key=dparams["conditioning_key"], use_ema=False, channels=4,
        image_size=list(latent[2:]), first_stage_key="video", cond_stage_key="caption",
        scale_factor=0.18215, uncond_type="empty_seq", num_timesteps_cond=1).eval()
    randomize_module_(model.model.diffusion_model, seed_w)
    DDIMSampler.register_buffer = lambda self, n, a: setattr(self, n, a)  # CPU shim (ddim.py:18-22)
    sampler = DDIMSampler(model)
    inp = synth_inputs(1, latent, params["context_dim"], seed_x)
    cond = dict(c_crossattn=[inp["context"]], c_concat=[inp["c_concat"]])
    uc = dict(c_crossattn=[inp["uc_context"]], c_concat=[inp["c_concat"]])
    uc_img = None
    if multicond:
        # "image kept, text dropped" (inference.py:196-203): the empty-prompt text tokens followed by the image tokens
        n_txt = 77
        uc_img = dict(c_crossattn=[torch.cat([inp["uc_context"][:, :n_txt], inp["context"][:, n_txt:]], 1)],
                      c_concat=[inp["c_concat"]])
        try:
            DDIMSampler(model).make_schedule(S, spacing, eta, verbose=False)
            raise AssertionError("the reference's multicond make_schedule was expected to fail on bf16 buffers")
        except TypeError as e:
            print(f"[ddim:{name}] reference make_schedule on the bf16 schedule: TypeError: {e}")
        model_buffers_bf16 = {k: getattr(model, k).clone() for k in
                              ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "scale_arr")}
        model.float()
        sampler = DDIMSampler(model)
    fs = torch.tensor([params["default_fs"]], dtype=torch.long)
    torch.manual_seed(seed_x + 99)
    samples, inter = sampler.sample(
        S=S, batch_size=1, shape=latent, conditioning=cond, unconditional_conditioning=uc,
        unconditional_guidance_scale=cfg_scale, eta=eta, x_T=inp["x"], fs=fs, verbose=False,
        timestep_spacing=spacing, guidance_rescale=rescale, cfg_img=cfg_img, round_info=(1, 1),
        gr_progress_bar=None, unconditional_conditioning_img_nonetext=uc_img)
    # ---- oracle arm: same seeds, same draw order
    buf = ddim_oracle.model_schedule(**dparams)
    for k in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "scale_arr"):
        if multicond:
            assert torch.equal(buf[k], model_buffers_bf16[k]) and torch.equal(buf[k].float(), getattr(model, k)), k
        else:
            assert torch.equal(buf[k], getattr(model, k)), k
    tab = ddim_oracle.ddim_tables(buf, S, spacing, eta)
    assert np.array_equal(tab["timesteps"], sampler.ddim_timesteps)
    for k, ref_v in (("alphas", sampler.ddim_alphas), ("alphas_prev", sampler.ddim_alphas_prev),
                     ("sigmas", sampler.ddim_sigmas), ("sqrt_one_minus_alphas", sampler.ddim_sqrt_one_minus_alphas)):
        assert [float(v) for v in ref_v] == tab[k], k
    sd = synth_state_dict(unet_oracle.unet_param_shapes(params), seed_w)
    fn = lambda xc, t, cc, fs_: unet_oracle.unet_forward(sd, params, xc, t, cc, fs_)
    torch.manual_seed(seed_x + 99)
    x_orc, _ = ddim_oracle.ddim_sample(fn, buf, inp["x"], cond, uc, S, spacing, eta, cfg_scale, rescale, fs,
                                       uncond_img=uc_img, cfg_img=cfg_img)
    e = rel_l2(x_orc, samples)
    print(f"[ddim:{name}] oracle vs reference final latent rel-L2 = {e:.3e}")
    assert e < 1e-4, e
    np.savez_compressed(
        os.path.join(GOLD, f"ddim_{name}.npz"), samples=samples.numpy(),
        pred_x0=inter["pred_x0"][-1].numpy(),
        alphas=np.array(tab["alphas"]), alphas_prev=np.array(tab["alphas_prev"]),
        sigmas=np.array(tab["sigmas"]), timesteps=tab["timesteps"],
        scale=np.array(tab.get("scale", [])), scale_prev=np.array(tab.get("scale_prev", [])),
        sqrt_ac=model.sqrt_alphas_cumprod.float().numpy(),
        sqrt_omac=model.sqrt_one_minus_alphas_cumprod.float().numpy(),
        meta=json.dumps(dict(params=params, dparams=dparams, latent=latent, S=S, cfg_scale=cfg_scale,
                             guidance_rescale=rescale, seed_w=seed_w, seed_x=seed_x, spacing=spacing,
                             eta=eta, noise_seed=seed_x + 99, multicond=multicond, cfg_img=cfg_img)))


def full_width_case(name="full512", seed_w=1234, seed_x=0, stride=8):
    """The REAL configuration (inference_512_v1.0.yaml U-Net, 1.44 B parameters, latent 4x16x40x64, context 333x1024)
    executed by the reference itself: one forward stored whole (fp32 and fp64), plus teacher-forced forwards at five
    DDIM timesteps of the 50-step trailing schedule with the cond and the uncond context — x_t = q_sample(x0, t) with
    the model's bf16 schedule buffers, so every timestep sees a latent of the right statistics — stored strided
    (every `stride`-th element) to keep the fixture small.  Pins the oracle AND the GPU path at full width."""
    params, latent = pcfg.UNET_512, pcfg.LATENT_512
    ref = build_reference_unet(params, seed_w)
    shapes = unet_oracle.unet_param_shapes(params)
    sd2 = synth_state_dict(shapes, seed_w)
    sd = ref.state_dict()
    assert set(shapes) == set(sd) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    inp = synth_inputs(1, latent, params["context_dim"], seed_x)
    buf = ddim_oracle.model_schedule(**pcfg.DIFFUSION_512)
    fs_t = torch.tensor([params["default_fs"]], dtype=torch.long)
    out = {}
    with torch.no_grad():
        x = torch.cat([inp["x"], inp["c_concat"]], 1)
        t = torch.tensor([499], dtype=torch.long)
        y_ref = ref(x, t, context=inp["context"], fs=fs_t)
        y_orc = unet_oracle.unet_forward(sd2, params, x, t, inp["context"], fs_t)
        e = rel_l2(y_orc, y_ref)
        print(f"[unet:{name}] t=499 oracle vs reference fp32 rel-L2 = {e:.3e}; out std = {float(y_ref.std()):.3f}", flush=True)
        assert e < 2e-5, e
        out["y"] = y_ref.numpy()
        steps = [999, 759, 499, 259, 19]
        x0 = inp["x"]
        noise = synth_inputs(1, latent, params["context_dim"], seed_x + 11)["x"]
        for tv in steps:
            tt = torch.tensor([tv], dtype=torch.long)
            xt = (buf["sqrt_alphas_cumprod"].float()[tv] * x0 + buf["sqrt_one_minus_alphas_cumprod"].float()[tv] * noise)
            xc = torch.cat([xt, inp["c_concat"]], 1)
            for tag, ctx in (("c", inp["context"]), ("u", inp["uc_context"])):
                v = ref(xc, tt, context=ctx, fs=fs_t)
                out[f"v_{tag}_{tv}"] = v.flatten()[::stride].numpy().copy()
                if tv in (999, 19) and tag == "c":
                    eo = rel_l2(unet_oracle.unet_forward(sd2, params, xc, tt, ctx, fs_t), v)
                    print(f"[unet:{name}] t={tv} oracle vs reference rel-L2 = {eo:.3e}", flush=True)
                    assert eo < 2e-5, eo
            print(f"[unet:{name}] teacher-forced t={tv} done", flush=True)
        y64 = ref.double()(x.double(), t, context=inp["context"].double(), fs=fs_t).float()
        print(f"[unet:{name}] reference fp32 vs fp64 = {rel_l2(y_ref, y64):.3e}", flush=True)
        out["y64"] = y64.numpy()
    np.savez_compressed(
        os.path.join(GOLD, f"unet_{name}.npz"), **out,
        meta=json.dumps(dict(params=params, latent=list(latent), seed_w=seed_w, seed_x=seed_x, t=499,
                             fs=params["default_fs"], steps=steps, stride=stride, noise_seed=seed_x + 11,
                             dparams=pcfg.DIFFUSION_512, per_frame_image=True)))


def vae_case(name, ddconfig, latent, seed_w, seed_x, scale_factor=0.18215):
    """The reference's own AutoencoderKL.decode (lvdm/models/autoencoder.py:104-107) on synthetic weights: per-frame
    decode of a small video latent, as LatentDiffusion.decode_core does (ddpm3d.py:630-651)."""
    from lvdm.models.autoencoder import AutoencoderKL
    m = AutoencoderKL(ddconfig=ddconfig, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4).eval()
    randomize_module_(m, seed_w)
    sd = {k: v.detach() for k, v in m.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    shapes = vae_oracle.vae_param_shapes(ddconfig)
    assert set(shapes) == set(sd), set(shapes) ^ set(sd)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in sd)
    sd2 = synth_state_dict(shapes, seed_w)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    z = synth_inputs(1, latent, 8, seed_x)["x"]                       # (1, 4, t, h, w) ~ N(0, 1)
    zz = z.permute(0, 2, 1, 3, 4).reshape(-1, *z.shape[1:2], *z.shape[3:])
    with torch.no_grad():
        ref = torch.cat([m.decode((1.0 / scale_factor) * zz[i:i + 1]) for i in range(zz.shape[0])], 0)
        ref64 = torch.cat([m.double().decode((1.0 / scale_factor) * zz[i:i + 1].double()) for i in range(zz.shape[0])], 0).float()
    orc = vae_oracle.decode_first_stage(sd2, z, scale_factor)
    ref5 = ref.reshape(1, zz.shape[0], *ref.shape[1:]).permute(0, 2, 1, 3, 4)
    e = rel_l2(orc, ref5)
    print(f"[vae:{name}] oracle vs reference fp32 rel-L2 = {e:.3e}; reference fp32 vs fp64 = {rel_l2(ref, ref64):.3e}; "
          f"out std = {float(ref.std()):.3f}", flush=True)
    assert e < 2e-5, e
    np.savez_compressed(
        os.path.join(GOLD, f"vae_{name}.npz"), y=ref5.numpy(),
        y64=ref64.reshape(1, zz.shape[0], *ref.shape[1:]).permute(0, 2, 1, 3, 4).numpy(),
        meta=json.dumps(dict(ddconfig=ddconfig, latent=list(latent), seed_w=seed_w, seed_x=seed_x,
                             scale_factor=scale_factor, z_sum=float(z.double().sum()))))


VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)   # inference_512_v1.0.yaml:62-76


def main():
    if "--vae" in sys.argv:          # first-stage decoder fixtures
        os.makedirs(GOLD, exist_ok=True)
        torch.set_num_threads(os.cpu_count())
        vae_case("full_width_8x12", VAE_DDCONFIG, (4, 2, 8, 12), 1234, 0)        # the real decoder, 2 frames, 64 x 96 px
        vae_case("narrow_16x8", dict(VAE_DDCONFIG, ch=64), (4, 3, 16, 8), 77, 3)
        return
    if "--multicond" in sys.argv:    # three-branch guidance fixtures
        os.makedirs(GOLD, exist_ok=True)
        torch.set_num_threads(os.cpu_count())
        small = pcfg.small_unet()
        sampler_case("multicond_s5_img2", small, pcfg.DIFFUSION_512, (4, 16, 8, 8), 5, 7.5, 0.7, 1234, 6,
                     "uniform_trailing", 1.0, multicond=True, cfg_img=2.0)
        sampler_case("multicond_s4_default", small, pcfg.DIFFUSION_1024, (4, 16, 8, 8), 4, 5.0, 0.0, 99, 7,
                     "uniform_trailing", 0.0, multicond=True, cfg_img=None)
        return
    if "--full" in sys.argv:         # ~15 min of CPU: the full-width fixture only
        os.makedirs(GOLD, exist_ok=True)
        torch.set_num_threads(os.cpu_count())
        return full_width_case()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    small = pcfg.small_unet()
    unet_case("small_perframe", small, (4, 16, 8, 8), True, 1234, 0, 999, 24)
    unet_case("small_t499_nofs", small, (4, 16, 16, 8), True, 4321, 3, 499, None)
    unet_case("small_sharedimg", small, (4, 16, 8, 8), False, 1234, 5, 19, 10)
    unet_case("small_b2ish", pcfg.small_unet(context_dim=64), (4, 16, 8, 16), True, 77, 8, 259, 15)
    sampler_case("s5_cfg75", small, pcfg.DIFFUSION_512, (4, 16, 8, 8), 5, 7.5, 0.0, 1234, 0, "uniform_trailing", 1.0)
    sampler_case("s8_cfg75_r07", small, pcfg.DIFFUSION_1024, (4, 16, 8, 8), 8, 7.5, 0.7, 1234, 1, "uniform_trailing", 1.0)
    sampler_case("s50_cfg75_r07", small, pcfg.DIFFUSION_512, (4, 16, 8, 8), 50, 7.5, 0.7, 1234, 4, "uniform_trailing", 1.0)
    sampler_case("s4_eta0", small, pcfg.DIFFUSION_512, (4, 16, 8, 8), 4, 4.0, 0.0, 99, 2, "uniform_trailing", 0.0)


if __name__ == "__main__":
    main()
