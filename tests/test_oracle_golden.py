"""CPU: the oracle restatements reproduce the outputs recorded from the reference itself.

The golden files were written by `oracle/make_golden.py`, which executed the unmodified reference
(`/root/reference/DynamiCrafter`) in the build container.
"""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import ddim_oracle, unet_oracle
from pandora_b200.synthetic import synth_inputs, synth_state_dict

UNET_CASES = ["unet_small_perframe", "unet_small_t499_nofs", "unet_small_sharedimg", "unet_small_b2ish"]
DDIM_CASES = ["ddim_s5_cfg75", "ddim_s8_cfg75_r07", "ddim_s4_eta0", "ddim_s50_cfg75_r07"]
#: three-branch guidance, produced by the reference's ddim_multiplecond.DDIMSampler (make_golden.py --multicond)
MULTICOND_CASES = ["ddim_multicond_s5_img2", "ddim_multicond_s4_default"]


def image_only_conditioning(inp, n_text=77):
    """"image kept, text dropped" conditioning of the third branch, as make_golden.py built it."""
    return dict(c_crossattn=[torch.cat([inp["uc_context"][:, :n_text], inp["context"][:, n_text:]], 1)],
                c_concat=[inp["c_concat"]])


def unet_case_inputs(meta):
    p = meta["params"]
    inp = synth_inputs(1, tuple(meta["latent"]), p["context_dim"], meta["seed_x"], meta["per_frame_image"])
    if not meta["per_frame_image"]:
        extra = synth_inputs(1, tuple(meta["latent"]), p["context_dim"], meta["seed_x"] + 7)["context"][:, :16]
        inp["context"] = torch.cat([inp["context"], extra], 1)
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([meta["t"]], dtype=torch.long)
    fs = None if meta["fs"] is None else torch.tensor([meta["fs"]], dtype=torch.long)
    return x, t, inp["context"], fs


@pytest.mark.parametrize("name", UNET_CASES)
def test_unet_oracle_matches_reference_golden(name):
    z, meta = load_golden(name)
    p = meta["params"]
    sd = synth_state_dict(unet_oracle.unet_param_shapes(p), meta["seed_w"])
    x, t, ctx, fs = unet_case_inputs(meta)
    assert abs(float(x.double().sum()) - meta["x_sum"]) < 1e-6 * max(1.0, abs(meta["x_sum"]))
    w_sum = float(sum(v.double().sum() for v in sd.values()))
    assert abs(w_sum - meta["w_sum"]) < 1e-6 * max(1.0, abs(meta["w_sum"]))
    y = unet_oracle.unet_forward(sd, p, x, t, ctx, fs)
    assert rel_l2(y, z["y"]) < 2e-5
    assert rel_l2(y, z["y64"]) < 2e-5


def test_unet_oracle_matches_full_width_reference_golden():
    """The REAL configuration (1.44 B parameters, latent 4x16x40x64, context 333x1024): the oracle against the forward
    the reference itself computed in the build container (`oracle/make_golden.py --full`)."""
    from pandora_b200.config import UNET_512
    z, meta = load_golden("unet_full512")
    assert meta["params"] == UNET_512
    sd = synth_state_dict(unet_oracle.unet_param_shapes(UNET_512), meta["seed_w"])
    x, t, ctx, fs = unet_case_inputs(meta)
    y = unet_oracle.unet_forward(sd, UNET_512, x, t, ctx, fs)
    assert rel_l2(y, z["y"]) < 2e-5
    assert rel_l2(y, z["y64"]) < 2e-5


@pytest.mark.parametrize("name", ["vae_full_width_8x12", "vae_narrow_16x8"])
def test_vae_decoder_oracle_matches_reference_golden(name):
    """First-stage decoder: the oracle against `AutoencoderKL.decode` outputs of the reference itself."""
    from oracle import vae_oracle
    z, meta = load_golden(name)
    sd = synth_state_dict(vae_oracle.vae_param_shapes(meta["ddconfig"]), meta["seed_w"])
    lat = synth_inputs(1, tuple(meta["latent"]), 8, meta["seed_x"])["x"]
    assert abs(float(lat.double().sum()) - meta["z_sum"]) < 1e-6 * max(1.0, abs(meta["z_sum"]))
    y = vae_oracle.decode_first_stage(sd, lat, meta["scale_factor"])
    assert rel_l2(y, z["y"]) < 2e-5 and rel_l2(y, z["y64"]) < 2e-5


def test_param_inventory_full_size():
    from pandora_b200.config import UNET_512
    shapes = unet_oracle.unet_param_shapes(UNET_512)
    assert len(shapes) == 1516
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 1438854980
    assert shapes["input_blocks.1.0.temopral_conv.conv1.2.weight"] == (320, 320, 3, 1, 1)
    assert shapes["init_attn.0.proj_in.weight"] == (512, 320, 1)
    assert shapes["input_blocks.1.1.transformer_blocks.0.attn2.to_k_ip.weight"] == (320, 1024)


def test_timestep_embedding_uses_bf16_table():
    t = torch.tensor([999, 19], dtype=torch.long)
    e = unet_oracle.timestep_embedding(t, 320)
    half = 160
    f32 = torch.exp(-np.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    e32 = torch.cat([torch.cos(t[:, None].float() * f32), torch.sin(t[:, None].float() * f32)], -1)
    assert e.shape == (2, 320)
    assert float((e - e32).abs().max()) > 0.5          # the bf16 table really matters


@pytest.mark.parametrize("name", DDIM_CASES + MULTICOND_CASES)
def test_ddim_oracle_matches_reference_golden(name):
    z, meta = load_golden(name)
    p, dp = meta["params"], meta["dparams"]
    buf = ddim_oracle.model_schedule(**dp)
    tab = ddim_oracle.ddim_tables(buf, meta["S"], meta["spacing"], meta["eta"])
    assert np.array_equal(tab["timesteps"], z["timesteps"])
    for k in ("alphas", "alphas_prev", "sigmas"):
        assert np.array_equal(np.array(tab[k]), z[k]), k
    assert np.array_equal(buf["sqrt_alphas_cumprod"].float().numpy(), z["sqrt_ac"])
    sd = synth_state_dict(unet_oracle.unet_param_shapes(p), meta["seed_w"])
    inp = synth_inputs(1, tuple(meta["latent"]), p["context_dim"], meta["seed_x"])
    cond = dict(c_crossattn=[inp["context"]], c_concat=[inp["c_concat"]])
    uc = dict(c_crossattn=[inp["uc_context"]], c_concat=[inp["c_concat"]])
    fs = torch.tensor([p["default_fs"]], dtype=torch.long)
    fn = lambda xc, t, cc, fs_: unet_oracle.unet_forward(sd, p, xc, t, cc, fs_)
    torch.manual_seed(meta["noise_seed"])
    uc_img = image_only_conditioning(inp) if meta.get("multicond") else None
    x, x0 = ddim_oracle.ddim_sample(fn, buf, inp["x"], cond, uc, meta["S"], meta["spacing"], meta["eta"],
                                    meta["cfg_scale"], meta["guidance_rescale"], fs, uncond_img=uc_img,
                                    cfg_img=meta.get("cfg_img"))
    assert rel_l2(x, z["samples"]) < 1e-4
    assert rel_l2(x0, z["pred_x0"]) < 1e-4


def test_trailing_schedule_quirks():
    buf = ddim_oracle.model_schedule()
    ac = buf["alphas_cumprod"].float()
    assert ac[0] == 1.0 and ac[1] == 1.0 and ac[999] == 0.0      # bf16 storage
    ts = ddim_oracle.ddim_timesteps("uniform_trailing", 50)
    assert ts[0] == 19 and ts[-1] == 999 and len(ts) == 50
    assert buf["scale_arr"].shape[0] == 1400
