"""GPU: the native first-stage decoder against outputs of the reference's own
`AutoencoderKL.decode` (tests/golden/vae_*.npz) and, at the real frame size, against the CPU oracle."""
import pytest
import torch

from helpers import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _native(ddconfig, seed, dtype=torch.float32):
    from pandora_b200 import AutoencoderKL
    from pandora_b200.synthetic import randomize_module_
    m = AutoencoderKL(ddconfig=ddconfig, lossconfig={"target": "torch.nn.Identity"}, embed_dim=4).eval()
    randomize_module_(m, seed)
    return m.to(device="cuda", dtype=dtype)


@pytest.mark.parametrize("name", ["vae_full_width_8x12", "vae_narrow_16x8"])
def test_vae_decode_matches_reference_golden(name):
    from pandora_b200.synthetic import synth_inputs
    z, meta = load_golden(name)
    m = _native(meta["ddconfig"], meta["seed_w"])
    lat = synth_inputs(1, tuple(meta["latent"]), 8, meta["seed_x"])["x"].cuda()
    for mode, tol in (("fp32", 1e-3), ("bf16", 2e-2)):
        m.compute_mode = mode
        m.invalidate_packed_weights()
        y = m.decode_first_stage(lat, meta["scale_factor"]).float().cpu()
        e = rel_l2(y, z["y"])
        print(f"{name}: {mode} mode rel-L2 vs the reference = {e:.3e}")
        assert y.shape == z["y"].shape and e < tol
    # frame batching does not change the result (frames are independent): exact mode, fp32 round-off only
    m.compute_mode = "fp32"
    m.invalidate_packed_weights()
    ya = m.decode_first_stage(lat, meta["scale_factor"]).float().cpu()
    m.frames_per_pass = 1
    y1 = m.decode_first_stage(lat, meta["scale_factor"]).float().cpu()
    assert rel_l2(y1, ya) < 1e-5
    # bf16 weights / inputs select the bf16 mode by themselves and return the input dtype
    mb = _native(meta["ddconfig"], meta["seed_w"], torch.bfloat16)
    yb = mb.decode_first_stage(lat.to(torch.bfloat16), meta["scale_factor"])
    assert yb.dtype == torch.bfloat16 and rel_l2(yb.float().cpu(), z["y"]) < 2.5e-2


def test_vae_decode_real_frame_size_matches_oracle():
    """Two frames of the 16x320x512 clip's latents (4 x 40 x 64 -> 3 x 320 x 512) through the full-width decoder: the
    tcgen05 convolutions at 128 ... 512 channels, the 2560-token single-head attention, three x2 upsamples."""
    from oracle import vae_oracle
    from pandora_b200 import config
    from pandora_b200.synthetic import synth_inputs, synth_state_dict
    m = _native(config.VAE_DDCONFIG, 1234)
    lat = synth_inputs(1, (4, 2, 40, 64), 8, 3)["x"]
    sd = synth_state_dict(vae_oracle.vae_param_shapes(config.VAE_DDCONFIG), 1234)
    ref = vae_oracle.decode_first_stage(sd, lat, config.SCALE_FACTOR)
    for mode, tol in (("fp32", 1e-3), ("bf16", 2e-2)):
        m.compute_mode = mode
        m.invalidate_packed_weights()
        y = m.decode_first_stage(lat.cuda(), config.SCALE_FACTOR).float().cpu()
        y2 = m.decode_first_stage(lat.cuda(), config.SCALE_FACTOR).float().cpu()
        e = rel_l2(y, ref)
        print(f"VAE decode 2 x (4x40x64 -> 3x320x512), {mode} mode: rel-L2 vs oracle = {e:.3e}")
        assert torch.equal(y, y2) and torch.isfinite(y).all() and e < tol


def test_softmax_rows_kernel():
    from pandora_b200 import abi
    abi.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    for rows, n in ((7, 96), (33, 2560), (5, 9216), (3, 1001 * 4)):
        x = torch.randn(rows, n, device="cuda", generator=g) * 30
        ref = torch.softmax(x.double() * 0.0442, dim=1)
        for dt, tol in ((torch.float32, 2e-6), (torch.bfloat16, 4e-3)):
            y = torch.empty(rows, n, device="cuda", dtype=dt)
            abi.softmax_rows(x, y, 0.0442)
            assert rel_l2(y, ref) < tol
