"""GPU, BASELINE.json's full size (1.44 B-parameter U-Net, latent 4x16x40x64, context 333x1024):
one forward against the CPU oracle in both precision modes, plus size-independent properties
(determinism, batch consistency, CFG-step algebra) that need no oracle."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    from pandora_b200 import UNetModel, config
    from pandora_b200.synthetic import randomize_module_, synth_inputs
    with torch.device("cuda"):
        m = UNetModel(**config.UNET_512).eval()
    randomize_module_(m, 1234)
    inp = synth_inputs(1, config.LATENT_512, 1024, seed=0)
    return m, inp


def test_full_size_forward_matches_oracle_fp32_and_bf16(full):
    from oracle import unet_oracle
    from pandora_b200 import config
    from pandora_b200.synthetic import synth_state_dict
    m, inp = full
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([499])
    fs = torch.tensor([24])
    sd = synth_state_dict(unet_oracle.unet_param_shapes(config.UNET_512), 1234)
    ref = unet_oracle.unet_forward(sd, config.UNET_512, x, t, inp["context"], fs)
    del sd
    m.compute_mode = "fp32"
    y32 = m(x.cuda(), t.cuda(), context=inp["context"].cuda(), fs=fs.cuda()).cpu()
    e32 = rel_l2(y32, ref)
    m.compute_mode = "bf16"
    m.invalidate_packed_weights()
    yb = m(x.cuda(), t.cuda(), context=inp["context"].cuda(), fs=fs.cuda()).cpu()
    eb = rel_l2(yb, ref)
    print(f"full size 16x320x512: fp32 mode rel-L2 = {e32:.3e}, bf16 mode rel-L2 = {eb:.3e}")
    assert e32 < 1e-3 and eb < 2e-2


def test_full_size_matches_the_reference_itself(full):
    """Full width against outputs of the REFERENCE (tests/golden/unet_full512.npz, written by
    `oracle/make_golden.py --full` executing /root/reference/DynamiCrafter): the whole forward at t = 499, and
    teacher-forced forwards at five timesteps of the 50-step trailing schedule with the cond and the uncond context
    (x_t = q_sample(x0, t); strided fixture).  Per-step v and eps relative L2: fp32 mode <= 1e-3, bf16 <= 2e-2."""
    from helpers import load_golden
    from pandora_b200 import LatentDiffusionLite, config
    from pandora_b200.synthetic import synth_inputs
    z, meta = load_golden("unet_full512")
    m, _ = full
    inp = synth_inputs(1, tuple(meta["latent"]), 1024, meta["seed_x"])
    noise = synth_inputs(1, tuple(meta["latent"]), 1024, meta["noise_seed"])["x"]
    sched = LatentDiffusionLite(m, **config.DIFFUSION_512)
    fs = torch.tensor([meta["fs"]], device="cuda")
    z_c = inp["c_concat"].cuda()
    stride = meta["stride"]
    for mode, tol in (("fp32", 1e-3), ("bf16", 2e-2)):
        m.compute_mode = mode
        m.invalidate_packed_weights()
        y = m(inp["x"].cuda(), torch.tensor([499], device="cuda"), context=inp["context"].cuda(), fs=fs, x_concat=[z_c])
        e_full = rel_l2(y, z["y"])
        worst_v = worst_eps = 0.0
        for tv in meta["steps"]:
            sa = float(sched.sqrt_alphas_cumprod.float()[tv]); so = float(sched.sqrt_one_minus_alphas_cumprod.float()[tv])
            xt = sa * inp["x"] + so * noise
            for tag, ctx in (("c", inp["context"]), ("u", inp["uc_context"])):
                v = m(xt.cuda(), torch.tensor([tv], device="cuda"), context=ctx.cuda(), fs=fs, x_concat=[z_c]).float().cpu()
                v_s, xt_s = v.flatten()[::stride], xt.flatten()[::stride]
                v_ref = torch.from_numpy(z[f"v_{tag}_{tv}"])
                worst_v = max(worst_v, rel_l2(v_s, v_ref))
                worst_eps = max(worst_eps, rel_l2(sa * v_s + so * xt_s, sa * v_ref + so * xt_s))   # ddpm3d.py:243-247
        print(f"full width vs the reference, {mode} mode: forward rel-L2 {e_full:.3e}; teacher-forced worst v {worst_v:.3e}, "
              f"eps {worst_eps:.3e} over t = {meta['steps']} x (cond, uncond)")
        assert e_full < tol and worst_v < tol and worst_eps < tol
    m.compute_mode = None


def test_full_size_50_step_clip_bf16_vs_fp32_psnr(full):
    """BASELINE config 2 at its stated size and tolerance: a free-running 50-step CFG 7.5 / rescale 0.7 DDIM clip in
    bf16 mode against the same clip in fp32 mode (same x_T, same per-step noise): final-latent PSNR >= 40 dB."""
    from helpers import psnr
    from pandora_b200 import DDIMSampler, LatentDiffusionLite, config
    m, inp = full
    d = {k: v.cuda() for k, v in inp.items()}
    cond = dict(c_crossattn=[d["context"]], c_concat=[d["c_concat"]])
    uc = dict(c_crossattn=[d["uc_context"]], c_concat=[d["c_concat"]])
    model = LatentDiffusionLite(m, **config.DIFFUSION_512).cuda()
    outs = {}
    for mode in ("fp32", "bf16"):
        m.compute_mode = mode
        m.invalidate_packed_weights()
        torch.manual_seed(11)
        x, _ = DDIMSampler(model).sample(S=50, batch_size=1, shape=config.LATENT_512, conditioning=cond,
                                         unconditional_conditioning=uc, unconditional_guidance_scale=7.5, eta=1.0,
                                         x_T=d["x"], fs=torch.tensor([24], device="cuda"), verbose=False,
                                         timestep_spacing="uniform_trailing", guidance_rescale=0.7)
        outs[mode] = x.float().cpu()
    m.compute_mode = None
    m.invalidate_packed_weights()
    p = psnr(outs["bf16"], outs["fp32"])
    print(f"full-size 50-step clip, bf16 vs fp32 mode: PSNR = {p:.1f} dB, rel-L2 = {rel_l2(outs['bf16'], outs['fp32']):.3e}")
    assert torch.isfinite(outs["bf16"]).all() and p >= 40.0


def test_full_size_determinism_and_batch_consistency(full):
    m, inp = full
    x = inp["x"].cuda()
    z = inp["c_concat"].cuda()
    ctx = inp["context"].cuda()
    t = torch.tensor([259], device="cuda")
    # a batch of two different samples == the two samples run alone
    x2 = torch.cat([x, x.flip(3)], 0)
    z2 = torch.cat([z, z.flip(4)], 0)
    ctx2 = torch.cat([ctx, ctx.flip(1)], 0)
    t2 = torch.tensor([259, 19], device="cuda")
    res = {}
    for mode in ("fp32", "bf16"):
        m.compute_mode = mode
        y1 = m(x, t, context=ctx, x_concat=[z])
        y2 = m(x, t, context=ctx, x_concat=[z])
        assert torch.equal(y1, y2)                      # run-to-run deterministic (no atomics on the data path)
        assert torch.isfinite(y1).all()
        yb = m(x2, t2, context=ctx2, x_concat=[z2])
        ya = m(x2[1:], t2[1:], context=ctx2[1:].contiguous(), x_concat=[z2[1:]])
        res[mode] = (y1[0], ya[0], yb)
    # fp32 mode: the batched run is the single runs to fp32 round-off (tile shapes / split-K ways follow M,
    # so the summation order may differ, nothing else)
    y1, ya, yb = res["fp32"]
    assert rel_l2(yb[0], y1) < 1e-5 and rel_l2(yb[1], ya) < 1e-5
    # bf16 mode: a different summation order flips bf16 roundings, which decorrelates the rounding noise of
    # the ~1200 kernels downstream; both runs must sit inside the bf16 budget of the SAME fp32 answer
    r1, ra, _ = res["fp32"]
    y1, ya, yb = res["bf16"]
    errs = [rel_l2(y1, r1), rel_l2(yb[0], r1), rel_l2(ya, ra), rel_l2(yb[1], ra)]
    print("bf16 vs fp32 mode, single / batched:", ["%.3e" % e for e in errs])
    assert max(errs) < 2e-2
    assert rel_l2(yb[0], y1) < 2e-2 and rel_l2(yb[1], ya) < 2e-2


def test_full_size_cfg_step_algebra(full):
    """With identical cond/uncond conditioning the CFG combine must be a no-op for any scale, and a full
    graph-replayed step must equal the eager step."""
    from pandora_b200 import DDIMSampler, LatentDiffusionLite, config
    m, inp = full
    m.compute_mode = "bf16"
    model = LatentDiffusionLite(m, **config.DIFFUSION_512).cuda()
    d = {k: v.cuda() for k, v in inp.items()}
    cond = dict(c_crossattn=[d["context"]], c_concat=[d["c_concat"]])
    outs = []
    for graph in (False, True):
        for scale in (1.0, 7.5):
            s = DDIMSampler(model)
            s.use_cuda_graph = graph
            torch.manual_seed(3)
            x, _ = s.sample(S=5, batch_size=1, shape=config.LATENT_512, conditioning=cond,
                            unconditional_conditioning=cond, unconditional_guidance_scale=scale, eta=1.0,
                            x_T=d["x"], verbose=False, timestep_spacing="uniform_trailing")
            outs.append(x.float())
    assert rel_l2(outs[2], outs[0]) < 1e-6      # graph replay == eager (scale 1: one forward per step)
    assert rel_l2(outs[3], outs[1]) < 1e-6      # graph replay == eager (both run the batched cond|uncond forward)
    # CFG with uc == c is the identity up to fp32 round-off of the combine; scale 1 runs b=1 forwards and
    # scale 7.5 the batched b=2 forward, whose bf16 rounding noise is decorrelated (see the test above)
    assert rel_l2(outs[1], outs[0]) < 2e-2


# ------------------------------------------------------------------------------------------ 576 x 1024
@pytest.fixture(scope="module")
def full1024():
    """The same 1.44 B-parameter U-Net at BASELINE config 3's latent (4 x 16 x 72 x 128, 9216 tokens per frame)."""
    from pandora_b200 import UNetModel, config
    from pandora_b200.synthetic import randomize_module_, synth_inputs
    with torch.device("cuda"):
        m = UNetModel(**config.UNET_1024).eval()
    randomize_module_(m, 1234)
    inp = synth_inputs(1, config.LATENT_1024, 1024, seed=0)
    return m, inp


def test_1024_forward_matches_oracle_fp32_and_bf16(full1024):
    """One forward at 16x576x1024 against the CPU oracle (its attention walks the frames one at a time to bound the
    score matrix): fp32 mode <= 1e-3, bf16 mode <= 2e-2 (north_star tolerances), and bf16 vs fp32 mode."""
    from oracle import unet_oracle
    from pandora_b200 import config
    from pandora_b200.synthetic import synth_state_dict
    m, inp = full1024
    x = torch.cat([inp["x"], inp["c_concat"]], 1)
    t = torch.tensor([499])
    fs = torch.tensor([10])
    sd = synth_state_dict(unet_oracle.unet_param_shapes(config.UNET_1024), 1234)
    ref = unet_oracle.unet_forward(sd, config.UNET_1024, x, t, inp["context"], fs)
    del sd
    m.compute_mode = "fp32"
    y32 = m(x.cuda(), t.cuda(), context=inp["context"].cuda(), fs=fs.cuda()).cpu()
    e32 = rel_l2(y32, ref)
    m.compute_mode = "bf16"
    m.invalidate_packed_weights()
    yb = m(x.cuda(), t.cuda(), context=inp["context"].cuda(), fs=fs.cuda()).cpu()
    eb, ebf = rel_l2(yb, ref), rel_l2(yb, y32)
    print(f"full size 16x576x1024: fp32 mode rel-L2 = {e32:.3e}, bf16 mode rel-L2 = {eb:.3e} (vs fp32 mode {ebf:.3e})")
    assert e32 < 1e-3 and eb < 2e-2 and ebf < 2e-2


@pytest.mark.parametrize("P", [2, 4])
def test_1024_virtual_sequence_parallel_matches_single_gpu(full1024, P):
    """BASELINE config 3 on one GPU: the 576x1024 clip sharded over P virtual ranks (frames for the spatial layers,
    pixels for the temporal ones; at P = 4 the 9x16 level holds 36 pixels per rank and 4 frames of 9x16 = ragged
    bricks) equals the unsharded forward.  bf16 mode: both are held to the bf16 budget against each other."""
    from pandora_b200.dist import VirtualSequenceParallel
    m, inp = full1024
    d = {k: v.cuda() for k, v in inp.items()}
    t = torch.tensor([259], device="cuda")
    m.compute_mode = "bf16"
    ref = m(d["x"], t, context=d["context"], x_concat=[d["c_concat"]]).float()
    hub = VirtualSequenceParallel(P)
    m.sequence_parallel = hub
    try:
        outs = hub.run(lambda r: m(d["x"], t, context=d["context"], x_concat=[d["c_concat"]]).float())
    finally:
        m.sequence_parallel = None
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    err = rel_l2(outs[0], ref)
    print(f"16x576x1024 bf16, {P} virtual sequence-parallel ranks vs single GPU: rel-L2 = {err:.3e}")
    assert torch.isfinite(outs[0]).all() and err < 2e-2
