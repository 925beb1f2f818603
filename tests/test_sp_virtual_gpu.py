"""Single GPU: sequence parallelism with P VIRTUAL ranks (threads of one process, `VirtualSequenceParallel`).

The ranks run the real sharded forward — frames sharded for the per-frame layers, pixels for the temporal layers —
and exchange through the same kernels as the multi-GPU p2p transport (`pb_sp_reshard`, `pb_sp_groupnorm_affine`),
ordered by host barriers instead of device flags.  This exercises, on the driver's 1-GPU box, everything of
BASELINE.json config 3 except the NVLink hop itself: the frame<->pixel addressing, ragged shards (pixel counts no
128-row brick tiles), the clip-wide GroupNorm sums, cond|uncond batching under SP.  Reference: the two shardings are
`openaimodel3d.py:570` ((b t) c h w) and `attention.py:369-372` ((b h w) t c); the sharded result must equal the
unsharded one.
"""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _hub(P):
    from pandora_b200.dist import VirtualSequenceParallel
    return VirtualSequenceParallel(P)


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("samples", [1, 2])
def test_reshard_kernel_roundtrip_and_layout(P, samples):
    T, HW, C = 16, 36 * P, 320                      # 36 pixels per rank: not a power of two
    g = torch.Generator(device="cuda").manual_seed(P * 10 + samples)
    full = torch.randn(samples, T, HW, C, device="cuda", generator=g)
    hub = _hub(P)
    tl, hwl = T // P, HW // P

    def work(r):
        sp = hub.current()
        sp.prepare(samples * tl * HW * C, 1, samples, full.device)
        mine = full[:, r * tl:(r + 1) * tl].reshape(-1, C).contiguous()
        pix = sp.frames_to_pixels(mine, samples, tl, HW)
        want = full[:, :, r * hwl:(r + 1) * hwl].reshape(-1, C)
        ok = torch.equal(pix, want)
        # back, into the right-hand columns of a wider buffer (a skip-concatenation half: row pitch > cols)
        wide = torch.zeros(samples * tl * HW, C + 64, device="cuda")
        back = sp.pixels_to_frames(pix, samples, tl, HW, out=wide[:, 64:])
        ok = ok and back.data_ptr() == wide[:, 64:].data_ptr() and torch.equal(wide[:, 64:], mine)
        ok = ok and bool((wide[:, :64] == 0).all())
        # a producer that writes straight into the send slot is not copied again
        slot = sp.send_buffer(mine.shape[0], C, full.device)
        slot.copy_(mine * 2)
        pix2 = sp.frames_to_pixels(slot, samples, tl, HW)
        return ok and torch.equal(pix2, want * 2)

    assert all(hub.run(work))


@pytest.mark.parametrize("P", [2, 4])
def test_clip_wide_groupnorm_over_virtual_ranks(P):
    from pandora_b200 import abi
    b, T, HWL, C = 2, 16, 20, 128
    g = torch.Generator(device="cuda").manual_seed(7)
    full = torch.randn(b, T, P * HWL, C, device="cuda", generator=g) * 2 + 0.5
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    ref = torch.nn.functional.group_norm(full.double().permute(0, 3, 1, 2), 32, gamma.double(), beta.double(), 1e-5)
    ref = torch.nn.functional.silu(ref).permute(0, 2, 3, 1)          # (b, T, HW, C)
    hub = _hub(P)

    def work(r):
        sp = hub.current()
        sp.prepare(b * T * HWL * C, 1, b, full.device)
        x = full[:, :, r * HWL:(r + 1) * HWL].reshape(-1, C).contiguous()      # pixel-sharded rows (b, t, p_local)
        errs = []
        for _ in range(3):                                                     # alternates the two sums slots
            y = torch.empty_like(x)
            sp.groupnorm(abi, x, y, b, T * HWL, gamma, beta, 1e-5, True)
            errs.append(rel_l2(y, ref[:, :, r * HWL:(r + 1) * HWL].reshape(-1, C)))
        return max(errs)

    assert max(hub.run(work)) < 2e-6


def _small(dtype, latent, seed=21):
    from pandora_b200 import UNetModel, config
    from pandora_b200.synthetic import randomize_module_, synth_inputs
    p = config.small_unet()
    m = randomize_module_(UNetModel(**p).eval(), seed).to(device="cuda", dtype=dtype)
    inp = {k: v.cuda() for k, v in synth_inputs(1, latent, p["context_dim"], seed=5).items()}
    return p, m, inp


@pytest.mark.parametrize("P,latent", [(2, (4, 16, 24, 16)), (4, (4, 16, 16, 16)), (2, (4, 16, 16, 16))])
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_virtual_sequence_parallel_unet_matches_single_rank(P, latent, mode, tol):
    dtype = torch.bfloat16 if mode == "bf16" else torch.float32
    p, m, inp = _small(dtype, latent)
    t = torch.tensor([499], device="cuda")
    ref = m(inp["x"], t, context=inp["context"], x_concat=[inp["c_concat"]]).float()
    hub = _hub(P)
    m.sequence_parallel = hub
    outs = hub.run(lambda r: m(inp["x"], t, context=inp["context"], x_concat=[inp["c_concat"]]).float())
    m.sequence_parallel = None
    for o in outs[1:]:
        assert torch.equal(o, outs[0])              # every rank ends with the same gathered clip
    err = rel_l2(outs[0], ref)
    print(f"virtual SP P={P} latent={latent[2]}x{latent[3]} {mode}: rel-L2 vs single rank = {err:.3e}")
    assert err < tol


def test_virtual_sequence_parallel_batched_cfg_sampler_matches_single_rank():
    """cond|uncond batched in one forward (b = 2) under sequence parallelism, through the sampler's eager CFG step
    (eta = 0: the per-step noise, which virtual ranks would draw in a racy order, is multiplied by sigma = 0)."""
    from pandora_b200 import DDIMSampler, LatentDiffusionLite, config
    p, m, inp = _small(torch.float32, (4, 16, 16, 16))
    model = LatentDiffusionLite(m, **config.DIFFUSION_512).cuda()
    cond = dict(c_crossattn=[inp["context"]], c_concat=[inp["c_concat"]])
    uc = dict(c_crossattn=[inp["uc_context"]], c_concat=[inp["c_concat"]])

    def sample():
        s = DDIMSampler(model)
        s.use_cuda_graph = False
        x, _ = s.sample(S=4, batch_size=1, shape=(4, 16, 16, 16), conditioning=cond, unconditional_conditioning=uc,
                        unconditional_guidance_scale=7.5, eta=0.0, x_T=inp["x"], verbose=False,
                        timestep_spacing="uniform_trailing", guidance_rescale=0.7)
        return x.float()

    ref = sample()
    hub = _hub(2)
    m.sequence_parallel = hub
    outs = hub.run(lambda r: sample())
    m.sequence_parallel = None
    assert torch.equal(outs[0], outs[1])
    assert rel_l2(outs[0], ref) < 1e-4
