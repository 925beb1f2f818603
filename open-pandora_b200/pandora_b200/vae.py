"""B200-native first-stage DECODER: the KL-VAE that turns the sampler's latents into frames.

Reference (paths relative to /root/reference/DynamiCrafter):
  lvdm/models/autoencoder.py:14-107          AutoencoderKL (constructor kwargs, state-dict layout, decode = post_quant_conv -> Decoder)
  lvdm/modules/networks/ae_modules.py:466-579  Decoder; :152-213 ResnetBlock; :26-76 AttnBlock; :112-126 Upsample
  lvdm/models/ddpm3d.py:630-655              LatentDiffusion.decode_first_stage: z / scale_factor, frame by frame

Drop-in surface: `AutoencoderKL(ddconfig, lossconfig, embed_dim, ...)` keeps the reference constructor's keyword list and
reproduces its state-dict key names and shapes (encoder.*, decoder.*, quant_conv.*, post_quant_conv.*), so a reference
checkpoint loads with strict=True; `decode(z)` is the native path; `swap_vae_decoder(latent_diffusion)` re-routes an
existing reference model's `first_stage_model.decode` through it, sharing the Parameter objects.  The encoder runs once
per round on one image and stays with the reference (its parameters are held, `encode` raises).

Execution reuses the denoiser's kernels through the C ABI, in the same rows layout (n, y, x) x channels:
  conv3x3 / 1x1 / nearest-x2 + conv      pb_tc_gemm implicit GEMM (bias / residual epilogues, GroupNorm column statistics)
  GroupNorm(32, eps 1e-6) + swish          pb_groupnorm_* (swish = SiLU)
  AttnBlock (one head, d = C = 512)        three GEMMs per frame around pb_softmax_rows: S = q k^T (fp32), P = softmax(S C^-1/2),
                                           O = P v with v^T produced directly by a GEMM with swapped operands; the value bias is
                                           added after P v (softmax rows sum to one)
Precision modes as in the U-Net: fp32 = exact CUDA-core kernels; bf16 = bf16 operands on the tcgen05 kernels, fp32 trunk.
"""
import torch
import torch.nn as nn

from . import abi
from .unet import _Runner


class _Res(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.in_channels, self.out_channels = cin, cout
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.in_channels = c
        self.norm = nn.GroupNorm(32, c, eps=1e-6)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))


class _Conv(nn.Module):
    def __init__(self, c, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1 if stride == 1 else 0)


def _mid(c):
    m = nn.Module()
    m.block_1, m.attn_1, m.block_2 = _Res(c, c), _Attn(c), _Res(c, c)
    return m


class Decoder(nn.Module):
    """Parameter tree of ae_modules.Decoder (execution: `_DecodeRunner`)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if list(attn_resolutions) or use_linear_attn or attn_type != "vanilla" or give_pre_end or tanh_out \
                or not resamp_with_conv:
            raise NotImplementedError("pandora_b200 Decoder covers the DynamiCrafter first-stage configuration "
                                      "(no attention in the up path, vanilla mid attention, conv resampling)")
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, padding=1)
        self.mid = _mid(block_in)
        self.up = nn.ModuleList()
        for lv in reversed(range(self.num_resolutions)):
            up = nn.Module()
            up.block, up.attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[lv]
            for _ in range(num_res_blocks + 1):
                up.block.append(_Res(block_in, block_out))
                block_in = block_out
            if lv != 0:
                up.upsample = _Conv(block_in)
            self.up.insert(0, up)
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, padding=1)


class Encoder(nn.Module):
    """Parameter holder for ae_modules.Encoder (:364-463): state-dict compatibility only."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, **ignore_kwargs):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for lv in range(len(ch_mult)):
            down = nn.Module()
            down.block, down.attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_mult[lv], ch * ch_mult[lv]
            for _ in range(num_res_blocks):
                down.block.append(_Res(block_in, block_out))
                block_in = block_out
            if lv != len(ch_mult) - 1:
                down.downsample = _Conv(block_in, stride=2)
            self.down.append(down)
        self.mid = _mid(block_in)
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, padding=1)


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig=None, embed_dim=4, ckpt_path=None, ignore_keys=(), image_key="image",
                 colorize_nlabels=None, monitor=None, test=False, test_args=None, logdir=None, input_dim=4, **kwargs):
        super().__init__()
        assert ddconfig["double_z"]
        self.image_key, self.input_dim, self.embed_dim = image_key, input_dim, embed_dim
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        #: None = follow the weights / input dtype; or "fp32" / "bf16"
        self.compute_mode = None
        self.kernel_impl = "auto"
        self.bf16_norm_inputs = True
        #: frames decoded per pass (bounds the activation memory: one 576x1024 frame is 0.3 GB per 128-channel tensor)
        self.frames_per_pass = 8
        for name, m in self.named_modules():
            m._pfx = name + "." if name else ""
        self._packs = {}
        if ckpt_path is not None:
            sd = torch.load(ckpt_path, map_location="cpu")
            sd = sd.get("state_dict", sd)
            self.load_state_dict({k: v for k, v in sd.items() if not any(k.startswith(i) for i in ignore_keys)},
                                 strict=False)

    def _apply(self, fn, *a, **k):
        self._packs = {}
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._packs = {}
        return super().load_state_dict(*a, **k)

    def invalidate_packed_weights(self):
        self._packs = {}

    def encode(self, x, **kwargs):
        raise NotImplementedError("the first-stage ENCODER is outside the native path (it runs once per round on one "
                                  "image); keep the reference encoder and use swap_vae_decoder() for decode")

    @torch.no_grad()
    def decode(self, z, scale=1.0, **kwargs):
        """z: (n, z_channels, h, w) -> (n, out_ch, 8h, 8w) in z.dtype (autoencoder.py:104-107).  `scale` multiplies z
        first (decode_first_stage passes 1 / scale_factor; folded into the 1x1 post_quant_conv weights)."""
        if not z.is_cuda:
            raise RuntimeError("pandora_b200.AutoencoderKL.decode runs on CUDA only (no CPU fallback)")
        outs = []
        for i in range(0, z.shape[0], self.frames_per_pass):
            outs.append(_DecodeRunner(self, z[i:i + self.frames_per_pass].contiguous(), float(scale)).run())
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    @torch.no_grad()
    def decode_first_stage(self, z, scale_factor):
        """LatentDiffusion.decode_core (ddpm3d.py:630-651): video latents (b, c, t, h, w) or images (n, c, h, w)."""
        if z.dim() == 5:
            b, c, t, h, w = z.shape
            y = self.decode(z.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w), scale=1.0 / scale_factor)
            return y.reshape(b, t, *y.shape[1:]).permute(0, 2, 1, 3, 4)
        return self.decode(z, scale=1.0 / scale_factor)

    @classmethod
    def from_reference(cls, ref, ddconfig, embed_dim=4):
        """Build from an instantiated reference AutoencoderKL, sharing its Parameter objects."""
        m = cls(ddconfig=ddconfig, embed_dim=embed_dim)
        for name, p in ref.named_parameters():
            if name.startswith("loss."):
                continue
            mod_name, _, leaf = name.rpartition(".")
            setattr(m.get_submodule(mod_name), leaf, p)
        m.invalidate_packed_weights()
        return m


def swap_vae_decoder(latent_diffusion, ddconfig, embed_dim=4):
    """Route `latent_diffusion.first_stage_model.decode` through the native decoder (shared parameters); encode and
    everything else stay with the reference object."""
    ref = latent_diffusion.first_stage_model
    native = AutoencoderKL.from_reference(ref, ddconfig, embed_dim)
    ref.decode = lambda z, **kw: native.decode(z)
    return native


class _DecodeRunner(_Runner):
    """One decode pass over n frames; reuses the denoiser executor's buffer / weight-pack / kernel wrappers."""

    def __init__(self, model, z, scale):
        self.m, self.z, self.scale = model, z, scale
        self.dev = z.device
        p0 = next(model.parameters())
        if p0.device != z.device:
            raise RuntimeError(f"model on {p0.device}, input on {z.device}")
        mode = model.compute_mode
        if mode is None:
            mode = "bf16" if (p0.dtype == torch.bfloat16 or z.dtype in (torch.bfloat16, torch.float16)) else "fp32"
        self.mode = mode
        self.opdt = torch.bfloat16 if mode == "bf16" else torch.float32
        self.use_tc = mode == "bf16" and model.kernel_impl == "auto" and abi.supports_tc()
        self.pack = model._packs.setdefault(mode, {})
        self.sd = None
        self.sp = None
        self.n, _, self.H, self.W = z.shape
        self.b, self.T, self.Tf = self.n, 1, 1

    def res(self, blk, x, H, W):
        p = blk._pfx
        has_skip = blk.in_channels != blk.out_channels
        x_op = self.new(x.shape[0], x.shape[1], self.opdt) if (has_skip and x.dtype != self.opdt) else None
        a = self.gn(x, p + "norm1.", self.n, 1e-6, True, x_copy=x_op)
        mid_dt = self.opdt if (self.use_tc and self.m.bf16_norm_inputs) else torch.float32
        h = self.conv3x3(a, p + "conv1.weight", p + "conv1.bias", H, W, mid_dt, stats=True)
        a = self.gn(h, p + "norm2.", self.n, 1e-6, True)
        skip = x
        if has_skip:
            skip = self.linear(x if x_op is None else x_op, p + "nin_shortcut.weight", p + "nin_shortcut.bias", torch.float32)
        return self.conv3x3(a, p + "conv2.weight", p + "conv2.bias", H, W, torch.float32, residual=skip, stats=True)

    def attn(self, blk, x, H, W):
        """ae_modules.py:53-76: single head, d = C."""
        p = blk._pfx
        C, HW = x.shape[1], H * W
        a = self.gn(x, p + "norm.", self.n, 1e-6, False)
        q = self.linear(a, p + "q.weight", p + "q.bias", self.opdt)
        k = self.linear(a, p + "k.weight", p + "k.bias", self.opdt)
        wv = self.lin_w(p + "v.weight")                         # [C, C]: the A operand of the swapped product
        o = self.new(x.shape[0], C, self.opdt)
        s = self.new(HW, HW, torch.float32)
        pr = self.new(HW, HW, self.opdt)
        vt = self.new(C, HW, self.opdt)
        for f in range(self.n):
            rows = slice(f * HW, (f + 1) * HW)
            self.gemm(wv, a[rows], HW, self.opdt, out=vt)                      # v^T = W_v a^T (bias deferred)
            self.gemm(q[rows], k[rows], HW, torch.float32, out=s)              # S = q k^T
            abi.softmax_rows(s, pr, float(C) ** -0.5)
            self.gemm(pr, vt, C, self.opdt, bias=self.vec(p + "v.bias"), out=o[rows])   # P v + b_v (rows of P sum to 1)
        return self.linear(o, p + "proj_out.weight", p + "proj_out.bias", torch.float32, residual=x, stats=True)

    def run(self):
        m, dec = self.m, self.m.decoder
        n, H, W = self.n, self.H, self.W
        zc = self.z.shape[1]
        z = self.z if self.z.dtype in (torch.float32, torch.bfloat16) else self.z.float()
        h = self.new(n * H * W, zc, torch.float32)
        abi.video_to_rows(z.reshape(n, zc, 1, H, W), None, h, n, 1, H * W)
        f32 = torch.float32
        # post_quant_conv (1x1) with the 1 / scale_factor of decode_first_stage folded into its weights
        wq = self.w(("pq", self.scale), lambda: (self._param("post_quant_conv.weight").flatten(1).float() * self.scale))
        h = self.gemm(h, wq, wq.shape[0], f32, bias=self.vec("post_quant_conv.bias"), tc_ok=False)
        h = self.conv3x3(h, "decoder.conv_in.weight", "decoder.conv_in.bias", H, W, f32, wdtype=f32, stats=True)
        h = self.res(dec.mid.block_1, h, H, W)
        h = self.attn(dec.mid.attn_1, h, H, W)
        h = self.res(dec.mid.block_2, h, H, W)
        for lv in reversed(range(dec.num_resolutions)):
            for blk in dec.up[lv].block:
                h = self.res(blk, h, H, W)
            if lv != 0:
                q = dec.up[lv].upsample._pfx
                h = self.conv3x3(h, q + "conv.weight", q + "conv.bias", H, W, f32, up=True, stats=True)
                H, W = 2 * H, 2 * W
        oc = dec.conv_out.out_channels
        if self.use_tc:
            # out_ch = 3 on the tensor cores: output channels zero-padded to one 16-wide UMMA tile
            w16 = self.w(("convpad16", "decoder.conv_out.weight"), lambda: torch.cat(
                [self.conv_w("decoder.conv_out.weight"),
                 torch.zeros(16 - oc, 9 * h.shape[1], dtype=self.opdt, device=self.dev)], 0))
            b16 = self.w(("vecpad16", "decoder.conv_out.bias"), lambda: torch.cat(
                [self.vec("decoder.conv_out.bias"), torch.zeros(16 - oc, dtype=f32, device=self.dev)], 0))
            a = self.gn(h, "decoder.norm_out.", n, 1e-6, True)
            y = self.gemm(a, w16, 16, f32, bias=b16, mode=abi.A_CONV3X3, conv=(h.shape[1], H, W, H, W, 1, 0), M=n * H * W)
        else:
            # exact path: the CUDA-core GEMM moves 4-wide vectors -> pad the 3 output channels to 4
            w4 = self.w(("convpad4", "decoder.conv_out.weight"), lambda: torch.cat(
                [self.conv_w("decoder.conv_out.weight", f32), torch.zeros(4 - oc % 4 if oc % 4 else 0, 9 * h.shape[1],
                                                                          dtype=f32, device=self.dev)], 0))
            b4 = self.w(("vecpad4", "decoder.conv_out.bias"), lambda: torch.cat(
                [self.vec("decoder.conv_out.bias"), torch.zeros(w4.shape[0] - oc, dtype=f32, device=self.dev)], 0))
            a = self.gn(h, "decoder.norm_out.", n, 1e-6, True, out_dtype=f32)
            y = self.gemm(a, w4, w4.shape[0], f32, bias=b4, mode=abi.A_CONV3X3, conv=(h.shape[1], H, W, H, W, 1, 0),
                          M=n * H * W, tc_ok=False)
        out = torch.empty((n, oc, 1, H, W), dtype=z.dtype, device=self.dev)
        abi.rows_to_video(y, out, n, oc, 1, H * W)          # reads the first oc columns of every row
        return out[:, :, 0]
