"""ORACLE (test infrastructure, not product code): CPU restatement of the first-stage (KL-VAE) DECODER that turns the
sampler's latents into frames — SURVEY.md section 8(f) row 1.

Functional plain-PyTorch restatement operating on a state dict with the reference's key names.  Only `tests/`,
`__graft_entry__.smoke()` and bench.py's CPU legs may import it.

Parity pinning: `oracle/make_golden.py --vae` executes the reference's own `AutoencoderKL.decode`
(/root/reference/DynamiCrafter/lvdm/models/autoencoder.py) in the build container and stores its outputs in
`tests/golden/vae_*.npz`; `tests/test_oracle_golden.py` holds this restatement to them.

Reference anchors (paths relative to /root/reference/DynamiCrafter):
  AutoencoderKL.decode          lvdm/models/autoencoder.py:104-107      post_quant_conv (1x1) -> Decoder
  Decoder.forward               lvdm/modules/networks/ae_modules.py:540-579
  ResnetBlock.forward           ae_modules.py:190-213   (temb is None in the decoder: temb_ch = 0)
  AttnBlock.forward             ae_modules.py:53-76     single head, d = C, scale C^-0.5
  Upsample.forward              ae_modules.py:122-126   nearest x2 + conv3x3
  Normalize                     ae_modules.py:15-16     GroupNorm(32, C, eps=1e-6)
  LatentDiffusion.decode_core   lvdm/models/ddpm3d.py:630-651  z / scale_factor, per frame, (b t) c h w <-> b c t h w
"""
import torch
import torch.nn.functional as F


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + "weight"], sd[p + "bias"], padding=pad)


def resnet_block(sd, p, x):
    h = _conv(sd, p + "conv1.", _swish(_gn(sd, p + "norm1.", x)), 1)
    h = _conv(sd, p + "conv2.", _swish(_gn(sd, p + "norm2.", h)), 1)
    if p + "nin_shortcut.weight" in sd:
        x = _conv(sd, p + "nin_shortcut.", x, 0)
    return x + h


def attn_block(sd, p, x):
    h = _gn(sd, p + "norm.", x)
    q, k, v = (_conv(sd, p + n + ".", h, 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out.", h, 0)


def decoder_layout(sd):
    """(number of resolution levels, blocks per level) read off the key names."""
    levels = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.up."))
    blocks = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("decoder.up.0.block."))
    return levels, blocks


def decode(sd, z):
    """AutoencoderKL.decode: z (n, z_channels, h, w) -> images (n, 3, 8h, 8w)."""
    h = _conv(sd, "post_quant_conv.", z, 0)
    h = _conv(sd, "decoder.conv_in.", h, 1)
    h = resnet_block(sd, "decoder.mid.block_1.", h)
    h = attn_block(sd, "decoder.mid.attn_1.", h)
    h = resnet_block(sd, "decoder.mid.block_2.", h)
    levels, blocks = decoder_layout(sd)
    for lv in reversed(range(levels)):
        for i in range(blocks):
            h = resnet_block(sd, f"decoder.up.{lv}.block.{i}.", h)
        if lv != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{lv}.upsample.conv.", h, 1)
    h = _swish(_gn(sd, "decoder.norm_out.", h))
    return _conv(sd, "decoder.conv_out.", h, 1)


def decode_first_stage(sd, z, scale_factor):
    """LatentDiffusion.decode_core with perframe_ae (the arithmetic is per frame either way)."""
    if z.dim() == 5:
        b, c, t, hh, ww = z.shape
        zz = z.permute(0, 2, 1, 3, 4).reshape(b * t, c, hh, ww)
        out = torch.cat([decode(sd, (1.0 / scale_factor) * zz[i:i + 1]) for i in range(b * t)], 0)
        return out.reshape(b, t, *out.shape[1:]).permute(0, 2, 1, 3, 4)
    return torch.cat([decode(sd, (1.0 / scale_factor) * z[i:i + 1]) for i in range(z.shape[0])], 0)


def vae_param_shapes(ddconfig, embed_dim=4):
    """The reference AutoencoderKL's decoder-side parameter inventory (decoder.*, post_quant_conv.*)."""
    ch, mult, nrb, zc = ddconfig["ch"], list(ddconfig["ch_mult"]), ddconfig["num_res_blocks"], ddconfig["z_channels"]
    out_ch = ddconfig["out_ch"]
    s = {}

    def conv(p, i, o, k):
        s[p + "weight"] = (o, i, k, k)
        s[p + "bias"] = (o,)

    def norm(p, c):
        s[p + "weight"] = (c,)
        s[p + "bias"] = (c,)

    def res(p, i, o):
        norm(p + "norm1.", i); conv(p + "conv1.", i, o, 3); norm(p + "norm2.", o); conv(p + "conv2.", o, o, 3)
        if i != o:
            conv(p + "nin_shortcut.", i, o, 1)

    conv("post_quant_conv.", embed_dim, zc, 1)
    block_in = ch * mult[-1]
    conv("decoder.conv_in.", zc, block_in, 3)
    res("decoder.mid.block_1.", block_in, block_in)
    p = "decoder.mid.attn_1."
    norm(p + "norm.", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(p + n + ".", block_in, block_in, 1)
    res("decoder.mid.block_2.", block_in, block_in)
    for lv in reversed(range(len(mult))):
        block_out = ch * mult[lv]
        for i in range(nrb + 1):
            res(f"decoder.up.{lv}.block.{i}.", block_in, block_out)
            block_in = block_out
        if lv != 0:
            conv(f"decoder.up.{lv}.upsample.conv.", block_in, block_in, 3)
    norm("decoder.norm_out.", block_in)
    conv("decoder.conv_out.", block_in, out_ch, 3)
    return s
