"""ORACLE (test infrastructure, not product code): CPU restatement of the denoiser hot path.

A functional, plain-PyTorch fp32/fp64 restatement of DynamiCrafter's `UNetModel.forward` as used by
Open-Pandora, written against the reference's behaviour, operating directly on a state dict that
uses the reference's key names.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this file; the product package never does.

Parity pinning: the reference ships no tests or golden vectors for this path,
so this restatement is pinned against outputs of the reference's own modules executed in the build
container; `oracle/make_golden.py` is the generating script and `tests/golden/*.npz` the committed
fixtures (see `tests/test_oracle_golden.py`).

Reference anchors (paths relative to /root/reference/DynamiCrafter):
  timestep_embedding            lvdm/models/utils_diffusion.py:8-28   (bf16 frequency table!)
  UNetModel.__init__/forward    lvdm/modules/networks/openaimodel3d.py:314-607
  ResBlock._forward             openaimodel3d.py:213-239
  TemporalConvBlock.forward     openaimodel3d.py:275-282
  Downsample / Upsample         openaimodel3d.py:51-109
  SpatialTransformer.forward    lvdm/modules/attention.py:294-310
  TemporalTransformer.forward   attention.py:365-412
  BasicTransformerBlock._forward attention.py:242-246
  CrossAttention.forward        attention.py:81-144
  GEGLU / FeedForward           attention.py:415-442
"""
import math

import torch
import torch.nn.functional as F

TEXT_LEN = 77


# ----------------------------------------------------------------------------- architecture walk
def unet_layout(cfg):
    """Enumerate the blocks `UNetModel.__init__` instantiates (openaimodel3d.py:388-550).

    Returns dict(input=[...], middle=[...], output=[...]) where every entry is a list of layer
    tuples: ("conv", prefix, cin, cout) | ("res", prefix, cin, cout) | ("st", prefix, ch, heads) |
    ("tt", prefix, ch, heads, use_linear) | ("down", prefix, ch) | ("up", prefix, ch).
    """
    mc = cfg["model_channels"]
    nhc = cfg["num_head_channels"]
    att = set(cfg["attention_resolutions"])
    mult = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    use_linear = cfg.get("use_linear", False)
    temporal = cfg.get("temporal_attention", True)

    def attn_layers(prefix, idx, ch):
        heads = ch // nhc
        ls = [("st", f"{prefix}.{idx}.", ch, heads)]
        if temporal:
            ls.append(("tt", f"{prefix}.{idx + 1}.", ch, heads, use_linear))
        return ls

    inp = [[("conv", "input_blocks.0.0.", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            p = f"input_blocks.{len(inp)}"
            layers = [("res", f"{p}.0.", ch, m * mc)]
            ch = m * mc
            if ds in att:
                layers += attn_layers(p, 1, ch)
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", f"input_blocks.{len(inp)}.0.", ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", "middle_block.0.", ch, ch)] + attn_layers("middle_block", 1, ch)
    mid.append(("res", f"middle_block.{len(mid)}.", ch, ch))
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            p = f"output_blocks.{len(out)}"
            layers = [("res", f"{p}.0.", ch + ich, m * mc)]
            ch = m * mc
            if ds in att:
                layers += attn_layers(p, 1, ch)
            if level and i == nrb:
                layers.append(("up", f"{p}.{len(layers)}.", ch))
                ds //= 2
            out.append(layers)
    return dict(input=inp, middle=mid, output=out)


# ----------------------------------------------------------------------------------- primitives
def timestep_embedding(timesteps, dim, max_period=10000):
    """utils_diffusion.py:8-28. The frequency table is evaluated in bf16, then used in fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.bfloat16) / half)
    args = timesteps[:, None].float() * freqs[None].to(timesteps.device)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _lin(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd.get(p + "bias"))


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + "weight"], sd[p + "bias"], eps)


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], 1e-5)


def _heads(x, h):
    b, n, _ = x.shape
    return x.reshape(b, n, h, -1).permute(0, 2, 1, 3)  # b h n d


#: score-matrix elements above which `_softmax_attn` walks the batch (frames) one sample at a time.  Same arithmetic
#: (every sample's softmax is independent); it only bounds memory: at 576x1024 the reference's einsum path
#: materialises (16 frames x 5 heads) x 9216 x 9216 fp32 = 27 GB (attention.py:103-120).
ATTN_CHUNK_ELEMS = 1 << 29


#: "einsum" = the path the reference takes without xformers (materialised scores, attention.py:101-125);
#: "sdpa" = torch's fused attention as the stand-in for xformers.memory_efficient_attention (attention.py:175), used
#: only by bench.py's torch-on-GPU baseline arm.
ATTN_IMPL = "einsum"


def _softmax_attn(q, k, v, scale):
    if ATTN_IMPL == "sdpa":
        return F.scaled_dot_product_attention(q, k, v, scale=scale)
    if q.shape[0] > 1 and q.shape[0] * q.shape[1] * q.shape[2] * k.shape[2] > ATTN_CHUNK_ELEMS:
        return torch.cat([_softmax_attn(q[i:i + 1], k[i:i + 1], v[i:i + 1], scale) for i in range(q.shape[0])], 0)
    sim = torch.matmul(q, k.transpose(-1, -2)) * scale
    return torch.matmul(sim.softmax(dim=-1), v)


def cross_attention(sd, p, x, context, heads, image_cross):
    """attention.py:81-144 (einsum path; the xformers path :146-209 is the same math)."""
    q = _lin(sd, p + "to_q.", x)
    k_ip = v_ip = None
    if context is None:
        k, v = _lin(sd, p + "to_k.", x), _lin(sd, p + "to_v.", x)
    else:
        text = context[:, :TEXT_LEN]
        k, v = _lin(sd, p + "to_k.", text), _lin(sd, p + "to_v.", text)
        if image_cross:
            img = context[:, TEXT_LEN:]
            k_ip, v_ip = _lin(sd, p + "to_k_ip.", img), _lin(sd, p + "to_v_ip.", img)
    scale = (q.shape[-1] // heads) ** -0.5
    qh = _heads(q, heads)
    out = _softmax_attn(qh, _heads(k, heads), _heads(v, heads), scale)
    if k_ip is not None:
        out = out + 1.0 * _softmax_attn(qh, _heads(k_ip, heads), _heads(v_ip, heads), scale)
    b, h, n, d = out.shape
    out = out.permute(0, 2, 1, 3).reshape(b, n, h * d)
    return _lin(sd, p + "to_out.0.", out)


def transformer_block(sd, p, x, context, heads, image_cross):
    """attention.py:242-246; GEGLU feed-forward attention.py:415-442 (exact erf GELU)."""
    x = cross_attention(sd, p + "attn1.", _ln(sd, p + "norm1.", x), None, heads, False) + x
    x = cross_attention(sd, p + "attn2.", _ln(sd, p + "norm2.", x), context, heads, image_cross) + x
    y = _lin(sd, p + "ff.net.0.proj.", _ln(sd, p + "norm3.", x))
    a, gate = y.chunk(2, dim=-1)
    return _lin(sd, p + "ff.net.2.", a * F.gelu(gate)) + x


def spatial_transformer(sd, p, x, context, heads, use_linear, image_cross):
    """attention.py:294-310. x: (b t) c h w, context: (b t) L c."""
    n, c, hh, ww = x.shape
    y = _gn(sd, p + "norm.", x, 1e-6)
    if not use_linear:
        y = F.conv2d(y, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(n, hh * ww, -1)
    if use_linear:
        y = _lin(sd, p + "proj_in.", y)
    y = transformer_block(sd, p + "transformer_blocks.0.", y, context, heads, image_cross)
    if use_linear:
        y = _lin(sd, p + "proj_out.", y)
    y = y.reshape(n, hh, ww, -1).permute(0, 3, 1, 2)
    if not use_linear:
        y = F.conv2d(y, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return y + x


def temporal_transformer(sd, p, x, b, heads, use_linear):
    """attention.py:365-412 with only_self_att=True, no mask, no relative position.

    x: (b t) c h w.  GroupNorm statistics span all t frames (the reference normalises the
    5-D tensor b c t h w).  With use_linear=False (`init_attn`) the projections are Conv1d k=1.
    """
    n, c, hh, ww = x.shape
    t = n // b
    x5 = x.reshape(b, t, c, hh, ww).permute(0, 2, 1, 3, 4)  # b c t h w
    y = F.group_norm(x5, 32, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    y = y.permute(0, 3, 4, 2, 1).reshape(b * hh * ww, t, c)  # (b h w) t c
    w_in, w_out = sd[p + "proj_in.weight"], sd[p + "proj_out.weight"]
    y = F.linear(y, w_in.reshape(w_in.shape[0], -1), sd[p + "proj_in.bias"])
    y = transformer_block(sd, p + "transformer_blocks.0.", y, None, heads, False)
    y = F.linear(y, w_out.reshape(w_out.shape[0], -1), sd[p + "proj_out.bias"])
    y = y.reshape(b, hh, ww, t, c).permute(0, 4, 3, 1, 2)  # b c t h w
    y = y + x5
    return y.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)


def temporal_conv_block(sd, p, x5):
    """openaimodel3d.py:242-282: 4 x [GN32 -> SiLU -> Conv3d (3,1,1)], identity residual."""
    y = x5
    for i, ci in ((1, 2), (2, 3), (3, 3), (4, 3)):
        q = f"{p}conv{i}."
        y = F.silu(F.group_norm(y, 32, sd[q + "0.weight"], sd[q + "0.bias"], 1e-5))
        y = F.conv3d(y, sd[f"{q}{ci}.weight"], sd[f"{q}{ci}.bias"], padding=(1, 0, 0))
    return x5 + y


def res_block(sd, p, x, emb, b, temporal_conv):
    """openaimodel3d.py:213-239 (use_scale_shift_norm=False, no up/down variants)."""
    h = F.conv2d(F.silu(_gn(sd, p + "in_layers.0.", x, 1e-5)),
                 sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = _lin(sd, p + "emb_layers.1.", F.silu(emb))
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + "out_layers.0.", h, 1e-5)),
                 sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1)
    if p + "skip_connection.weight" in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    h = x + h
    if temporal_conv:
        n, c, hh, ww = h.shape
        h5 = h.reshape(b, n // b, c, hh, ww).permute(0, 2, 1, 3, 4)
        h5 = temporal_conv_block(sd, p + "temopral_conv.", h5)  # sic: reference attribute name
        h = h5.permute(0, 2, 1, 3, 4).reshape(n, c, hh, ww)
    return h


# ---------------------------------------------------------------------------------------- U-Net
@torch.no_grad()
def unet_forward(sd, cfg, x, timesteps, context, fs=None):
    """UNetModel.forward (openaimodel3d.py:552-607). x: b c t h w -> b out_c t h w."""
    b, _, t, hh, ww = x.shape
    mc = cfg["model_channels"]
    use_linear = cfg.get("use_linear", False)
    image_cross = cfg.get("image_cross_attention", False)
    tconv = cfg.get("temporal_conv", False)
    dt = x.dtype

    def mlp(p, v):
        return _lin(sd, p + "2.", F.silu(_lin(sd, p + "0.", v)))

    emb = mlp("time_embed.", timestep_embedding(timesteps, mc).to(dt))
    if context.shape[1] == TEXT_LEN + 16 * t:
        text = context[:, :TEXT_LEN].repeat_interleave(t, dim=0)
        img = context[:, TEXT_LEN:].reshape(b * t, 16, context.shape[-1])
        ctx = torch.cat([text, img], dim=1)
    else:
        ctx = context.repeat_interleave(t, dim=0)
    emb = emb.repeat_interleave(t, dim=0)
    if cfg.get("fs_condition", False):
        if fs is None:
            fs = torch.full((b,), cfg["default_fs"], dtype=torch.long, device=x.device)
        emb = emb + mlp("fps_embedding.", timestep_embedding(fs, mc).to(dt)).repeat_interleave(t, dim=0)

    h = x.permute(0, 2, 1, 3, 4).reshape(b * t, x.shape[1], hh, ww)
    lay = unet_layout(cfg)

    def run(layers, h):
        for L in layers:
            kind, p = L[0], L[1]
            if kind == "conv":
                h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
            elif kind == "res":
                h = res_block(sd, p, h, emb, b, tconv)
            elif kind == "st":
                h = spatial_transformer(sd, p, h, ctx, L[3], use_linear, image_cross)
            elif kind == "tt":
                h = temporal_transformer(sd, p, h, b, L[3], L[4])
            elif kind == "down":
                h = F.conv2d(h, sd[p + "op.weight"], sd[p + "op.bias"], stride=2, padding=1)
            elif kind == "up":
                h = F.interpolate(h.float(), scale_factor=2, mode="nearest").to(h.dtype)
                h = F.conv2d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
        return h

    hs = []
    for i, layers in enumerate(lay["input"]):
        h = run(layers, h)
        if i == 0 and cfg.get("addition_attention", False):
            h = temporal_transformer(sd, "init_attn.0.", h, b, 8, False)
        hs.append(h)
    h = run(lay["middle"], h)
    for layers in lay["output"]:
        h = run(layers, torch.cat([h, hs.pop()], dim=1))
    y = F.conv2d(F.silu(_gn(sd, "out.0.", h, 1e-5)), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    return y.reshape(b, t, -1, hh, ww).permute(0, 2, 1, 3, 4).contiguous()


def unet_param_shapes(cfg):
    """name -> shape of every state-dict entry of the reference UNetModel for `cfg`.

    Restates the constructor's parameter inventory so weights can be synthesised without
    instantiating the reference (which is absent on the GPU box)."""
    mc, cd = cfg["model_channels"], cfg["context_dim"]
    ted = 4 * mc
    use_linear = cfg.get("use_linear", False)
    S = {}

    def lin(p, i, o, bias=True):
        S[p + "weight"] = (o, i)
        if bias:
            S[p + "bias"] = (o,)

    def norm(p, c):
        S[p + "weight"] = (c,)
        S[p + "bias"] = (c,)

    def conv(p, i, o, *k):
        S[p + "weight"] = (o, i, *k)
        S[p + "bias"] = (o,)

    def tblock(p, dim, ctx, image_cross):
        for a, cdim, ip in (("attn1.", dim, False), ("attn2.", ctx or dim, image_cross)):
            lin(p + a + "to_q.", dim, dim, False)
            lin(p + a + "to_k.", cdim, dim, False)
            lin(p + a + "to_v.", cdim, dim, False)
            lin(p + a + "to_out.0.", dim, dim)
            if ip:
                lin(p + a + "to_k_ip.", cdim, dim, False)
                lin(p + a + "to_v_ip.", cdim, dim, False)
        lin(p + "ff.net.0.proj.", dim, dim * 8)
        lin(p + "ff.net.2.", dim * 4, dim)
        for n in ("norm1.", "norm2.", "norm3."):
            norm(p + n, dim)

    def st(p, ch):
        norm(p + "norm.", ch)
        if use_linear:
            lin(p + "proj_in.", ch, ch)
            lin(p + "proj_out.", ch, ch)
        else:
            conv(p + "proj_in.", ch, ch, 1, 1)
            conv(p + "proj_out.", ch, ch, 1, 1)
        tblock(p + "transformer_blocks.0.", ch, cd, cfg.get("image_cross_attention", False))

    def tt(p, ch, inner, linear):
        norm(p + "norm.", ch)
        if linear:
            lin(p + "proj_in.", ch, inner)
            lin(p + "proj_out.", inner, ch)
        else:
            conv(p + "proj_in.", ch, inner, 1)
            conv(p + "proj_out.", inner, ch, 1)
        tblock(p + "transformer_blocks.0.", inner, None, False)

    def res(p, cin, cout):
        norm(p + "in_layers.0.", cin)
        conv(p + "in_layers.2.", cin, cout, 3, 3)
        lin(p + "emb_layers.1.", ted, cout)
        norm(p + "out_layers.0.", cout)
        conv(p + "out_layers.3.", cout, cout, 3, 3)
        if cin != cout:
            conv(p + "skip_connection.", cin, cout, 1, 1)
        if cfg.get("temporal_conv", False):
            for i, ci in ((1, 2), (2, 3), (3, 3), (4, 3)):
                norm(f"{p}temopral_conv.conv{i}.0.", cout)
                conv(f"{p}temopral_conv.conv{i}.{ci}.", cout, cout, 3, 1, 1)

    lin("time_embed.0.", mc, ted)
    lin("time_embed.2.", ted, ted)
    if cfg.get("fs_condition", False):
        lin("fps_embedding.0.", mc, ted)
        lin("fps_embedding.2.", ted, ted)
    lay = unet_layout(cfg)
    first = True
    for group in lay["input"] + [lay["middle"]] + lay["output"]:
        for L in group:
            kind, p = L[0], L[1]
            if kind == "conv":
                conv(p, L[2], L[3], 3, 3)
            elif kind == "res":
                res(p, L[2], L[3])
            elif kind == "st":
                st(p, L[2])
            elif kind == "tt":
                tt(p, L[2], L[2], L[4])
            elif kind == "down":
                conv(p + "op.", L[2], L[2], 3, 3)
            elif kind == "up":
                conv(p + "conv.", L[2], L[2], 3, 3)
        if first and cfg.get("addition_attention", False):
            tt("init_attn.0.", mc, 8 * cfg["num_head_channels"], False)
        first = False
    norm("out.0.", mc)
    conv("out.2.", mc, cfg["out_channels"], 3, 3)
    return S
