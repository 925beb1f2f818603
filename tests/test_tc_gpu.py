"""GPU: tcgen05/TMA kernels against fp64 PyTorch on the same bf16 operands (bit-level inputs shared)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from test_kernels_gpu import conv_weight_rows, ref_attn, rnd

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module")
def abi():
    from pandora_b200 import abi as a
    a.load()
    if not a.supports_tc():
        pytest.skip("device has no tcgen05")
    return a


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 320, 320), (2560, 1280, 1280), (1232, 640, 1024),
                                   (4096, 960, 320), (640, 256, 2560), (77, 192, 128)])
@pytest.mark.parametrize("cdt", [torch.float32, BF])
def test_tc_gemm_linear(abi, M, N, K, cdt):
    a, w = rnd(M, K, seed=1, dtype=BF), rnd(N, K, seed=2, dtype=BF, scale=K ** -0.5)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
    c = torch.full((M, N), float("nan"), dtype=cdt, device="cuda")
    assert abi.gemm(a, w, c, bias=bias, residual=res, tc=True)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    assert rel_l2(c, ref) < (3e-6 if cdt == torch.float32 else 4e-3)


def test_tc_gemm_plain_and_rowbias_and_views(abi):
    M, N, K = 512, 320, 192
    big_a = rnd(M, K + 64, seed=5, dtype=BF)
    a = big_a[:, 64:]
    w = rnd(N, K, seed=6, dtype=BF, scale=K ** -0.5)
    big_c = torch.zeros(M, N + 64, device="cuda")
    c = big_c[:, 32:32 + N]
    rb = rnd(4, N, seed=7)
    assert abi.gemm(a, w, c, rowbias=rb, rowbias_rows=128, tc=True)
    ref = a.double() @ w.double().t() + rb.double().repeat_interleave(128, 0)
    assert rel_l2(c, ref) < 3e-6
    assert float(big_c[:, :32].abs().max()) == 0 and float(big_c[:, 32 + N:].abs().max()) == 0


@pytest.mark.parametrize("M,C", [(256, 64), (1000, 320), (2560, 1280)])
def test_tc_gemm_geglu(abi, M, C):
    a = rnd(M, C, seed=8, dtype=BF)
    w = rnd(8 * C, C, seed=9, dtype=BF, scale=C ** -0.5)
    bias = rnd(8 * C, seed=10)
    y = torch.empty(M, 4 * C, dtype=BF, device="cuda")
    assert abi.gemm(a, w, y, bias=bias, geglu=True, tc=True)
    u = a.double() @ w.double().t() + bias.double()
    ref = u[:, :4 * C] * F.gelu(u[:, 4 * C:])
    assert rel_l2(y, ref) < 4e-3


@pytest.mark.parametrize("n,cin,cout,H,W", [(16, 64, 64, 8, 8), (16, 128, 64, 4, 4), (2, 320, 320, 40, 64),
                                            (16, 640, 320, 10, 16), (16, 1280, 1280, 5, 8), (4, 64, 128, 20, 32),
                                            (2, 64, 64, 18, 32), (1, 64, 64, 72, 128)])
def test_tc_conv3x3(abi, n, cin, cout, H, W):
    x = rnd(n, cin, H, W, seed=11, dtype=BF)
    w = rnd(cout, cin, 3, 3, seed=12, dtype=BF, scale=(9 * cin) ** -0.5)
    bias = rnd(cout, seed=13)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    rows = x.permute(0, 2, 3, 1).reshape(n * H * W, cin).contiguous()
    emb = rnd(n // 2 if n > 1 else 1, cout, seed=14)
    rb_rows = 2 * H * W if n > 1 else H * W
    res = rnd(n * H * W, cout, seed=15)
    c = torch.full((n * H * W, cout), float("nan"), device="cuda")
    ok = abi.gemm(rows, conv_weight_rows(w), c, bias=bias, residual=res, rowbias=emb, rowbias_rows=rb_rows,
                  mode=abi.A_CONV3X3, conv=(cin, H, W, H, W, 1, 0), tc=True)
    assert ok, "expected the TMA path to take this shape"
    ref_rows = ref.permute(0, 2, 3, 1).reshape(n * H * W, cout) + res.double() \
        + emb.double().repeat_interleave(rb_rows, 0)
    assert rel_l2(c, ref_rows) < 1e-5      # K up to 23040 accumulated in the tensor core's fp32


def test_tc_conv3x3_declines_what_it_cannot_tile(abi):
    rows = rnd(16 * 2 * 2, 64, seed=16, dtype=BF)
    w = rnd(64, 9 * 64, seed=17, dtype=BF)
    c = torch.empty(64, 64, device="cuda")
    assert abi.gemm(rows, w, c, mode=abi.A_CONV3X3, conv=(64, 2, 2, 2, 2, 1, 0), tc=True) == 0
    assert abi.gemm(rows, w, c, mode=abi.A_CONV3X3, conv=(64, 2, 2, 4, 4, 1, 1), tc=True, M=256) == 0


@pytest.mark.parametrize("b,T,HW,C", [(1, 16, 160, 320), (2, 16, 64, 64), (1, 16, 40, 1280), (1, 16, 2560, 64),
                                      (1, 16, 16, 128)])
def test_tc_convt3(abi, b, T, HW, C):
    x = rnd(b, C, T, HW, 1, seed=18, dtype=BF)
    w = rnd(C, C, 3, 1, 1, seed=19, dtype=BF, scale=(3 * C) ** -0.5)
    bias = rnd(C, seed=20)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=(1, 0, 0))
    rows = x[..., 0].permute(0, 2, 3, 1).reshape(b * T * HW, C).contiguous()
    wr = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    res = rnd(b * T * HW, C, seed=21)
    c = torch.full((b * T * HW, C), float("nan"), device="cuda")
    assert abi.gemm(rows, wr, c, bias=bias, residual=res, mode=abi.A_CONVT3, convt=(C, T, HW), tc=True)
    ref_rows = ref[..., 0].permute(0, 2, 3, 1).reshape(b * T * HW, C) + res.double()
    assert rel_l2(c, ref_rows) < 3e-6


def test_tc_gemm_back_to_back_launches_are_deterministic(abi):
    a, w = rnd(5000, 640, seed=22, dtype=BF), rnd(640, 640, seed=23, dtype=BF, scale=640 ** -0.5)
    outs = []
    for _ in range(3):
        c = torch.empty(5000, 640, device="cuda")
        abi.gemm(a, w, c, tc=True)
        outs.append(c)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


# ------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,heads,nq,nk", [(2, 2, 128, 128), (3, 5, 160, 160), (2, 1, 640, 640), (1, 2, 2560, 2560),
                                           (2, 3, 200, 77), (2, 2, 64, 16), (1, 1, 300, 129),
                                           # the spatial token counts of the 576x1024 clip (72x128 ... 9x16 latents)
                                           (1, 1, 9216, 9216), (1, 2, 2304, 2304), (2, 2, 576, 576), (2, 3, 144, 144)])
def test_tc_attention_frame_batched(abi, B, heads, nq, nk):
    C = heads * 64
    qkv = rnd(B * nq, 3 * C, seed=30, dtype=BF)
    kv = rnd(B * nk, 2 * C, seed=31, dtype=BF)
    q = qkv[:, :C].reshape(B, nq, C)
    if nq == nk:
        k, v = qkv[:, C:2 * C].reshape(B, nk, C), qkv[:, 2 * C:].reshape(B, nk, C)
        kop = abi.operand(qkv, 1, nk * 3 * C, 0, 3 * C, C)
        vop = abi.operand(qkv, 1, nk * 3 * C, 0, 3 * C, 2 * C)
    else:
        k, v = kv[:, :C].reshape(B, nk, C), kv[:, C:].reshape(B, nk, C)
        kop = abi.operand(kv, 1, nk * 2 * C, 0, 2 * C, 0)
        vop = abi.operand(kv, 1, nk * 2 * C, 0, 2 * C, C)
    out = torch.full((B * nq, C), float("nan"), dtype=BF, device="cuda")
    qop = abi.operand(qkv, 1, nq * 3 * C, 0, 3 * C, 0)
    assert abi.attention(qop, kop, vop, out, (1, nq * C, 0, C), B, heads, nq, nk, 0.125, abi.BF16, tc=True)
    ref = ref_attn(q, k, v, heads).reshape(B * nq, C)
    e = rel_l2(out, ref)
    assert e < 6e-3, e
    assert abi.attention(qop, kop, vop, out, (1, nq * C, 0, C), B, heads, nq, nk, 0.125, abi.BF16, accumulate=True, tc=True)
    assert rel_l2(out, 2 * ref) < 1e-2


@pytest.mark.parametrize("n,ramp", [(9216, (0.05, 8.0)), (2560, (8.0, 0.05)), (2304, (0.02, 12.0))])
def test_tc_attention_online_softmax_with_drifting_maxima(abi, n, ramp):
    """Key norms ramp over the sequence, so the row maxima drift by tens of log2 units across the key tiles: the
    accumulator in tensor memory is rescaled lazily (only when a maximum grows by more than 2^8) and stale maxima
    are used in between; an un-normalised P reaches 2^8 there.  Both directions, 18 .. 72 key tiles."""
    heads, C = 1, 64
    q = rnd(n, C, seed=41, dtype=BF)
    k = (rnd(n, C, seed=42) * torch.linspace(ramp[0], ramp[1], n, device="cuda")[:, None]).to(BF)
    v = rnd(n, C, seed=43, dtype=BF)
    out = torch.full((n, C), float("nan"), dtype=BF, device="cuda")
    op = lambda t: abi.operand(t, 1, n * C, 0, C, 0)
    assert abi.attention(op(q), op(k), op(v), out, (1, n * C, 0, C), 1, heads, n, n, 0.125, abi.BF16, tc=True)
    ref = ref_attn(q[None], k[None], v[None], heads)[0]
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 8e-3


def test_tc_attention_text_kv_shared_across_frames(abi):
    b, T, heads, nq, nk = 2, 4, 2, 160, 77
    C = heads * 64
    q = rnd(b * T * nq, C, seed=32, dtype=BF)
    kv = rnd(b * nk, 2 * C, seed=33, dtype=BF)
    out = torch.empty(b * T * nq, C, dtype=BF, device="cuda")
    assert abi.attention(abi.operand(q, 1, nq * C, 0, C, 0), abi.operand(kv, T, nk * 2 * C, 0, 2 * C, 0),
                         abi.operand(kv, T, nk * 2 * C, 0, 2 * C, C), out, (1, nq * C, 0, C), b * T, heads, nq, nk,
                         0.125, abi.BF16, tc=True)
    k = kv[:, :C].reshape(b, nk, C).repeat_interleave(T, 0)
    v = kv[:, C:].reshape(b, nk, C).repeat_interleave(T, 0)
    assert rel_l2(out, ref_attn(q.reshape(b * T, nq, C), k, v, heads).reshape(-1, C)) < 6e-3


@pytest.mark.parametrize("nq,heads", [(160, 2), (2560, 5), (100, 1)])
def test_tc_attention_two_segment_cross(abi, nq, heads):
    """kv_split: text (77) and image (16) keys of a frame in one tile, two softmaxes, outputs summed
    (attention.py:128-142) == two separate attentions added."""
    B, nt, ni = 4, 77, 16
    C = heads * 64
    q = rnd(B * nq, C, seed=36, dtype=BF)
    kv = rnd(B * (nt + ni), 2 * C, seed=37, dtype=BF)
    out = torch.full((B * nq, C), float("nan"), dtype=BF, device="cuda")
    n = nt + ni
    assert abi.attention(abi.operand(q, 1, nq * C, 0, C, 0), abi.operand(kv, 1, n * 2 * C, 0, 2 * C, 0),
                         abi.operand(kv, 1, n * 2 * C, 0, 2 * C, C), out, (1, nq * C, 0, C), B, heads, nq, n, 0.125,
                         abi.BF16, tc=True, kv_split=nt)
    k3, v3 = kv[:, :C].reshape(B, n, C), kv[:, C:].reshape(B, n, C)
    q3 = q.reshape(B, nq, C)
    ref = ref_attn(q3, k3[:, :nt], v3[:, :nt], heads) + ref_attn(q3, k3[:, nt:], v3[:, nt:], heads)
    assert rel_l2(out, ref.reshape(B * nq, C)) < 6e-3


def test_tc_attention_declines_temporal_layout(abi):
    qkv = rnd(16 * 24, 3 * 64, seed=34, dtype=BF)
    out = torch.empty(16 * 24, 64, dtype=BF, device="cuda")
    op = lambda off: abi.operand(qkv, 24, 16 * 24 * 192, 192, 24 * 192, off)
    assert abi.attention(op(0), op(64), op(128), out, (24, 16 * 24 * 64, 64, 24 * 64), 24, 1, 16, 16, 0.125,
                         abi.BF16, tc=True) in (False, 0)


@pytest.mark.parametrize("n,cin,cout,H,W", [(16, 64, 64, 8, 16), (2, 320, 320, 40, 64), (16, 640, 640, 20, 32),
                                            (16, 1280, 1280, 10, 16), (4, 64, 64, 36, 64)])
def test_tc_conv3x3_stride2_by_tma_element_strides(abi, n, cin, cout, H, W):
    x = rnd(n, cin, H, W, seed=41, dtype=BF)
    w = rnd(cout, cin, 3, 3, seed=42, dtype=BF, scale=(9 * cin) ** -0.5)
    bias = rnd(cout, seed=43)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=2, padding=1)
    oh, ow = ref.shape[-2:]
    rows = x.permute(0, 2, 3, 1).reshape(n * H * W, cin).contiguous()
    c = torch.full((n * oh * ow, cout), float("nan"), device="cuda")
    ok = abi.gemm(rows, conv_weight_rows(w), c, bias=bias, mode=abi.A_CONV3X3, conv=(cin, H, W, oh, ow, 2, 0), tc=True)
    assert ok
    assert rel_l2(c, ref.permute(0, 2, 3, 1).reshape(n * oh * ow, cout)) < 1e-5


@pytest.mark.parametrize("n,cin,cout,H,W", [(4, 64, 128, 9, 16), (8, 128, 64, 9, 16), (3, 64, 64, 18, 32), (2, 64, 64, 36, 64),
                                            (5, 64, 64, 10, 16)])
def test_tc_conv3x3_ragged_bricks(abi, n, cin, cout, H, W):
    """Images that no exact 128-row brick tiles (576x1024 latents sharded over 4 ranks: 9x16 x 4 frames): edge tiles
    hang over the tensor, the copy engine zero-fills / clips them."""
    x = rnd(n, cin, H, W, seed=51, dtype=BF)
    w = rnd(cout, cin, 3, 3, seed=52, dtype=BF, scale=(9 * cin) ** -0.5)
    bias, res = rnd(cout, seed=53), rnd(n * H * W, cout, seed=54)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    rows = x.permute(0, 2, 3, 1).reshape(n * H * W, cin).contiguous()
    wide = torch.full((n * H * W + 64, cout), 7.0, device="cuda")           # guard rows after the tensor
    c = wide[:n * H * W]
    c.fill_(float("nan"))
    assert abi.gemm(rows, conv_weight_rows(w), c, bias=bias, residual=res, mode=abi.A_CONV3X3,
                    conv=(cin, H, W, H, W, 1, 0), tc=True), "expected the TMA path to take this shape"
    assert rel_l2(c, ref.permute(0, 2, 3, 1).reshape(n * H * W, cout) + res.double()) < 1e-5
    assert float((wide[n * H * W:] - 7.0).abs().max()) == 0


@pytest.mark.parametrize("nb,C,HW,T", [(1, 64, 36, 16), (2, 128, 36, 16), (1, 64, 144, 16), (1, 64, 100, 8)])
def test_tc_convt3_ragged_bricks(abi, nb, C, HW, T):
    x = rnd(nb, C, T, HW, seed=55, dtype=BF)
    w = rnd(C, C, 3, seed=56, dtype=BF, scale=(3 * C) ** -0.5)
    bias = rnd(C, seed=57)
    ref = F.conv1d(x.double().permute(0, 3, 1, 2).reshape(nb * HW, C, T), w.double(), bias.double(), padding=1)
    ref = ref.reshape(nb, HW, C, T).permute(0, 3, 1, 2).reshape(nb * T * HW, C)
    rows = x.permute(0, 2, 3, 1).reshape(nb * T * HW, C).contiguous()
    wr = w.permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    c = torch.full((nb * T * HW, C), float("nan"), device="cuda")
    assert abi.gemm(rows, wr, c, bias=bias, mode=abi.A_CONVT3, convt=(C, T, HW), tc=True)
    assert rel_l2(c, ref) < 1e-5


def _slab_stats(c):
    """fp64 reference of pb_gemm_args.colstats: per 32-row slab, per column (sum, sumsq)."""
    M, N = c.shape
    v = c.double().reshape(M // 32, 32, N)
    return torch.stack([v.sum(1), (v * v).sum(1)], dim=-1).reshape(M // 32, 2 * N)


@pytest.mark.parametrize("mode,shape", [("linear", (4096, 320, 320)), ("linear", (640, 1280, 320)),
                                        ("conv", (4, 64, 320, 40, 64)), ("conv", (16, 128, 64, 20, 32)),
                                        ("conv", (32, 64, 320, 10, 16)), ("convt", (2, 64, 2560)),
                                        ("convt", (1, 128, 160))])
def test_tc_gemm_colstats_and_fused_groupnorm(abi, mode, shape):
    """The epilogue's column statistics equal a reduction of the stored output, and GroupNorm from them equals
    GroupNorm from a pass over the tensor (also through a column slice of a wider buffer)."""
    if mode == "linear":
        M, N, K = shape
        a, w = rnd(M, K, seed=41, dtype=BF), rnd(N, K, seed=42, dtype=BF, scale=K ** -0.5)
        kw, n_inst = {}, 4
    elif mode == "conv":
        n, cin, N, H, W = shape
        M, K = n * H * W, 9 * cin
        a, w = rnd(M, cin, seed=41, dtype=BF), rnd(N, K, seed=42, dtype=BF, scale=K ** -0.5)
        kw, n_inst = dict(mode=abi.A_CONV3X3, conv=(cin, H, W, H, W, 1, 0)), n
    else:
        nb, C, HW = shape
        M, K, N = nb * 16 * HW, 3 * C, C
        a, w = rnd(M, C, seed=41, dtype=BF), rnd(N, K, seed=42, dtype=BF, scale=K ** -0.5)
        kw, n_inst = dict(mode=abi.A_CONVT3, convt=(C, 16, HW)), nb
    bias, res = rnd(N, seed=43), rnd(M, N, seed=44)
    wide = torch.zeros(M, N + 64, device="cuda")
    wide_cs = torch.full((M // 32, 2 * (N + 64)), float("nan"), device="cuda")
    c, cs = wide[:, 64:], wide_cs[:, 128:]
    rc = abi.gemm(a, w, c, bias=bias, residual=res, tc=True, colstats=cs, **kw)
    assert rc == 2, "expected the tcgen05 path to write column statistics for this shape"
    ref = _slab_stats(c)
    assert rel_l2(cs, ref) < 1e-5
    assert torch.isnan(wide_cs[:, :128]).all()
    # GroupNorm from the fused statistics == GroupNorm with its own statistics pass
    gamma, beta = rnd(N, seed=45), rnd(N, seed=46)
    rows = M // n_inst
    y0 = torch.empty(M, N, dtype=BF, device="cuda")
    y1 = torch.empty(M, N, dtype=BF, device="cuda")
    abi.groupnorm(c, y0, n_inst, rows, gamma, beta, 1e-5, True)
    abi.groupnorm(c, y1, n_inst, rows, gamma, beta, 1e-5, True, colstats=cs)
    assert rel_l2(y1, y0.double()) < 2e-3
    want = F.silu(F.group_norm(c.double().reshape(n_inst, rows, N).permute(0, 2, 1), 32, gamma.double(),
                               beta.double(), 1e-5)).permute(0, 2, 1).reshape(M, N)
    assert rel_l2(y1, want) < 4e-3


def test_tc_gemm_colstats_declined(abi):
    """Shapes whose tiles do not map to aligned 32-row slabs (5x8 images) or that use split-K report 1, not 2."""
    n, cin, N, H, W = 16, 64, 64, 5, 8
    M = n * H * W
    a, w = rnd(M, cin, seed=41, dtype=BF), rnd(N, 9 * cin, seed=42, dtype=BF, scale=0.05)
    c = torch.empty(M, N, device="cuda")
    cs = torch.zeros(M // 32, 2 * N, device="cuda")
    assert abi.gemm(a, w, c, tc=True, colstats=cs, mode=abi.A_CONV3X3, conv=(cin, H, W, H, W, 1, 0)) == 1
