"""GPU: every C-ABI kernel against the matching plain-PyTorch fp32/fp64 op on seeded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def abi():
    from pandora_b200 import abi as a
    a.load()
    return a


def rnd(*shape, seed=0, dtype=torch.float32, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


DT = [torch.float32, torch.bfloat16]


def tol(dtype):
    return 2e-5 if dtype == torch.float32 else 6e-3


# ----------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("M,N,K", [(1, 1280, 320), (130, 64, 72), (257, 320, 1024), (64, 2560, 320)])
def test_gemm_linear_bias_residual(abi, dtype, M, N, K):
    a, w = rnd(M, K, seed=1, dtype=dtype), rnd(N, K, seed=2, dtype=dtype, scale=K ** -0.5)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4)
    for cdt in DT:
        c = torch.empty(M, N, dtype=cdt, device="cuda")
        abi.gemm(a, w, c, bias=bias, residual=res)
        ref = a.double() @ w.double().t() + bias.double() + res.double()
        assert rel_l2(c, ref) < (2e-6 if cdt == torch.float32 else 4e-3)


def test_gemm_strided_views_and_rowbias(abi):
    M, N, K = 96, 128, 64
    big_a, big_c = rnd(M, K + 64, seed=5), torch.zeros(M, N + 32, device="cuda")
    a, c = big_a[:, 32:32 + K], big_c[:, 16:16 + N]
    w = rnd(N, K, seed=6, scale=K ** -0.5)
    rb = rnd(3, N + 8, seed=7)[:, 4:4 + N]
    abi.gemm(a, w, c, rowbias=rb, rowbias_rows=32)
    ref = a.double() @ w.double().t() + rb.double().repeat_interleave(32, 0)
    assert rel_l2(c, ref) < 2e-6
    assert float(big_c[:, :16].abs().max()) == 0 and float(big_c[:, 16 + N:].abs().max()) == 0


def conv_weight_rows(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n,cin,cout,H,W,stride,up", [
    (3, 8, 64, 8, 8, 1, False), (2, 64, 128, 10, 16, 1, False), (2, 64, 64, 8, 12, 2, False),
    (2, 64, 64, 5, 8, 1, True), (1, 128, 4, 6, 6, 1, False), (2, 64, 64, 5, 7, 2, False)])
def test_gemm_conv3x3(abi, dtype, n, cin, cout, H, W, stride, up):
    x = rnd(n, cin, H, W, seed=8, dtype=dtype)
    w = rnd(cout, cin, 3, 3, seed=9, dtype=dtype, scale=(9 * cin) ** -0.5)
    bias = rnd(cout, seed=10)
    xin = F.interpolate(x.double(), scale_factor=2, mode="nearest") if up else x.double()
    ref = F.conv2d(xin, w.double(), bias.double(), stride=stride, padding=1)
    oh, ow = ref.shape[-2:]
    rows = x.permute(0, 2, 3, 1).reshape(n * H * W, cin).contiguous()
    res = rnd(n * oh * ow, cout, seed=11)
    c = torch.empty(n * oh * ow, cout, device="cuda")
    abi.gemm(rows, conv_weight_rows(w), c, bias=bias, residual=res, mode=abi.A_CONV3X3,
             conv=(cin, H, W, oh, ow, stride, int(up)))
    ref_rows = ref.permute(0, 2, 3, 1).reshape(n * oh * ow, cout) + res.double()
    assert rel_l2(c, ref_rows) < 2e-6


@pytest.mark.parametrize("dtype", DT)
def test_gemm_convt3(abi, dtype):
    b, T, HW, C = 2, 16, 12, 64
    x = rnd(b, C, T, HW, 1, seed=12, dtype=dtype)
    w = rnd(C, C, 3, 1, 1, seed=13, dtype=dtype, scale=(3 * C) ** -0.5)
    bias = rnd(C, seed=14)
    ref = F.conv3d(x.double(), w.double(), bias.double(), padding=(1, 0, 0))
    rows = x[..., 0].permute(0, 2, 3, 1).reshape(b * T * HW, C).contiguous()
    wr = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    c = torch.empty(b * T * HW, C, device="cuda")
    abi.gemm(rows, wr, c, bias=bias, mode=abi.A_CONVT3, convt=(C, T, HW))
    ref_rows = ref[..., 0].permute(0, 2, 3, 1).reshape(b * T * HW, C)
    assert rel_l2(c, ref_rows) < 2e-6


def test_gemm_rejects_bad_arguments(abi):
    a, w, c = rnd(8, 6), rnd(8, 6), torch.empty(8, 8, device="cuda")
    with pytest.raises(RuntimeError, match="multiples of 4"):
        abi.gemm(a, w, c)


# ---------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("xdt", DT)
@pytest.mark.parametrize("n_inst,R,C,silu,eps", [
    (16, 64, 64, True, 1e-5), (2, 16 * 40, 320, False, 1e-6), (3, 50, 1280, True, 1e-5),
    (1, 7, 2560, True, 1e-5), (5, 33, 960, True, 1e-5), (2, 128, 128, False, 1e-6)])
def test_groupnorm(abi, xdt, n_inst, R, C, silu, eps):
    x = (rnd(n_inst * R, C, seed=20) * 2 + 0.7).to(xdt)
    gamma, beta = 1 + 0.1 * rnd(C, seed=21), 0.1 * rnd(C, seed=22)
    xr = x.double().reshape(n_inst, R, C).permute(0, 2, 1)
    ref = F.group_norm(xr, 32, gamma.double(), beta.double(), eps)
    ref = (F.silu(ref) if silu else ref).permute(0, 2, 1).reshape(n_inst * R, C)
    for ydt in DT:
        y = torch.empty(n_inst * R, C, dtype=ydt, device="cuda")
        abi.groupnorm(x, y, n_inst, R, gamma, beta, eps, silu)
        assert rel_l2(y, ref) < tol(ydt)


@pytest.mark.parametrize("n_inst,R,C", [(2, 640, 1280), (32, 40, 1280), (16, 160, 640), (1, 333, 128)])
def test_groupnorm_single_launch_small_instances(abi, n_inst, R, C):
    """Small tensors take the one-launch kernel (a CTA per (instance, group)); same result as the three-pass path."""
    x = rnd(n_inst * R, C, seed=23) * 3 - 0.4
    gamma, beta = 1 + 0.1 * rnd(C, seed=24), 0.1 * rnd(C, seed=25)
    ref = F.group_norm(x.double().reshape(n_inst, R, C).permute(0, 2, 1), 32, gamma.double(), beta.double(), 1e-5)
    ref = F.silu(ref).permute(0, 2, 1).reshape(n_inst * R, C)
    l0 = abi.launches
    y = torch.empty(n_inst * R, C, device="cuda")
    abi.groupnorm(x, y, n_inst, R, gamma, beta, 1e-5, True)
    assert abi.launches - l0 == (1 if n_inst * 32 >= 32 else 3)
    assert rel_l2(y, ref) < 2e-6
    yb = torch.empty(n_inst * R, C, device="cuda", dtype=torch.bfloat16)
    abi.groupnorm(x, yb, n_inst, R, gamma, beta, 1e-5, True)
    assert rel_l2(yb, ref) < 4e-3


def test_groupnorm_leaves_operand_copy_of_its_input(abi):
    """x_copy: the apply pass also writes x itself in the output dtype (the 1x1 skip conv's operand)."""
    n_inst, R, C = 4, 2560, 320
    x = rnd(n_inst * R, C, seed=26) * 2
    gamma, beta = 1 + 0.1 * rnd(C, seed=27), 0.1 * rnd(C, seed=28)
    y = torch.empty(n_inst * R, C, device="cuda", dtype=torch.bfloat16)
    wide = torch.zeros(n_inst * R, C + 64, device="cuda", dtype=torch.bfloat16)
    abi.groupnorm(x, y, n_inst, R, gamma, beta, 1e-5, True, x_copy=wide[:, 64:])
    assert torch.equal(wide[:, 64:], x.to(torch.bfloat16)) and bool((wide[:, :64] == 0).all())
    y2 = torch.empty_like(y)
    abi.groupnorm(x, y2, n_inst, R, gamma, beta, 1e-5, True)
    assert torch.equal(y, y2)


@pytest.mark.parametrize("xdt", DT)
@pytest.mark.parametrize("rows,C", [(77, 64), (1000, 320), (33, 512), (64, 640), (9, 1280)])
def test_layernorm(abi, xdt, rows, C):
    x = (rnd(rows, C, seed=23) * 3 - 1).to(xdt)
    gamma, beta = 1 + 0.1 * rnd(C, seed=24), 0.1 * rnd(C, seed=25)
    ref = F.layer_norm(x.double(), (C,), gamma.double(), beta.double(), 1e-5)
    for ydt in DT:
        y = torch.empty(rows, C, dtype=ydt, device="cuda")
        abi.layernorm(x, y, gamma, beta)
        assert rel_l2(y, ref) < tol(ydt)


# ------------------------------------------------------------------------------------ attention
def ref_attn(q, k, v, heads):
    B, n, C = q.shape
    sp = lambda t: t.double().reshape(t.shape[0], t.shape[1], heads, 64).permute(0, 2, 1, 3)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.permute(0, 2, 1, 3).reshape(B, n, C)


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("B,heads,nq,nk", [(3, 2, 160, 160), (2, 5, 100, 77), (4, 1, 40, 16), (1, 3, 300, 300)])
def test_attention_frame_batched(abi, dtype, B, heads, nq, nk):
    C = heads * 64
    qkv = rnd(B * nq, 3 * C, seed=30, dtype=dtype)
    kv = rnd(B * nk, 2 * C, seed=31, dtype=dtype)
    q = qkv[:, :C].reshape(B, nq, C)
    if nq == nk:
        k, v = qkv[:, C:2 * C].reshape(B, nk, C), qkv[:, 2 * C:].reshape(B, nk, C)
        kop = abi.operand(qkv, 1, nk * 3 * C, 0, 3 * C, C)
        vop = abi.operand(qkv, 1, nk * 3 * C, 0, 3 * C, 2 * C)
    else:
        k, v = kv[:, :C].reshape(B, nk, C), kv[:, C:].reshape(B, nk, C)
        kop = abi.operand(kv, 1, nk * 2 * C, 0, 2 * C, 0)
        vop = abi.operand(kv, 1, nk * 2 * C, 0, 2 * C, C)
    out = torch.zeros(B * nq, C, dtype=dtype, device="cuda")
    code = abi.BF16 if dtype == torch.bfloat16 else abi.F32
    abi.attention(abi.operand(qkv, 1, nq * 3 * C, 0, 3 * C, 0), kop, vop, out, (1, nq * C, 0, C),
                  B, heads, nq, nk, 0.125, code)
    ref = ref_attn(q, k, v, heads).reshape(B * nq, C)
    assert rel_l2(out, ref) < tol(dtype)
    # accumulate: out += attention (text + image cross-attention sum, attention.py:142)
    abi.attention(abi.operand(qkv, 1, nq * 3 * C, 0, 3 * C, 0), kop, vop, out, (1, nq * C, 0, C),
                  B, heads, nq, nk, 0.125, code, accumulate=True)
    assert rel_l2(out, 2 * ref) < 2 * tol(dtype)


@pytest.mark.parametrize("dtype", DT)
def test_attention_text_kv_shared_across_frames(abi, dtype):
    b, T, heads, nq, nk = 2, 4, 2, 48, 77
    C = heads * 64
    q = rnd(b * T * nq, C, seed=32, dtype=dtype)
    kv = rnd(b * nk, 2 * C, seed=33, dtype=dtype)
    out = torch.empty(b * T * nq, C, dtype=dtype, device="cuda")
    code = abi.BF16 if dtype == torch.bfloat16 else abi.F32
    abi.attention(abi.operand(q, 1, nq * C, 0, C, 0), abi.operand(kv, T, nk * 2 * C, 0, 2 * C, 0),
                  abi.operand(kv, T, nk * 2 * C, 0, 2 * C, C), out, (1, nq * C, 0, C), b * T, heads, nq, nk, 0.125, code)
    k = kv[:, :C].reshape(b, nk, C).repeat_interleave(T, 0)
    v = kv[:, C:].reshape(b, nk, C).repeat_interleave(T, 0)
    assert rel_l2(out, ref_attn(q.reshape(b * T, nq, C), k, v, heads).reshape(-1, C)) < tol(dtype)


@pytest.mark.parametrize("dtype", DT)
def test_attention_temporal_in_place_layout(abi, dtype):
    """rows ordered (b, t, p); attention over t for every pixel p without any transpose."""
    b, T, HW, heads = 2, 16, 24, 2
    C = heads * 64
    qkv = rnd(b * T * HW, 3 * C, seed=34, dtype=dtype)
    out = torch.empty(b * T * HW, C, dtype=dtype, device="cuda")
    ld = 3 * C
    op = lambda off: abi.operand(qkv, HW, T * HW * ld, ld, HW * ld, off)
    code = abi.BF16 if dtype == torch.bfloat16 else abi.F32
    abi.attention(op(0), op(C), op(2 * C), out, (HW, T * HW * C, C, HW * C), b * HW, heads, T, T, 0.125, code)
    x = qkv.reshape(b, T, HW, 3 * C).permute(0, 2, 1, 3).reshape(b * HW, T, 3 * C)
    ref = ref_attn(x[..., :C], x[..., C:2 * C], x[..., 2 * C:], heads)
    ref = ref.reshape(b, HW, T, C).permute(0, 2, 1, 3).reshape(b * T * HW, C)
    assert rel_l2(out, ref) < tol(dtype)


@pytest.mark.parametrize("T,HW,heads", [(16, 40, 5), (8, 24, 2), (5, 17, 1), (16, 2560, 5)])
def test_attention_temporal_bf16_short_and_ragged(abi, T, HW, heads):
    """mma.sync temporal kernel: sequences shorter than 16 (key masking, rows not stored) and odd pixel counts."""
    b, C = 1, heads * 64
    qkv = rnd(b * T * HW, 3 * C, seed=35, dtype=torch.bfloat16)
    out = torch.full((b * T * HW, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    ld = 3 * C
    op = lambda off: abi.operand(qkv, HW, T * HW * ld, ld, HW * ld, off)
    abi.attention(op(0), op(C), op(2 * C), out, (HW, T * HW * C, C, HW * C), b * HW, heads, T, T, 0.125, abi.BF16)
    x = qkv.reshape(b, T, HW, 3 * C).permute(0, 2, 1, 3).reshape(b * HW, T, 3 * C)
    ref = ref_attn(x[..., :C], x[..., C:2 * C], x[..., 2 * C:], heads)
    ref = ref.reshape(b, HW, T, C).permute(0, 2, 1, 3).reshape(b * T * HW, C)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < tol(torch.bfloat16)


# ----------------------------------------------------------------------------------- small ops
def test_geglu_silu_add_copy(abi):
    u = rnd(37, 256, seed=40)
    y = torch.empty(37, 128, device="cuda")
    abi.geglu(u, y)
    assert rel_l2(y, u[:, :128].double() * F.gelu(u[:, 128:].double())) < 1e-6
    yb = torch.empty(37, 128, dtype=torch.bfloat16, device="cuda")
    abi.geglu(u.bfloat16(), yb)
    assert rel_l2(yb, u.bfloat16()[:, :128].double() * F.gelu(u.bfloat16()[:, 128:].double())) < 6e-3
    x = rnd(1000, seed=41)
    s = torch.empty_like(x)
    abi.silu(x, s)
    assert rel_l2(s, F.silu(x.double())) < 1e-6
    abi.add(x, s, s)
    assert rel_l2(s, x.double() + F.silu(x.double())) < 1e-6
    dst = torch.zeros(37, 400, dtype=torch.bfloat16, device="cuda")
    abi.copy_rows(u, dst[:, 100:356])
    assert torch.equal(dst[:, 100:356], u.bfloat16()) and float(dst[:, :100].abs().max()) == 0


def test_upsample2x_rows(abi):
    n, H, W, C = 3, 5, 6, 64
    x = rnd(n, C, H, W, seed=44)
    rows = x.permute(0, 2, 3, 1).reshape(n * H * W, C).contiguous()
    out = torch.empty(n * 4 * H * W, C, dtype=torch.bfloat16, device="cuda")
    abi.upsample2x_rows(rows, out, n, H, W)
    ref = F.interpolate(x, scale_factor=2, mode="nearest").permute(0, 2, 3, 1).reshape(-1, C).bfloat16()
    assert torch.equal(out, ref)


def test_timestep_embedding_bf16_table(abi):
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.bfloat16) / half).float().cuda()
    t = torch.tensor([999, 19, 0, 500], dtype=torch.long, device="cuda")
    out = torch.empty(4, 2 * half, device="cuda")
    abi.timestep_embedding(t, freqs, out)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert float((out - ref).abs().max()) < 2e-6


@pytest.mark.parametrize("dtype", DT)
def test_video_rows_roundtrip(abi, dtype):
    b, c1, c2, T, H, W = 2, 4, 4, 3, 5, 6
    x1, x2 = rnd(b, c1, T, H, W, seed=42, dtype=dtype), rnd(b, c2, T, H, W, seed=43, dtype=dtype)
    rows = torch.empty(b * T * H * W, c1 + c2, device="cuda")
    abi.video_to_rows(x1, x2, rows, b, T, H * W)
    ref = torch.cat([x1, x2], 1).float().permute(0, 2, 3, 4, 1).reshape(-1, c1 + c2)
    assert torch.equal(rows, ref)
    back = torch.empty(b, c1 + c2, T, H, W, dtype=dtype, device="cuda")
    abi.rows_to_video(rows, back, b, c1 + c2, T, H * W)
    assert torch.equal(back, torch.cat([x1, x2], 1))


@pytest.mark.parametrize("phi", [0.0, 0.7])
@pytest.mark.parametrize("cfg", [True, False])
def test_ddim_cfg_step_bit_exact_vs_torch_fp32(abi, phi, cfg):
    b, n = 2, 4 * 16 * 8 * 8
    x, vc, vu, nz = (rnd(b, n, seed=s) for s in (50, 51, 52, 53))
    coef = torch.tensor([[0.9, 0.43, 0.8, 0.31, 1.0156, 0.52, 0, 0], [0.5, 0.86, 0.95, 0.2, 1.0, 0.22, 0, 0]]).cuda()
    idx = torch.tensor([1], dtype=torch.int32, device="cuda")
    xp, x0 = torch.empty_like(x), torch.empty_like(x)
    ws = torch.empty(4 * b, dtype=torch.float64, device="cuda")
    abi.ddim_cfg_step(x, vc, vu if cfg else None, nz, xp, x0, 7.5, phi, 1.0, coef, idx, True, ws)
    assert int(idx[0]) == 2
    sa, so, sap, sig, resc, dirc = (coef[1, i] for i in range(6))
    v = vc
    if cfg:
        v = vu + 7.5 * (vc - vu)
        if phi > 0:
            std_c = vc.std(dim=1, keepdim=True)
            std_g = v.std(dim=1, keepdim=True)
            v = phi * (v * (std_c / std_g)) + (1 - phi) * v
    e = sa * v + so * x
    r0 = (sa * x - so * v) * resc
    rp = sap * r0 + dirc * e + sig * nz * 1.0
    if phi > 0 and cfg:
        assert rel_l2(xp, rp) < 1e-6 and rel_l2(x0, r0) < 1e-6
    else:
        assert torch.equal(x0, r0) and torch.equal(xp, rp)


@pytest.mark.parametrize("M,N,K", [(1, 1280, 320), (2, 20160, 1280), (8, 1000, 1284), (3, 36, 64)])
def test_gemm_few_rows_embedding_path(abi, M, N, K):
    """The timestep / fps embedding MLPs and the folded ResBlock emb projections (1-2 rows): warp-per-column GEMV."""
    a, w, bias = rnd(M, K, seed=70), rnd(N, K, seed=71, scale=K ** -0.5), rnd(N, seed=72)
    big = torch.zeros(M, N + 8, device="cuda")
    c = big[:, 4:4 + N]
    assert abi.gemm(a, w, c, bias=bias)
    assert rel_l2(c, a.double() @ w.double().t() + bias.double()) < 2e-6
    assert float(big[:, :4].abs().max()) == 0 and float(big[:, 4 + N:].abs().max()) == 0
    c2 = torch.empty(M, N, device="cuda")
    assert abi.gemm(a, w, c2)
    assert rel_l2(c2, a.double() @ w.double().t()) < 2e-6
