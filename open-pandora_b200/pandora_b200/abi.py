"""ctypes binding of libpandora_b200.so (the C ABI declared in include/pandora_b200.h).

PyTorch is used only for device memory and streams: every wrapper takes torch tensors, passes raw
device pointers + strides, and launches on `torch.cuda.current_stream()`.  There is no CPU or
PyTorch fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os
import threading

import torch

F32, BF16 = 0, 1
A_LINEAR, A_CONV3X3, A_CONVT3 = 0, 1, 2
EPI_BIAS, EPI_ROWBIAS, EPI_RESIDUAL, EPI_GEGLU = 1, 2, 4, 8
DDIM_SQRT_AC, DDIM_SQRT_OMAC, DDIM_SQRT_A_PREV, DDIM_SIGMA, DDIM_X0_RESCALE, DDIM_DIR_COEF = range(6)
DDIM_NCOEF = 8
E_UNSUPPORTED = -3

#: PANDORA_B200_LIB selects a profiling build of the SAME library (build.py `variant`); tools only
_LIB_PATH = os.environ.get("PANDORA_B200_LIB") or os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libpandora_b200.so")


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p), ("c", C.c_void_p),
        ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("residual", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("lda", C.c_int64), ("ldc", C.c_int64), ("ldr", C.c_int64), ("ld_rowbias", C.c_int64),
        ("rowbias_rows", C.c_int64),
        ("a_dtype", C.c_int32), ("c_dtype", C.c_int32), ("epilogue", C.c_int32), ("a_mode", C.c_int32),
        ("cin", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("out_h", C.c_int32),
        ("out_w", C.c_int32), ("stride", C.c_int32), ("upsample", C.c_int32),
        ("frames", C.c_int32), ("frame_rows", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("colstats", C.c_void_p), ("ld_colstats", C.c_int64),
    ]


class AttnOperand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("bdiv", C.c_int64), ("bs_outer", C.c_int64),
                ("bs_inner", C.c_int64), ("rs", C.c_int64)]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", AttnOperand), ("k", AttnOperand), ("v", AttnOperand), ("out", C.c_void_p),
        ("o_bdiv", C.c_int64), ("o_bs_outer", C.c_int64), ("o_bs_inner", C.c_int64), ("o_rs", C.c_int64),
        ("batch", C.c_int64), ("heads", C.c_int64), ("n_q", C.c_int64), ("n_kv", C.c_int64),
        ("scale", C.c_float), ("dtype", C.c_int32), ("accumulate", C.c_int32), ("kv_split", C.c_int32),
    ]


SP_MAX_RANKS = 16


class SpBarrier(C.Structure):
    _fields_ = [("flags", C.c_void_p * SP_MAX_RANKS), ("epoch", C.c_void_p)]


class SpReshardArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * SP_MAX_RANKS), ("dst", C.c_void_p),
        ("ld_src", C.c_int64), ("ld_dst", C.c_int64), ("cols", C.c_int64),
        ("samples", C.c_int64), ("t_local", C.c_int64), ("hw_local", C.c_int64),
        ("P", C.c_int32), ("rank", C.c_int32), ("direction", C.c_int32), ("elem_bytes", C.c_int32),
        ("barrier", SpBarrier),
    ]


# every symbol include/pandora_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32, _f, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_size_t
SYMBOLS = {
    "pb_version": (C.c_int, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_device_supports_tc": (C.c_int, []),
    "pb_gemm": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "pb_tc_gemm": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "pb_tc_gemm_workspace_bytes": (_sz, [C.POINTER(GemmArgs)]),
    "pb_tc_gemm_colstats_ok": (C.c_int, [C.POINTER(GemmArgs)]),
    "pb_groupnorm_cols_ws_bytes": (_sz, [_i64, _i64, _i32]),
    "pb_groupnorm_apply_cols_ws_bytes": (_sz, [_i64, _i64, _i32]),
    "pb_groupnorm_apply_cols": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _vp, _i64, _i64, _i64, _i64, _i32, _f, _vp, _vp,
                                          _i32, _vp, _sz, _vp, _i64, _vp]),
    "pb_groupnorm_stats_cols": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pb_groupnorm_ws_bytes": (_sz, [_i64, _i64, _i64]),
    "pb_groupnorm_stats": (C.c_int, [_vp, _i32, _i64, _i64, _i64, _i64, _i32, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pb_groupnorm_apply": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _i32, _vp, _i64, _vp]),
    "pb_groupnorm_small": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _i32, _f, _vp, _vp, _i32, _vp]),
    "pb_groupnorm_sums": (C.c_int, [_vp, _i32, _i64, _i64, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
    "pb_groupnorm_affine": (C.c_int, [_vp, _i64, _i64, _i32, C.c_double, _f, _vp, _vp, _vp, _vp]),
    "pb_layernorm": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _f, _vp]),
    "pb_softmax_rows": (C.c_int, [_vp, _i64, _vp, _i32, _i64, _i64, _i64, _f, _vp]),
    "pb_attention": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "pb_tc_attention": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "pb_geglu": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp]),
    "pb_silu": (C.c_int, [_vp, _vp, _i64, _vp]),
    "pb_add": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "pb_copy_rows": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp]),
    "pb_upsample2x_rows": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _i64, _vp]),
    "pb_timestep_embedding": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "pb_video_to_rows": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "pb_rows_to_video": (C.c_int, [_vp, _i64, _vp, _i32, _i64, _i64, _i64, _i64, _vp]),
    "pb_sp_reshard": (C.c_int, [C.POINTER(SpReshardArgs), _vp]),
    "pb_sp_groupnorm_affine": (C.c_int, [C.POINTER(C.c_void_p), _i32, _i32, C.POINTER(SpBarrier), _i64, _i64, _i32,
                                         C.c_double, _f, _vp, _vp, _vp, _vp]),
    "pb_ddim_ws_bytes": (_sz, [_i64]),
    "pb_ddim_cfg_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _vp, _vp, _i32, _vp, _sz, _vp]),
    "pb_ddim_cfg3_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _vp, _vp, _i32, _vp,
                                    _sz, _vp]),
}

_lib = None
_lock = threading.Lock()
launches = 0  # kernels launched through this binding (bench.py reports it as gpu_launches)


# ---- optional per-launch timing (bench.py roofline / breakdown); off by default ----------------
_prof = None


def profile_begin():
    """Start recording (name, algorithmic flops, algorithmic bytes, CUDA-event pair) per C-ABI call."""
    global _prof
    _prof = []


def profile_end():
    """Stop recording; returns [dict(name, ms, flops, bytes)] (synchronises the device once)."""
    global _prof
    items, _prof = _prof or [], None
    torch.cuda.synchronize()
    return [dict(name=n, flops=f, bytes=b, detail=d, ms=e0.elapsed_time(e1)) for (n, f, b, d, e0, e1) in items]


def _tick(name, flops=0, nbytes=0, detail=""):
    if _prof is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return (name, flops, nbytes, detail() if callable(detail) else detail, e0)


def _tock(tok):
    if tok is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _prof.append(tok + (e1,))


def _nbytes(*ts):
    return sum(t.shape[0] * t.shape[1] * t.element_size() for t in ts if t is not None)


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library (once). Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} not found: build it with `python -m pandora_b200.build` "
                "(or __graft_entry__.build()). There is no CPU / PyTorch fallback for the hot path.")
        lib = C.CDLL(_LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def dt(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {load().pb_last_error().decode()}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _rows(t):
    """(ptr, ld) of a 2-D row view whose last dim is contiguous."""
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), (t.shape, t.stride())
    return t.data_ptr(), t.stride(0)


_tc = None


def supports_tc():
    global _tc
    if _tc is None:
        _tc = load().pb_device_supports_tc() == 1
    return _tc


def gemm(a, w, c, bias=None, rowbias=None, rowbias_rows=0, residual=None, geglu=False,
         mode=A_LINEAR, conv=None, convt=None, tc=False, M=None, K=None, colstats=None):
    """c = epilogue(gather(a) @ w.T).  a, c, residual are 2-D row views; w is [N, K] contiguous.
    colstats: optional fp32 row view [ceil(M/32), >= 2*N_out] that the tcgen05 epilogue fills with per-slab
    column (sum, sumsq) for the GroupNorm that follows.  Returns 0 (shape not supported by the tcgen05 path),
    1 (launched) or 2 (launched and colstats written)."""
    global launches
    g = GemmArgs()
    g.a, g.lda = _rows(a)
    g.w = w.data_ptr()
    g.c, g.ldc = _rows(c)
    g.N, g.K = w.shape[0], (w.shape[1] if K is None else K)
    g.M = c.shape[0] if M is None else M
    g.a_dtype, g.c_dtype = dt(a), dt(c)
    assert w.dtype == a.dtype and w.is_contiguous()
    epi = 0
    if bias is not None:
        assert bias.dtype == torch.float32
        g.bias, epi = bias.data_ptr(), epi | EPI_BIAS
    if rowbias is not None:
        assert rowbias.dtype == torch.float32
        g.rowbias, g.ld_rowbias = _rows(rowbias)
        g.rowbias_rows, epi = rowbias_rows, epi | EPI_ROWBIAS
    if residual is not None:
        assert residual.dtype == torch.float32
        g.residual, g.ldr = _rows(residual)
        epi |= EPI_RESIDUAL
    if geglu:
        epi |= EPI_GEGLU
    g.epilogue, g.a_mode = epi, mode
    if mode == A_CONV3X3:
        g.cin, g.in_h, g.in_w, g.out_h, g.out_w, g.stride, g.upsample = conv
    elif mode == A_CONVT3:
        g.cin, g.frames, g.frame_rows = convt
    lib = load()
    fn = lib.pb_tc_gemm if tc else lib.pb_gemm
    ws = None
    wrote_stats = False
    if tc and colstats is not None:
        assert colstats.dtype == torch.float32 and colstats.shape[0] * 32 >= g.M
        g.colstats, g.ld_colstats = _rows(colstats)
    if tc and g.M <= 9472:                     # split-K scratch only matters for launches with < 74 row tiles
        nb = lib.pb_tc_gemm_workspace_bytes(C.byref(g))
        if nb:
            ws = torch.empty(nb, dtype=torch.uint8, device=a.device)
            g.workspace, g.workspace_bytes = ws.data_ptr(), nb
    tok = _tick(("tc_" if tc else "simt_") + ("gemm", "conv3x3", "convt3")[mode] + ("_geglu" if geglu else ""),
                2 * g.M * g.N * g.K, (g.M * g.K // (1, 9, 3)[mode] + g.N * g.K) * a.element_size() + _nbytes(c, residual),
                lambda: f"M{g.M} N{g.N} K{g.K} epi{epi} out{'f32' if g.c_dtype == F32 else 'bf16'}"
                        + (f" s{g.stride}" if mode == A_CONV3X3 and g.stride != 1 else ""))
    if tc and colstats is not None:
        wrote_stats = lib.pb_tc_gemm_colstats_ok(C.byref(g)) == 1
    rc = fn(C.byref(g), _stream())
    if rc == E_UNSUPPORTED and tc:
        return 0
    _check(rc, "pb_tc_gemm" if tc else "pb_gemm")
    _tock(tok)
    launches += 1
    return 2 if wrote_stats else 1


def groupnorm(x, y, n_inst, rows_per_inst, gamma, beta, eps, silu, groups=32, colstats=None, x_copy=None):
    """y = GroupNorm(x) [* SiLU]; x, y 2-D row views [n_inst*rows_per_inst, C].  colstats: the column statistics
    the producer of x left behind (see `gemm`), which replace the statistics pass over x.  x_copy: optional row view of
    y's dtype that receives x itself (converted) from the same pass."""
    global launches
    lib = load()
    Cc = x.shape[1]
    xp, ldx = _rows(x)
    yp, ldy = _rows(y)
    st = _stream()
    cp_, ldcp = (None, 0) if x_copy is None else _rows(x_copy)
    assert x_copy is None or (x_copy.dtype == y.dtype and x_copy.shape == x.shape)
    if colstats is not None:
        assert rows_per_inst % 32 == 0 and colstats.shape[0] * 32 == x.shape[0] and colstats.shape[1] == 2 * Cc
        cp, ldcs = _rows(colstats)
        tok = _tick("groupnorm_fused", 0, _nbytes(x) + _nbytes(y), f"inst{n_inst} rows{rows_per_inst} C{Cc}")
        wsb = lib.pb_groupnorm_apply_cols_ws_bytes(n_inst, rows_per_inst, groups)
        ws = torch.empty(wsb // 8, dtype=torch.float64, device=x.device)
        _check(lib.pb_groupnorm_apply_cols(xp, dt(x), ldx, yp, dt(y), ldy, cp, ldcs, n_inst, rows_per_inst, Cc, groups,
                                           eps, gamma.data_ptr(), beta.data_ptr(), int(silu), ws.data_ptr(), wsb, cp_, ldcp,
                                           st), "pb_groupnorm_apply_cols")
        _tock(tok)
        launches += 2
        return None
    else:
        # small tensors (<= 8 MiB): one launch, a CTA per (instance, group); the second pass over the slab hits L1 / L2
        if x_copy is None and x.numel() * x.element_size() <= (8 << 20) and n_inst * groups >= 32 \
                and (Cc // groups) % 4 == 0:
            tok = _tick("groupnorm_small", 0, _nbytes(x) + _nbytes(y), f"inst{n_inst} rows{rows_per_inst} C{Cc}")
            rc = lib.pb_groupnorm_small(xp, dt(x), ldx, yp, dt(y), ldy, n_inst, rows_per_inst, Cc, groups, eps,
                                        gamma.data_ptr(), beta.data_ptr(), int(silu), st)
            if rc == 0:
                _tock(tok)
                launches += 1
                return None
            if rc != E_UNSUPPORTED:
                _check(rc, "pb_groupnorm_small")
        ws_bytes = lib.pb_groupnorm_ws_bytes(n_inst, rows_per_inst, Cc)
        ws = torch.empty(ws_bytes // 4 + n_inst * 2 * Cc, dtype=torch.float32, device=x.device)
        stats = ws[ws_bytes // 4:]
        tok = _tick("groupnorm", 0, 2 * _nbytes(x) + _nbytes(y), f"inst{n_inst} rows{rows_per_inst} C{Cc}")
        _check(lib.pb_groupnorm_stats(xp, dt(x), ldx, n_inst, rows_per_inst, Cc, groups, eps, gamma.data_ptr(),
                                      beta.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws_bytes, st),
               "pb_groupnorm_stats")
        launches += 3
    _check(lib.pb_groupnorm_apply(xp, dt(x), ldx, yp, dt(y), ldy, n_inst, rows_per_inst, Cc,
                                  stats.data_ptr(), int(silu), cp_, ldcp, st), "pb_groupnorm_apply")
    _tock(tok)
    return stats


def groupnorm_sp(x, y, n_inst, rows_per_inst, gamma, beta, eps, silu, all_reduce, global_rows, groups=32):
    """GroupNorm whose statistics span ranks: local fp64 sums -> all_reduce(sums) -> affine -> apply."""
    global launches
    lib = load()
    Cc = x.shape[1]
    ws_bytes = lib.pb_groupnorm_ws_bytes(n_inst, rows_per_inst, Cc)
    ws = torch.empty(ws_bytes // 4 + n_inst * 2 * Cc, dtype=torch.float32, device=x.device)
    stats = ws[ws_bytes // 4:]
    sums = torch.empty(n_inst * groups * 2, dtype=torch.float64, device=x.device)
    xp, ldx = _rows(x)
    yp, ldy = _rows(y)
    st = _stream()
    _check(lib.pb_groupnorm_sums(xp, dt(x), ldx, n_inst, rows_per_inst, Cc, groups, sums.data_ptr(), ws.data_ptr(),
                                 ws_bytes, st), "pb_groupnorm_sums")
    all_reduce(sums)
    _check(lib.pb_groupnorm_affine(sums.data_ptr(), n_inst, Cc, groups, float(global_rows * (Cc // groups)), eps,
                                   gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), st), "pb_groupnorm_affine")
    _check(lib.pb_groupnorm_apply(xp, dt(x), ldx, yp, dt(y), ldy, n_inst, rows_per_inst, Cc,
                                  stats.data_ptr(), int(silu), None, 0, st), "pb_groupnorm_apply")
    launches += 4


def _sp_barrier(flag_ptrs, epoch_ptr):
    b = SpBarrier()
    if epoch_ptr:
        for i, f in enumerate(flag_ptrs):
            b.flags[i] = f
        b.epoch = epoch_ptr
    return b


def sp_reshard(src_ptrs, ld_src, dst, direction, rank, samples, t_local, hw_local, cols, elem_bytes,
               flag_ptrs=None, epoch_ptr=None):
    """Pull this rank's share of every rank's send buffer (`src_ptrs`, device addresses valid on THIS device) into the
    row view `dst`: direction 0 = frames -> pixels, 1 = pixels -> frames (see pb_sp_reshard)."""
    global launches
    a = SpReshardArgs()
    P = len(src_ptrs)
    for i, p in enumerate(src_ptrs):
        a.src[i] = p
    a.dst, a.ld_dst = _rows(dst)
    a.ld_src, a.cols = ld_src, cols
    a.samples, a.t_local, a.hw_local = samples, t_local, hw_local
    a.P, a.rank, a.direction, a.elem_bytes = P, rank, direction, elem_bytes
    a.barrier = _sp_barrier(flag_ptrs, epoch_ptr)
    nb = samples * t_local * P * hw_local * cols * elem_bytes
    tok = _tick("sp_reshard", 0, 2 * nb, f"P{P} dir{direction} rows{samples * t_local * P * hw_local} C{cols}")
    _check(load().pb_sp_reshard(C.byref(a), _stream()), "pb_sp_reshard")
    _tock(tok)
    launches += 1


def groupnorm_sp_peers(x, y, n_inst, rows_per_inst, gamma, beta, eps, silu, my_sums, sums_ptrs, rank, global_rows,
                       flag_ptrs=None, epoch_ptr=None, before_reduce=None, groups=32):
    """GroupNorm whose statistics span ranks, without NCCL: local fp64 sums into this rank's peer-visible slot
    (`my_sums`), one kernel that adds every rank's partials (`sums_ptrs`) behind a flag barrier, apply."""
    global launches
    lib = load()
    Cc = x.shape[1]
    ws_bytes = lib.pb_groupnorm_ws_bytes(n_inst, rows_per_inst, Cc)
    ws = torch.empty(ws_bytes // 4 + n_inst * 2 * Cc, dtype=torch.float32, device=x.device)
    stats = ws[ws_bytes // 4:]
    xp, ldx = _rows(x)
    yp, ldy = _rows(y)
    st = _stream()
    assert my_sums.dtype == torch.float64 and my_sums.numel() >= n_inst * groups * 2
    _check(lib.pb_groupnorm_sums(xp, dt(x), ldx, n_inst, rows_per_inst, Cc, groups, my_sums.data_ptr(), ws.data_ptr(),
                                 ws_bytes, st), "pb_groupnorm_sums")
    if before_reduce is not None:
        before_reduce()
    P = len(sums_ptrs)
    arr = (C.c_void_p * P)(*sums_ptrs)
    bar = _sp_barrier(flag_ptrs, epoch_ptr)
    _check(lib.pb_sp_groupnorm_affine(arr, P, rank, C.byref(bar), n_inst, Cc, groups,
                                      float(global_rows * (Cc // groups)), eps, gamma.data_ptr(), beta.data_ptr(),
                                      stats.data_ptr(), st), "pb_sp_groupnorm_affine")
    _check(lib.pb_groupnorm_apply(xp, dt(x), ldx, yp, dt(y), ldy, n_inst, rows_per_inst, Cc,
                                  stats.data_ptr(), int(silu), None, 0, st), "pb_groupnorm_apply")
    launches += 4


def layernorm(x, y, gamma, beta, eps=1e-5):
    global launches
    xp, ldx = _rows(x)
    yp, ldy = _rows(y)
    tok = _tick("layernorm", 0, _nbytes(x, y), f"rows{x.shape[0]} C{x.shape[1]}")
    _check(load().pb_layernorm(xp, dt(x), ldx, yp, dt(y), ldy, x.shape[0], x.shape[1],
                               gamma.data_ptr(), beta.data_ptr(), eps, _stream()), "pb_layernorm")
    _tock(tok)
    launches += 1


def softmax_rows(x, y, scale):
    """y[r] = softmax(x[r] * scale); x fp32 rows, y fp32 / bf16 rows of the same shape."""
    global launches
    assert x.dtype == torch.float32 and x.shape == y.shape
    xp, ldx = _rows(x)
    yp, ldy = _rows(y)
    tok = _tick("softmax_rows", 0, _nbytes(x, y), f"rows{x.shape[0]} n{x.shape[1]}")
    _check(load().pb_softmax_rows(xp, ldx, yp, dt(y), ldy, x.shape[0], x.shape[1], scale, _stream()), "pb_softmax_rows")
    _tock(tok)
    launches += 1


def operand(t, bdiv, bs_outer, bs_inner, rs, offset=0):
    o = AttnOperand()
    o.ptr = t.data_ptr() + offset * t.element_size()
    o.bdiv, o.bs_outer, o.bs_inner, o.rs = bdiv, bs_outer, bs_inner, rs
    return o


def attention(q, k, v, out, o_addr, batch, heads, n_q, n_kv, scale, dtype, accumulate=False, tc=False, kv_split=0):
    """q, k, v: AttnOperand; out: tensor; o_addr = (bdiv, bs_outer, bs_inner, rs)."""
    global launches
    a = AttnArgs()
    a.q, a.k, a.v = q, k, v
    a.out = out.data_ptr()
    a.o_bdiv, a.o_bs_outer, a.o_bs_inner, a.o_rs = o_addr
    a.batch, a.heads, a.n_q, a.n_kv = batch, heads, n_q, n_kv
    a.scale, a.dtype, a.accumulate, a.kv_split = scale, dtype, int(accumulate), int(kv_split)
    lib = load()
    fn = lib.pb_tc_attention if tc else lib.pb_attention
    esz = 2 if dtype == BF16 else 4
    kind = "temporal" if n_q <= 32 else ("self" if n_q == n_kv else "cross")
    tok = _tick(("tc_" if tc else "simt_") + "attn_" + kind, 4 * batch * heads * n_q * n_kv * 64,
                batch * heads * 64 * esz * (2 * n_q + 2 * n_kv), f"B{batch} h{heads} nq{n_q} nkv{n_kv}")
    rc = fn(C.byref(a), _stream())
    if rc == E_UNSUPPORTED and tc:
        return False
    _check(rc, "pb_tc_attention" if tc else "pb_attention")
    _tock(tok)
    launches += 1
    return True


def geglu(u, y):
    global launches
    up, ldu = _rows(u)
    yp, ldy = _rows(y)
    tok = _tick("geglu", 0, _nbytes(u, y))
    _check(load().pb_geglu(up, dt(u), ldu, yp, dt(y), ldy, u.shape[0], y.shape[1], _stream()), "pb_geglu")
    _tock(tok)
    launches += 1


def silu(x, y):
    global launches
    _check(load().pb_silu(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "pb_silu")
    launches += 1


def add(a, b, y):
    global launches
    _check(load().pb_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), y.numel(), _stream()), "pb_add")
    launches += 1


def copy_rows(src, dst):
    global launches
    sp, lds = _rows(src)
    dp, ldd = _rows(dst)
    tok = _tick("copy_rows", 0, _nbytes(src, dst), f"rows{src.shape[0]} C{src.shape[1]} {src.dtype}->{dst.dtype}")
    _check(load().pb_copy_rows(sp, dt(src), lds, dp, dt(dst), ldd, src.shape[0], src.shape[1], _stream()),
           "pb_copy_rows")
    _tock(tok)
    launches += 1


def upsample2x_rows(src, dst, n_img, H, W):
    global launches
    sp, lds = _rows(src)
    dp, ldd = _rows(dst)
    tok = _tick("upsample2x", 0, _nbytes(src, dst))
    _check(load().pb_upsample2x_rows(sp, dt(src), lds, dp, dt(dst), ldd, n_img, H, W, src.shape[1], _stream()),
           "pb_upsample2x_rows")
    _tock(tok)
    launches += 1


def timestep_embedding(t, freqs, out):
    global launches
    assert t.dtype == torch.int64 and t.is_contiguous() and freqs.dtype == torch.float32
    _check(load().pb_timestep_embedding(t.data_ptr(), freqs.data_ptr(), out.data_ptr(), t.shape[0],
                                        freqs.shape[0], _stream()), "pb_timestep_embedding")
    launches += 1


def video_to_rows(x1, x2, rows, b, t, hw):
    global launches
    assert x1.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x1.dtype))
    rp, ld = _rows(rows)
    _check(load().pb_video_to_rows(x1.data_ptr(), x1.shape[1], _ptr(x2), 0 if x2 is None else x2.shape[1],
                                   dt(x1), rp, ld, b, t, hw, _stream()), "pb_video_to_rows")
    launches += 1


def rows_to_video(rows, out, b, c, t, hw):
    global launches
    rp, ld = _rows(rows)
    assert out.is_contiguous() and rows.dtype == torch.float32
    _check(load().pb_rows_to_video(rp, ld, out.data_ptr(), dt(out), b, c, t, hw, _stream()), "pb_rows_to_video")
    launches += 1


def ddim_cfg_step(x, v_cond, v_uncond, noise, x_prev, pred_x0, cfg_scale, guidance_rescale, temperature,
                  tables, step_index, advance, ws, v_uncond_img=None, cfg_img=0.0):
    """p_sample_ddim after the U-Net calls; `v_uncond_img` selects the three-branch combine of
    ddim_multiplecond.py:234 (pb_ddim_cfg3_step)."""
    global launches
    b = x.shape[0]
    n = x.numel() // b
    for t_ in (x, v_cond, noise, x_prev, pred_x0):
        assert t_.dtype == torch.float32 and t_.is_contiguous()
    for t_ in (v_uncond, v_uncond_img):
        assert t_ is None or (t_.dtype == torch.float32 and t_.is_contiguous())
    assert v_uncond_img is None or v_uncond is not None
    assert tables.dtype == torch.float32 and step_index.dtype == torch.int32
    _check(load().pb_ddim_cfg3_step(x.data_ptr(), v_cond.data_ptr(), _ptr(v_uncond), _ptr(v_uncond_img),
                                    noise.data_ptr(), x_prev.data_ptr(), pred_x0.data_ptr(), b, n, cfg_scale,
                                    float(cfg_img), guidance_rescale, temperature, tables.data_ptr(),
                                    step_index.data_ptr(), int(advance), _ptr(ws),
                                    0 if ws is None else ws.numel() * ws.element_size(), _stream()),
           "pb_ddim_cfg3_step")
    launches += 1 + int(bool(advance)) + (2 if guidance_rescale > 0 and v_uncond is not None else 0)
