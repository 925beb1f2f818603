// norm.cu — LayerNorm as an HBM-streaming kernel (GroupNorm lives in groupnorm.cu).
//
// LayerNorm (attention.py:225-227): per token over C channels, eps 1e-5; reads the fp32 trunk, writes the bf16 (or
// fp32) operand of the GEMM that follows.
//   layernorm_rows_kernel  the shapes of this U-Net (C = 320 / 512 / 640 / 1280): a row is spread over L = 16 or 32
//                          lanes with EXACTLY NV vectors per lane (no idle lanes, no per-vector predicates), two rows
//                          per warp when L = 16; gamma / beta live in registers (NV <= 5), the next row group is
//                          prefetched under the two shuffle reductions; ~8 instructions per element.
//   layernorm_kernel       any C % 4 == 0 up to 1280: one warp per row, predicated vectors.
// (ncu on the generic kernel at C = 320: 21 instructions per element, issue slots 58 % busy at 47 % of the DRAM rate —
// bound by instruction issue, not by memory.)
#include "common.cuh"

namespace pb {

// One warp per row, the row held in registers (C % 4 == 0, C <= 1280 = 32 lanes x 10 vectors).  A warp walks
// rows with a grid stride and loads row i+1 before it reduces row i, so the two dependent warp reductions of a
// row overlap the next row's HBM latency; NV = vectors per lane is a template parameter so the C = 320 rows
// (3 vectors) do not pay for the C = 1280 ones (10).
constexpr int LN_MAXV = 10;
template <typename TX, typename TY, int NV, bool PREFETCH>
__global__ void __launch_bounds__(256)
layernorm_kernel(const TX* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy, int64_t rows,
                 int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  pb::pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int cv = C >> 2;
  const float inv_c = 1.0f / (float)C;
  float cur[NV][4], nxt[NV][4];
  auto load_row = [&](int64_t row, float (&dst)[NV][4]) {
    const TX* xr = x + row * ldx;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < cv) load4(xr + v * 4, dst[i]);
      else { dst[i][0] = dst[i][1] = dst[i][2] = dst[i][3] = 0.f; }
    }
  };
  if (warp0 < rows) load_row(warp0, cur);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const bool more = PREFETCH && row + nwarps < rows;
    if (more) load_row(row + nwarps, nxt);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (cur[i][0] + cur[i][1]) + (cur[i][2] + cur[i][3]);
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < cv) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = cur[i][j] - mean; q = fmaf(d, d, q); }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    TY* yr = y + row * ldy;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + 32 * i;
      if (v < cv) {
        float g4[4], b4[4], o[4];
        load4(gamma + v * 4, g4);
        load4(beta + v * 4, b4);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (cur[i][j] - mean) * rstd * g4[j] + b4[j];
        store4(yr + v * 4, o);
      }
    }
    if (PREFETCH) {
      if (more) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) cur[i][j] = nxt[i][j];
        }
      }
    } else if (row + nwarps < rows) {
      load_row(row + nwarps, cur);
    }
  }
}

// ---- exact-fit variant: cv == L * NV
// PF = false: one row group per warp and no prefetch buffer (short inputs: C = 1280 at the deep levels is 5120 rows =
// 640 CTAs; without the second row buffer the kernel fits two CTAs per SM, i.e. half as many waves of load latency).
template <typename TX, typename TY, int L, int NV, bool HOIST, bool PF = true>
__global__ void __launch_bounds__(256, PF ? 1 : 2)
layernorm_rows_kernel(const TX* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy, int64_t rows,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  pb::pdl_trigger();
  constexpr int RPW = 32 / L;                       // rows per warp pass
  constexpr float INV_C = 1.0f / (float)(4 * L * NV);
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, sl = lane % L;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t groups = (rows + RPW - 1) / RPW;
  float gh[HOIST ? NV : 1][4], bh[HOIST ? NV : 1][4];
  if (HOIST) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { load4(gamma + (sl + L * i) * 4, gh[HOIST ? i : 0]); load4(beta + (sl + L * i) * 4, bh[HOIST ? i : 0]); }
  }
  float cur[NV][4], nxt[PF ? NV : 1][4];
  auto load_row = [&](int64_t g, float (&dst)[NV][4]) {
    int64_t row = g * RPW + sub;
    if (row >= rows) row = rows - 1;                // tail: load a valid row, the store is skipped
    const TX* xr = x + row * ldx;
#pragma unroll
    for (int i = 0; i < NV; ++i) load4(xr + (sl + L * i) * 4, dst[i]);
  };
  if (warp0 < groups) load_row(warp0, cur);
  for (int64_t g = warp0; g < groups; g += nwarps) {
    const bool more = g + nwarps < groups;
    if constexpr (PF) {
      if (more) load_row(g + nwarps, nxt);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (cur[i][0] + cur[i][1]) + (cur[i][2] + cur[i][3]);
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * INV_C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = cur[i][j] - mean; q = fmaf(d, d, q); }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * INV_C + eps);
    const float nb = -mean * rstd;
    const int64_t row = g * RPW + sub;
    if (row < rows) {
      TY* yr = y + row * ldy;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float g4[4], b4[4], o[4];
        if (HOIST) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { g4[j] = gh[HOIST ? i : 0][j]; b4[j] = bh[HOIST ? i : 0][j]; }
        } else {
          load4(gamma + (sl + L * i) * 4, g4);
          load4(beta + (sl + L * i) * 4, b4);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = fmaf(fmaf(cur[i][j], rstd, nb), g4[j], b4[j]);
        store4(yr + (sl + L * i) * 4, o);
      }
    }
    if constexpr (PF) {
      if (more) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
          for (int j = 0; j < 4; ++j) cur[i][j] = nxt[i][j];
        }
      }
    } else {
      break;                                        // launched with one row group per warp
    }
  }
}

// ---- row softmax (the single-head, d = C attention of the VAE decoder's AttnBlock, ae_modules.py:53-67: the score
// matrix of one frame is [HW, HW] fp32 out of a GEMM; this turns every row into probabilities in the operand dtype of
// the P V product).  One CTA per row, online (max, sum) in one pass, second pass writes.
template <typename TY>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy, int n, float scale_log2) {
  pb::pdl_trigger();
  __shared__ float red_m[8], red_l[8];
  const float* xr = x + (int64_t)blockIdx.x * ldx;
  TY* yr = y + (int64_t)blockIdx.x * ldy;
  const int nv = n >> 2;
  float m = -INFINITY, l = 0.f;
  for (int v = threadIdx.x; v < nv; v += 256) {
    float t[4];
    load4(xr + v * 4, t);
    const float mv = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])) * scale_log2;
    if (mv > m) { l *= exp2f(m - mv); m = mv; }
#pragma unroll
    for (int j = 0; j < 4; ++j) l += exp2f(fmaf(t[j], scale_log2, -m));
  }
  for (int i = nv * 4 + threadIdx.x; i < n; i += 256) {
    const float s = xr[i] * scale_log2;
    if (s > m) { l *= exp2f(m - s); m = s; }
    l += exp2f(s - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float mm = fmaxf(m, m2);
    l = (m == -INFINITY ? 0.f : l * exp2f(m - mm)) + (m2 == -INFINITY ? 0.f : l2 * exp2f(m2 - mm));
    m = mm;
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red_m[w] = m; red_l[w] = l; }
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) M = fmaxf(M, red_m[i]);
  float L = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) L += red_m[i] == -INFINITY ? 0.f : red_l[i] * exp2f(red_m[i] - M);
  const float inv = 1.0f / L;
  for (int v = threadIdx.x; v < nv; v += 256) {
    float t[4], o[4];
    load4(xr + v * 4, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = exp2f(fmaf(t[j], scale_log2, -M)) * inv;
    store4(yr + v * 4, o);
  }
  for (int i = nv * 4 + threadIdx.x; i < n; i += 256) yr[i] = from_f<TY>(exp2f(fmaf(xr[i], scale_log2, -M)) * inv);
}

}  // namespace pb

extern "C" {

int pb_softmax_rows(const float* x, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy, int64_t rows, int64_t n,
                    float scale, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && rows > 0 && n > 0 && rows < ((int64_t)1 << 31) && n < ((int64_t)1 << 30), "pb_softmax_rows: bad arguments");
  PB_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldx >= n && ldy >= n, "pb_softmax_rows: row pitches must be multiples of 4");
  const float scale_log2 = scale * 1.4426950408889634f;
  PB_DISPATCH_DTYPE(y_dtype, TY, {
    softmax_rows_kernel<TY><<<(unsigned)rows, 256, 0, to_stream(stream)>>>(x, ldx, reinterpret_cast<TY*>(y), ldy, (int)n,
                                                                          scale_log2);
  })
  PB_LAUNCH_CHECK("softmax_rows_kernel");
  return PB_OK;
}


int pb_layernorm(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                 int64_t rows, int64_t C, const float* gamma, const float* beta, float eps,
                 pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && gamma && beta, "pb_layernorm: null pointer");
  PB_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0,
             "pb_layernorm: need rows > 0, C %% 4 == 0, ld %% 4 == 0");
  PB_REQUIRE(C <= 128 * LN_MAXV, "pb_layernorm: C=%lld exceeds the register-resident limit %d",
             (long long)C, 128 * LN_MAXV);
  cudaStream_t st = to_stream(stream);
  // exact-fit shapes of the U-Net: C = 320 (16 lanes x 5 vectors, two rows per warp), 512, 640, 1280
  {
    const int cv = (int)(C / 4);
    int64_t fb = 148 * 8;                            // grid-stride over row groups, prefetching the next one
#define PB_LN_ROWS(L_, NV_, H_)                                                                                   \
    do {                                                                                                           \
      const int64_t groups = ceil_div(rows, 32 / (L_));                                                            \
      const bool single = ceil_div(groups, 8) <= fb;       /* one row group per warp: no prefetch buffer */         \
      if (single) fb = ceil_div(groups, 8);                                                                        \
      PB_DISPATCH_DTYPE(x_dtype, TX, {                                                                             \
        PB_DISPATCH_DTYPE(y_dtype, TY, {                                                                           \
          if (single && (NV_) >= 10)                                                                               \
            layernorm_rows_kernel<TX, TY, L_, NV_, H_, false><<<(unsigned)fb, 256, 0, st>>>(                       \
                reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows, gamma, beta, eps);       \
          else                                                                                                     \
            layernorm_rows_kernel<TX, TY, L_, NV_, H_><<<(unsigned)fb, 256, 0, st>>>(                              \
                reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows, gamma, beta, eps);       \
        })                                                                                                         \
      })                                                                                                           \
      PB_LAUNCH_CHECK("layernorm_rows_kernel");                                                                    \
      return PB_OK;                                                                                                \
    } while (0)
    if (cv == 80) PB_LN_ROWS(16, 5, true);
    if (cv == 128) PB_LN_ROWS(32, 4, true);
    if (cv == 160) PB_LN_ROWS(32, 5, true);
    if (cv == 320) PB_LN_ROWS(32, 10, false);
#undef PB_LN_ROWS
  }
  int64_t blocks = ceil_div(rows * 32, 256);
  // long inputs: 8 CTAs x 8 warps per SM, warps stride over rows and prefetch the next row; short inputs (about
  // one row per warp) keep one row per warp and the smaller register footprint
  const bool prefetch = blocks > 148 * 8 * 2;
  if (prefetch) blocks = 148 * 8;
  const int nv = (int)ceil_div(C / 4, 32);
#define PB_LN_LAUNCH(NV)                                                                                       \
  if (prefetch)                                                                                                \
    layernorm_kernel<TX, TY, NV, true><<<(unsigned)blocks, 256, 0, st>>>(                                      \
        reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows, (int)C, gamma, beta, eps);   \
  else                                                                                                         \
    layernorm_kernel<TX, TY, NV, false><<<(unsigned)blocks, 256, 0, st>>>(                                     \
        reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows, (int)C, gamma, beta, eps)
  PB_DISPATCH_DTYPE(x_dtype, TX, {
    PB_DISPATCH_DTYPE(y_dtype, TY, {
      if (nv <= 3) { PB_LN_LAUNCH(3); }
      else if (nv <= 5) { PB_LN_LAUNCH(5); }
      else { PB_LN_LAUNCH(LN_MAXV); }
    })
  })
#undef PB_LN_LAUNCH
  PB_LAUNCH_CHECK("layernorm_kernel");
  return PB_OK;
}

}  // extern "C"
