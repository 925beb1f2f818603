// norm.cu — LayerNorm as an HBM-streaming kernel (GroupNorm lives in groupnorm.cu).
//
// GroupNorm (basics.py:76-88; openaimodel3d.py:155,179,259-268,547; attention.py:265,331):
//   rows layout [n_inst * R, C]; statistics per (instance, group) over R rows x C/G channels.
//   Pass 1  gn_partial_kernel : grid (splits, n_inst); every thread owns one 4-channel vector
//           column and walks rows with a fixed stride (4 loads in flight), so loads are fully
//           coalesced and the accumulators live in registers; per-channel partial (sum, sumsq) go to
//           the workspace.
//   Pass 2  gn_finalize_kernel: one warp per (instance, group) folds splits x C/G partials in fp64
//           -> mean, rstd -> per-channel (scale, shift).
//   Pass 3  gn_apply_kernel   : y = x*scale + shift (+SiLU), vectorised, no integer division in the
//           loop, written in the operand dtype the next GEMM wants (bf16 or fp32).
// LayerNorm (attention.py:225-227): one warp per row, two-pass in registers.
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

}  // namespace pb

extern "C" {

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
