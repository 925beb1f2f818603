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

// One warp per row, the row held in registers (C % 4 == 0, C <= 1280 = 32 lanes x 10 vectors).
constexpr int LN_MAXV = 10;
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
layernorm_kernel(const TX* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy, int64_t rows,
                 int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int cv = C >> 2;
  float buf[LN_MAXV][4];
  float s = 0.f;
  const TX* xr = x + row * ldx;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < cv) {
      load4(xr + v * 4, buf[i]);
      s += (buf[i][0] + buf[i][1]) + (buf[i][2] + buf[i][3]);
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < cv) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { float d = buf[i][j] - mean; q = fmaf(d, d, q); }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  TY* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < cv) {
      float g4[4], b4[4], o[4];
      load4(gamma + v * 4, g4);
      load4(beta + v * 4, b4);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = (buf[i][j] - mean) * rstd * g4[j] + b4[j];
      store4(yr + v * 4, o);
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
  const int64_t blocks = ceil_div(rows * 32, 256);
  PB_DISPATCH_DTYPE(x_dtype, TX, {
    PB_DISPATCH_DTYPE(y_dtype, TY, {
      layernorm_kernel<TX, TY><<<(unsigned)blocks, 256, 0, st>>>(
          reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows, (int)C, gamma, beta, eps);
    })
  })
  PB_LAUNCH_CHECK("layernorm_kernel");
  return PB_OK;
}

}  // extern "C"
