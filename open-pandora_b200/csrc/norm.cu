// norm.cu — GroupNorm(32)(+SiLU) and LayerNorm as HBM-streaming kernels.
//
// GroupNorm (basics.py:76-88; openaimodel3d.py:155,179,259-268,547; attention.py:265,331):
//   rows layout [n_inst * R, C]; statistics per (instance, group) over R rows x C/G channels.
//   Pass 1  gn_partial_kernel : grid (splits, n_inst); every thread owns one 4-channel vector
//           column and walks rows with a fixed stride, so loads are fully coalesced and the
//           accumulators live in registers; per-channel partial (sum, sumsq) go to the workspace.
//   Pass 2  gn_finalize_kernel: one warp per (instance, group) folds splits x C/G partials in fp64
//           -> mean, rstd.
//   Pass 3  gn_apply_kernel   : y = (x-mean)*rstd*gamma+beta (+SiLU), vectorised, written in the
//           operand dtype the next GEMM wants (bf16 or fp32).
// LayerNorm (attention.py:225-227): one warp per row, two-pass in registers.
#include "common.cuh"

namespace pb {

constexpr int GN_THREADS = 256;

static inline int gn_splits(int64_t n_inst, int64_t R) {
  // enough CTAs to fill 148 SMs a few times over, but at least 8 rows per CTA
  int64_t want = ceil_div(148 * 4, n_inst);
  int64_t maxs = ceil_div(R, 8);
  int64_t s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  if (s > 1024) s = 1024;
  return (int)s;
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const T* __restrict__ x, int64_t ldx, int64_t R, int C, int splits,
                  float* __restrict__ ws /* [n_inst, splits, C, 2] */) {
  const int inst = blockIdx.y, sp = blockIdx.x;
  const int cv = C >> 2;                       // vectors per row
  const int lanes = GN_THREADS / cv > 0 ? GN_THREADS / cv : 1;   // row lanes per CTA
  const int64_t r_begin = R * sp / splits, r_end = R * (sp + 1) / splits;
  const T* base = x + (int64_t)inst * R * ldx;
  float* out = ws + ((int64_t)inst * splits + sp) * C * 2;
  extern __shared__ float red[];               // [lanes][C][2] when lanes > 1

  // cv may exceed the CTA (C=2560 -> 640 vectors): loop over vector columns
  for (int v0 = 0; v0 < cv; v0 += GN_THREADS) {
    int v, lane;
    if (cv >= GN_THREADS) { v = v0 + threadIdx.x; lane = 0; }
    else { v = threadIdx.x % cv; lane = threadIdx.x / cv; }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    const bool active = (v < cv) && (lane < lanes);
    if (active) {
      const int step = cv >= GN_THREADS ? 1 : lanes;
      for (int64_t r = r_begin + lane; r < r_end; r += step) {
        float t[4];
        load4(base + r * ldx + (int64_t)v * 4, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] += t[i]; q[i] = fmaf(t[i], t[i], q[i]); }
      }
    }
    if (cv >= GN_THREADS) {
      if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { out[(v * 4 + i) * 2] = s[i]; out[(v * 4 + i) * 2 + 1] = q[i]; }
      }
    } else {
      // fold the row lanes through shared memory
      if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          red[((lane * C) + v * 4 + i) * 2] = s[i];
          red[((lane * C) + v * 4 + i) * 2 + 1] = q[i];
        }
      }
      __syncthreads();
      for (int c = threadIdx.x; c < C; c += GN_THREADS) {
        float ss = 0.f, qq = 0.f;
        for (int l = 0; l < lanes; ++l) { ss += red[(l * C + c) * 2]; qq += red[(l * C + c) * 2 + 1]; }
        out[c * 2] = ss; out[c * 2 + 1] = qq;
      }
      break;  // cv < GN_THREADS: single pass
    }
  }
}

__global__ void gn_finalize_kernel(const float* __restrict__ ws, int splits, int C, int G, int64_t R,
                                   float eps, float* __restrict__ mean_rstd, int total) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= total) return;
  const int inst = warp / G, g = warp % G;
  const int cpg = C / G;
  double s = 0.0, q = 0.0;
  const int n = splits * cpg;
  for (int i = lane; i < n; i += 32) {
    int sp = i / cpg, c = g * cpg + i % cpg;
    const float* p = ws + (((int64_t)inst * splits + sp) * C + c) * 2;
    s += (double)p[0]; q += (double)p[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    double cnt = (double)R * cpg;
    double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[((int64_t)inst * G + g) * 2] = (float)mean;
    mean_rstd[((int64_t)inst * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const TX* __restrict__ x, int64_t ldx, TY* __restrict__ y, int64_t ldy, int64_t R,
                int C, int G, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                const float* __restrict__ beta, int silu, int64_t total_vec) {
  const int cv = C >> 2, cpg = C / G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cv;
    const int v = (int)(i - row * cv);
    const int64_t inst = row / R;
    float t[4], g4[4], b4[4], o[4];
    load4(x + row * ldx + (int64_t)v * 4, t);
    load4(gamma + v * 4, g4);
    load4(beta + v * 4, b4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int grp = (v * 4 + j) / cpg;
      const float mean = __ldg(mean_rstd + (inst * G + grp) * 2);
      const float rstd = __ldg(mean_rstd + (inst * G + grp) * 2 + 1);
      float u = (t[j] - mean) * rstd * g4[j] + b4[j];
      o[j] = silu ? silu_f(u) : u;
    }
    store4(y + row * ldy + (int64_t)v * 4, o);
  }
}

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

size_t pb_groupnorm_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int64_t C) {
  if (n_inst <= 0 || rows_per_inst <= 0 || C <= 0) return 0;
  return (size_t)n_inst * pb::gn_splits(n_inst, rows_per_inst) * C * 2 * sizeof(float);
}

int pb_groupnorm_stats(const void* x, int32_t x_dtype, int64_t ldx, int64_t n_inst,
                       int64_t rows_per_inst, int64_t C, int32_t groups, float eps,
                       float* mean_rstd, void* ws, size_t ws_bytes, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && mean_rstd && ws, "pb_groupnorm_stats: null pointer");
  PB_REQUIRE(n_inst > 0 && rows_per_inst > 0 && C > 0 && groups > 0, "pb_groupnorm_stats: empty input");
  PB_REQUIRE(C % 4 == 0 && C % groups == 0 && ldx % 4 == 0,
             "pb_groupnorm_stats: need C %% 4 == 0, C %% groups == 0, ldx %% 4 == 0 (C=%lld)", (long long)C);
  PB_REQUIRE(n_inst <= 65535, "pb_groupnorm_stats: n_inst too large");
  const int splits = gn_splits(n_inst, rows_per_inst);
  PB_REQUIRE(ws_bytes >= pb_groupnorm_ws_bytes(n_inst, rows_per_inst, C),
             "pb_groupnorm_stats: workspace too small");
  cudaStream_t st = to_stream(stream);
  const int cv = (int)(C / 4);
  const int lanes = cv >= GN_THREADS ? 1 : GN_THREADS / cv;
  const size_t smem = cv >= GN_THREADS ? 0 : (size_t)lanes * C * 2 * sizeof(float);
  PB_REQUIRE(smem <= 48 * 1024, "pb_groupnorm_stats: C=%lld needs %zu B smem", (long long)C, smem);
  dim3 grid(splits, (unsigned)n_inst);
  PB_DISPATCH_DTYPE(x_dtype, T, {
    gn_partial_kernel<T><<<grid, GN_THREADS, smem, st>>>(reinterpret_cast<const T*>(x), ldx, rows_per_inst,
                                                         (int)C, splits, reinterpret_cast<float*>(ws));
  })
  PB_LAUNCH_CHECK("gn_partial_kernel");
  const int total = (int)(n_inst * groups);
  gn_finalize_kernel<<<(unsigned)ceil_div((int64_t)total * 32, 256), 256, 0, st>>>(
      reinterpret_cast<const float*>(ws), splits, (int)C, groups, rows_per_inst, eps, mean_rstd, total);
  PB_LAUNCH_CHECK("gn_finalize_kernel");
  return PB_OK;
}

int pb_groupnorm_apply(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype,
                       int64_t ldy, int64_t n_inst, int64_t rows_per_inst, int64_t C,
                       int32_t groups, const float* mean_rstd, const float* gamma,
                       const float* beta, int32_t silu, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && mean_rstd && gamma && beta, "pb_groupnorm_apply: null pointer");
  PB_REQUIRE(n_inst > 0 && rows_per_inst > 0 && C > 0 && groups > 0, "pb_groupnorm_apply: empty input");
  PB_REQUIRE(C % 4 == 0 && C % groups == 0 && ldx % 4 == 0 && ldy % 4 == 0,
             "pb_groupnorm_apply: need C %% 4 == 0, C %% groups == 0, ld %% 4 == 0");
  const int64_t total_vec = n_inst * rows_per_inst * (C / 4);
  int64_t blocks = ceil_div(total_vec, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(x_dtype, TX, {
    PB_DISPATCH_DTYPE(y_dtype, TY, {
      gn_apply_kernel<TX, TY><<<(unsigned)blocks, 256, 0, st>>>(
          reinterpret_cast<const TX*>(x), ldx, reinterpret_cast<TY*>(y), ldy, rows_per_inst, (int)C,
          groups, mean_rstd, gamma, beta, silu, total_vec);
    })
  })
  PB_LAUNCH_CHECK("gn_apply_kernel");
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
