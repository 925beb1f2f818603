// groupnorm.cu — GroupNorm(32)(+SiLU) as HBM-streaming kernels.
//
// (basics.py:76-88; openaimodel3d.py:155,179,259-268,547; attention.py:265,331)
//   rows layout [n_inst * R, C]; statistics per (instance, group) over R rows x C/G channels.
//   Pass 1  gn_partial_kernel : grid (splits, n_inst); every thread owns one 4-channel vector column and walks
//           rows with a fixed stride (4 loads in flight): coalesced, register accumulators; the CTA folds its
//           channels into per-group (sum, sumsq) and writes [n_inst, splits, G, 2].
//   Pass 2  gn_finalize_kernel: one CTA per (instance, group) folds the splits in fp64 -> mean, rstd ->
//           per-channel (scale, shift).
//   Pass 3  gn_apply_kernel   : y = x*scale + shift (+SiLU), vectorised, no integer division in the loop,
//           written in the operand dtype the next GEMM wants (bf16 or fp32).
#include <string.h>

#include "common.cuh"

namespace pb {

constexpr int GN_THREADS = 256;
// apply pass: rows per CTA = 8 * TY (each thread streams 8 rows of its column, 4 loads in flight)

// Thread mapping shared by both passes: TX threads across 4-channel vector columns, TY row lanes.
// A thread keeps its column(s) for the whole CTA lifetime, so per-column data sits in registers and the
// inner loops contain no integer division.
static inline void gn_shape(int cv, int* TX, int* TY) {
  *TX = cv < GN_THREADS ? cv : GN_THREADS;
  *TY = GN_THREADS / *TX;
}

static inline int gn_splits(int64_t n_inst, int64_t R) {
  // ~8 CTAs per SM overall, at least 8 rows per CTA
  int64_t want = ceil_div(148 * 8, n_inst);
  int64_t maxs = ceil_div(R, 8);
  int64_t s = want < 1 ? 1 : want;
  if (s > maxs) s = maxs;
  if (s > 2048) s = 2048;
  return (int)s;
}

constexpr int GN_MAX_C = 2560;

// slab-range splits of the column-statistics fold: aim at ~4 CTAs per SM, at least 16 slabs per CTA
static inline int gn_cols_splits(int64_t n_inst, int64_t R, int groups) {
  const int64_t ctas = n_inst * groups, slabs = R >> 5;
  int64_t s = ceil_div(148 * 4, ctas);
  if (s > slabs / 16) s = slabs / 16;
  return s < 1 ? 1 : (int)s;
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_partial_kernel(const T* __restrict__ x, int64_t ldx, int64_t R, int C, int G, int splits, int TX, int TY,
                  float* __restrict__ ws /* [n_inst, splits, G, 2] */) {
  pb::pdl_trigger();
  __shared__ float red[GN_THREADS * 8];
  __shared__ float chan[GN_MAX_C * 2];           // per-channel (sum, sumsq) of this CTA's rows
  const int inst = blockIdx.y, sp = blockIdx.x;
  const int cv = C >> 2;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int64_t r_begin = R * sp / splits, r_end = R * (sp + 1) / splits;
  const T* base = x + (int64_t)inst * R * ldx;
  for (int v0 = 0; v0 < cv; v0 += TX) {
    const int v = v0 + tx;
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (v < cv && ty < TY) {
      const T* col = base + (int64_t)v * 4;
      int64_t r = r_begin + ty;
      for (; r + 3 * TY < r_end; r += 4 * TY) {       // 4 independent loads in flight
        float t0[4], t1[4], t2[4], t3[4];
        load4(col + r * ldx, t0);
        load4(col + (r + TY) * ldx, t1);
        load4(col + (r + 2 * TY) * ldx, t2);
        load4(col + (r + 3 * TY) * ldx, t3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[i] += (t0[i] + t1[i]) + (t2[i] + t3[i]);
          q[i] += (t0[i] * t0[i] + t1[i] * t1[i]) + (t2[i] * t2[i] + t3[i] * t3[i]);
        }
      }
      for (; r < r_end; r += TY) {
        float t[4];
        load4(col + r * ldx, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i] += t[i]; q[i] = fmaf(t[i], t[i], q[i]); }
      }
    }
    if (TY > 1) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) { red[threadIdx.x * 8 + i] = s[i]; red[threadIdx.x * 8 + 4 + i] = q[i]; }
      __syncthreads();
      if (ty == 0 && v < cv) {
        for (int l = 1; l < TY; ++l) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { s[i] += red[(l * TX + tx) * 8 + i]; q[i] += red[(l * TX + tx) * 8 + 4 + i]; }
        }
      }
    }
    if (ty == 0 && v < cv) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { chan[(v * 4 + i) * 2] = s[i]; chan[(v * 4 + i) * 2 + 1] = q[i]; }
    }
  }
  __syncthreads();
  // channels -> groups (one thread per (group, stat))
  if (threadIdx.x < 2 * G) {
    const int g = threadIdx.x >> 1, st = threadIdx.x & 1, cpg = C / G;
    float a = 0.f;
    for (int c = 0; c < cpg; ++c) a += chan[(g * cpg + c) * 2 + st];
    ws[(((int64_t)inst * splits + sp) * G + g) * 2 + st] = a;
  }
}

// One 128-thread CTA per (instance, group): fold the splits in fp64 -> mean, rstd, then emit the per-channel
// affine  y = x * scale + shift  (scale = rstd*gamma, shift = beta - mean*rstd*gamma).
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float* __restrict__ ws, int splits, int C, int G, int64_t R, float eps,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ scale_shift /* [n_inst, 2, C] */) {
  pb::pdl_trigger();
  __shared__ double red[2][4];
  __shared__ float stat[2];
  const int inst = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const float* base = ws + ((int64_t)inst * splits * G + g) * 2;
  double s = 0.0, q = 0.0;
  for (int sp = threadIdx.x; sp < splits; sp += 128) {
    const float2 v = *reinterpret_cast<const float2*>(base + (int64_t)sp * G * 2);
    s += (double)v.x; q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double ss = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    const double qq = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const double cnt = (double)R * cpg;
    const double mean = ss / cnt;
    double var = qq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float meanf = stat[0], rstd = stat[1];
  float* sc = scale_shift + (int64_t)inst * 2 * C;
  for (int i = threadIdx.x; i < cpg; i += 128) {
    const int c = g * cpg + i;
    const float a = rstd * gamma[c];
    sc[c] = a;
    sc[C + c] = beta[c] - meanf * a;
  }
}

// Fold of the column statistics written by the producing GEMM's epilogue (pb_gemm_args.colstats): one CTA per
// (instance, group) sums (R/32 slabs) x (C/G columns) float2 entries in fp64, then emits the per-channel affine.
__global__ void __launch_bounds__(256)
gn_finalize_cols_kernel(const float* __restrict__ cs, int64_t ld_cs, int C, int G, int64_t R, float eps,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* __restrict__ scale_shift /* [n_inst, 2, C] */) {
  pb::pdl_trigger();
  __shared__ double red[2][8];
  __shared__ float stat[2];
  const int inst = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  const int64_t slabs = R >> 5;
  const float* base = cs + (int64_t)inst * slabs * ld_cs + 2 * (int64_t)g * cpg;
  const int64_t total = slabs * cpg;
  double s = 0.0, q = 0.0;
  for (int64_t i = threadIdx.x; i < total; i += 256) {
    const int64_t sl = i / cpg;
    const int c = (int)(i - sl * cpg);
    const float2 v = *reinterpret_cast<const float2*>(base + sl * ld_cs + 2 * c);
    s += (double)v.x; q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0, qq = 0.0;
    for (int i = 0; i < 8; ++i) { ss += red[0][i]; qq += red[1][i]; }
    const double cnt = (double)R * cpg;
    const double mean = ss / cnt;
    double var = qq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float meanf = stat[0], rstd = stat[1];
  float* sc = scale_shift + (int64_t)inst * 2 * C;
  for (int i = threadIdx.x; i < cpg; i += 256) {
    const int c = g * cpg + i;
    const float a = rstd * gamma[c];
    sc[c] = a;
    sc[C + c] = beta[c] - meanf * a;
  }
}

// Few (instance, group) pairs but many slabs (clip-wide norms: 2 x 32 CTAs would each walk thousands of slabs):
// grid (n_inst * G, splits) partial fp64 sums -> ws[(inst*G + g) * splits + sp]; gn_finalize_f64_kernel folds them.
__global__ void __launch_bounds__(256)
gn_cols_partial_kernel(const float* __restrict__ cs, int64_t ld_cs, int C, int G, int64_t R, int splits,
                       double* __restrict__ ws) {
  pb::pdl_trigger();
  __shared__ double red[2][8];
  const int inst = blockIdx.x / G, g = blockIdx.x % G, sp = blockIdx.y;
  const int cpg = C / G;
  const int64_t slabs = R >> 5;
  const int64_t s0 = slabs * sp / splits, s1 = slabs * (sp + 1) / splits;
  const float* base = cs + ((int64_t)inst * slabs + s0) * ld_cs + 2 * (int64_t)g * cpg;
  const int total = (int)(s1 - s0) * cpg;          // < 2^31: slabs of one instance x channels of one group
  double s = 0.0, q = 0.0;
  // 32-bit indexing and four independent loads in flight per thread: the kernel moves a few MB and is pure latency
  // (the first version did a 64-bit division per element and had one load outstanding)
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * 256) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i < total) {
        const int sl = i / cpg, c = i - sl * cpg;
        v[u] = *reinterpret_cast<const float2*>(base + (int64_t)sl * ld_cs + 2 * c);
      } else {
        v[u] = make_float2(0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { s += (double)v[u].x; q += (double)v[u].y; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0, qq = 0.0;
    for (int i = 0; i < 8; ++i) { ss += red[0][i]; qq += red[1][i]; }
    ws[((int64_t)blockIdx.x * splits + sp) * 2] = ss;
    ws[((int64_t)blockIdx.x * splits + sp) * 2 + 1] = qq;
  }
}

__global__ void __launch_bounds__(128)
gn_finalize_f64_kernel(const double* __restrict__ ws, int splits, int C, int G, int64_t R, float eps,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       float* __restrict__ scale_shift) {
  pb::pdl_trigger();
  __shared__ float stat[2];
  const int inst = blockIdx.x / G, g = blockIdx.x % G;
  const int cpg = C / G;
  if (threadIdx.x == 0) {
    double ss = 0.0, qq = 0.0;
    for (int i = 0; i < splits; ++i) { ss += ws[((int64_t)blockIdx.x * splits + i) * 2]; qq += ws[((int64_t)blockIdx.x * splits + i) * 2 + 1]; }
    const double cnt = (double)R * cpg;
    const double mean = ss / cnt;
    double var = qq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float meanf = stat[0], rstd = stat[1];
  float* sc = scale_shift + (int64_t)inst * 2 * C;
  for (int i = threadIdx.x; i < cpg; i += 128) {
    const int c = g * cpg + i;
    const float a = rstd * gamma[c];
    sc[c] = a;
    sc[C + c] = beta[c] - meanf * a;
  }
}

// Sequence-parallel variant of the fold: partials -> fp64 (sum, sumsq) per (instance, group), to be
// all-reduced across ranks by the caller, then turned into the per-channel affine with the GLOBAL count.
__global__ void __launch_bounds__(128)
gn_sums_kernel(const float* __restrict__ ws, int splits, int G, double* __restrict__ sums /* [n_inst, G, 2] */) {
  pb::pdl_trigger();
  __shared__ double red[2][4];
  const int inst = blockIdx.x / G, g = blockIdx.x % G;
  const float* base = ws + ((int64_t)inst * splits * G + g) * 2;
  double s = 0.0, q = 0.0;
  for (int sp = threadIdx.x; sp < splits; sp += 128) {
    const float2 v = *reinterpret_cast<const float2*>(base + (int64_t)sp * G * 2);
    s += (double)v.x; q += (double)v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = s; red[1][w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    sums[(int64_t)blockIdx.x * 2] = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    sums[(int64_t)blockIdx.x * 2 + 1] = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  }
}

__global__ void gn_affine_kernel(const double* __restrict__ sums, int C, int G, double count, float eps,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ scale_shift, int total /* n_inst * C */) {
  pb::pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int inst = i / C, c = i - inst * C, g = c / (C / G);
  const double mean = sums[((int64_t)inst * G + g) * 2] / count;
  double var = sums[((int64_t)inst * G + g) * 2 + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float a = rstd * gamma[c];
  scale_shift[(int64_t)inst * 2 * C + c] = a;
  scale_shift[(int64_t)inst * 2 * C + C + c] = beta[c] - (float)mean * a;
}

// Folded-finalize variant (part != nullptr): the per-(instance, group) fp64 partial sums of gn_cols_partial_kernel are
// folded by every CTA in its prologue (G threads x `splits` loads), so no separate finalize launch and no
// scale/shift buffer exist; the affine is built in registers from gamma / beta.
struct GnFold {
  const double* part;     // [(inst * G + g) * splits + sp][2]
  const float* gamma;
  const float* beta;
  int splits, G;
  double count;           // elements per (instance, group)
  float eps;
};

// y * sigmoid(y).  The quarter-rate XU pipe (MUFU + conversions) is what a streaming norm+SiLU pass runs out of first
// (2 MUFU + the bf16 conversion per element is 60 % of that pipe at the HBM rate), so:
//   bf16 output: sigmoid(y) = 0.5 + 0.5 tanh(y / 2) with ONE MUFU (tanh.approx, max relative error 2^-11, an eighth of
//                the bf16 rounding that follows): h = y/2, h + h * tanh(h); the conversion is done on the integer pipe;
//   fp32 output: ex2 + rcp (~2 ulp), the exact mode's budget.
template <bool LOWP>
__device__ __forceinline__ float silu_fast(float y) {
  if (LOWP) {
    const float h = 0.5f * y;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
  }
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return y * r;
}
// four floats -> four bf16 (round to nearest even) on the integer pipe; finite inputs
__device__ __forceinline__ void store4_bf16_alu(bf16* p, const float (&v)[4]) {
  uint32_t u[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t x = __float_as_uint(v[j]);
    u[j] = x + 0x7FFFu + ((x >> 16) & 1u);
  }
  uint2 t;
  t.x = __byte_perm(u[0], u[1], 0x7632);
  t.y = __byte_perm(u[2], u[3], 0x7632);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ void store4_fast(bf16* p, const float (&v)[4]) { store4_bf16_alu(p, v); }
__device__ __forceinline__ void store4_fast(float* p, const float (&v)[4]) { store4(p, v); }

// Apply pass.  A CTA owns `rows_per_cta` consecutive rows of one instance and walks them as ONE flat stream of 4-channel
// vectors (thread t takes vectors t, t + 256, ...: every thread busy for any channel count, neighbouring threads on
// neighbouring 16-byte words), 4 loads in flight per thread.  The per-channel affine (scale, shift) is built once per CTA
// in shared memory — from the folded fp64 partials when `fold.part` is set — and read back as float4 pairs; the row /
// column of a vector is advanced incrementally (no integer division in the loop).  CTAs are sized by the host so that a
// launch is a few waves of fat CTAs: the earlier version's 24-row CTAs spent as long in their prologue (fold, sqrt,
// gamma / beta loads) as in streaming their 30 KB, and ran at 47 % of the DRAM rate.
template <typename TX_, typename TY_>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const TX_* __restrict__ x, int64_t ldx, TY_* __restrict__ y, int64_t ldy, int64_t R, int C,
                const float* __restrict__ scale_shift, int silu, int rows_per_cta, GnFold fold,
                TY_* __restrict__ xc /* optional: plain copy of x in y's dtype, row pitch ldc */, int64_t ldc) {
  pb::pdl_trigger();
  extern __shared__ float4 gn_ab[];               // [cv] scale vectors, then [cv] shift vectors
  __shared__ float g_mean[32], g_rstd[32];
  const int inst = blockIdx.y;
  const int cv = C >> 2;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int nrows = (int)(r_begin + rows_per_cta < R ? rows_per_cta : R - r_begin);
  const TX_* xb = x + ((int64_t)inst * R + r_begin) * ldx;
  TY_* yb = y + ((int64_t)inst * R + r_begin) * ldy;
  TY_* cb = xc == nullptr ? nullptr : xc + ((int64_t)inst * R + r_begin) * ldc;
  if (fold.part != nullptr) {
    if (threadIdx.x < fold.G) {
      const double* pp = fold.part + ((int64_t)inst * fold.G + threadIdx.x) * fold.splits * 2;
      double ss = 0.0, qq = 0.0;
      for (int i = 0; i < fold.splits; ++i) { ss += pp[2 * i]; qq += pp[2 * i + 1]; }
      const double mean = ss / fold.count;
      double var = qq / fold.count - mean * mean;
      if (var < 0.0) var = 0.0;
      g_mean[threadIdx.x] = (float)mean;
      g_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)fold.eps));
    }
    __syncthreads();
    const int cpg = C / fold.G;
    for (int v = threadIdx.x; v < cv; v += GN_THREADS) {
      float a[4], b[4], gm[4], bt[4];
      load4(fold.gamma + v * 4, gm);
      load4(fold.beta + v * 4, bt);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (v * 4 + j) / cpg;
        a[j] = g_rstd[g] * gm[j];
        b[j] = bt[j] - g_mean[g] * a[j];
      }
      gn_ab[v] = make_float4(a[0], a[1], a[2], a[3]);
      gn_ab[cv + v] = make_float4(b[0], b[1], b[2], b[3]);
    }
  } else {
    const float* sc = scale_shift + (int64_t)inst * 2 * C;
    for (int v = threadIdx.x; v < cv; v += GN_THREADS) {
      gn_ab[v] = *reinterpret_cast<const float4*>(sc + v * 4);
      gn_ab[cv + v] = *reinterpret_cast<const float4*>(sc + C + v * 4);
    }
  }
  __syncthreads();
  // Flat walk over nrows x cv vectors.  Everything inside the CTA is 32-bit: vector offsets relative to the CTA's first
  // row (the host bounds rows_per_cta so that they fit), advanced incrementally — one step is 256 vectors = (drow rows,
  // dcol columns), with a wrap into the next row.
  const int total = nrows * cv;
  const int ldxv = (int)(ldx >> 2), ldyv = (int)(ldy >> 2);
  const int drow = GN_THREADS / cv, dcol = GN_THREADS - drow * cv;
  const int ldcv = (int)(ldc >> 2);
  const int stepx = drow * ldxv + dcol, stepy = drow * ldyv + dcol, wrapx = ldxv - cv, wrapy = ldyv - cv;
  const int stepc = drow * ldcv + dcol, wrapc = ldcv - cv;
  int col = (int)threadIdx.x % cv;
  int offx = ((int)threadIdx.x / cv) * ldxv + col, offy = ((int)threadIdx.x / cv) * ldyv + col;
  int offc = ((int)threadIdx.x / cv) * ldcv + col;
  auto advance = [&]() {
    col += dcol; offx += stepx; offy += stepy; offc += stepc;
    if (col >= cv) { col -= cv; offx += wrapx; offy += wrapy; offc += wrapc; }
  };
  auto emit = [&](const float (&t)[4], int c, int oy, int oc) {
    if (cb != nullptr) store4_fast(cb + (int64_t)oc * 4, t);
    const float4 a = gn_ab[c], b = gn_ab[cv + c];
    float o[4];
    o[0] = fmaf(t[0], a.x, b.x); o[1] = fmaf(t[1], a.y, b.y);
    o[2] = fmaf(t[2], a.z, b.z); o[3] = fmaf(t[3], a.w, b.w);
    if (silu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = silu_fast<sizeof(TY_) == 2>(o[j]);
    }
    store4_fast(yb + (int64_t)oy * 4, o);
  };
  int i = threadIdx.x;
  for (; i + 3 * GN_THREADS < total; i += 4 * GN_THREADS) {      // 4 independent 16-byte loads in flight per thread
    float t[4][4];
    int cc[4], oy[4], oc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      cc[u] = col; oy[u] = offy; oc[u] = offc;
      load4(xb + (int64_t)offx * 4, t[u]);
      advance();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(t[u], cc[u], oy[u], oc[u]);
  }
  for (; i < total; i += GN_THREADS) {
    float t[4];
    load4(xb + (int64_t)offx * 4, t);
    emit(t, col, offy, offc);
    advance();
  }
}

// One launch for SMALL instances (the H/8 and H/4 levels, where the producer left no column statistics): a CLUSTER of
// GN_CLUSTER CTAs owns one (instance, group) — R rows x C/G channels, a few KB to ~100 KB.  Each CTA sums its share of
// the rows, the partial sums meet through distributed shared memory (every CTA folds them in rank order, so all get
// the same bits), and the CTA streams its rows a second time (now an L1 / L2 hit) to apply the affine (+SiLU).
// Replaces partial + finalize + apply (three launches of ~7 us each on tensors that take ~2 us to move).  The first
// version ran ONE CTA per (instance, group) with one load in flight per thread: 64 CTAs on 148 SMs, 25 dependent
// round trips per pass, 19 us for 6.6 MB.  Needs C/G % 4 == 0.
constexpr int GN_CLUSTER = 4;

__device__ __forceinline__ uint32_t gn_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void gn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double gn_ld_dsmem_f64(const void* local_smem, uint32_t rank) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(local_smem)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}

template <typename TX_, typename TY_, int CS>       // CS = CTAs per (instance, group): GN_CLUSTER, or 1 for short instances
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(GN_THREADS)
gn_group_kernel(const TX_* __restrict__ x, int64_t ldx, TY_* __restrict__ y, int64_t ldy, int R, int C, int G, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu) {
  pb::pdl_trigger();
  __shared__ double red[2][GN_THREADS / 32];
  __shared__ double part[2];
  __shared__ float stat[2];
  const int inst = blockIdx.y, g = blockIdx.x / CS;
  const int rank = CS > 1 ? (int)gn_cluster_rank() : 0;
  const int cpg = C / G, vpg = cpg >> 2;
  const int rows_cta = (R + CS - 1) / CS;
  const int r0 = rank * rows_cta, r1 = (r0 + rows_cta < R) ? r0 + rows_cta : R;
  const TX_* xb = x + ((int64_t)inst * R + r0) * ldx + g * cpg;
  TY_* yb = y + ((int64_t)inst * R + r0) * ldy + g * cpg;
  const int total = (r1 > r0 ? r1 - r0 : 0) * vpg;
  // vector i of this CTA's slice -> (row, column vector); 4 loads in flight per thread
  float s = 0.f, q = 0.f;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * GN_THREADS) {
    float t[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * GN_THREADS;
      if (i < total) {
        const int r = i / vpg, c = i - r * vpg;
        load4(xb + (int64_t)r * ldx + c * 4, t[u]);
      } else {
        t[u][0] = t[u][1] = t[u][2] = t[u][3] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s += (t[u][0] + t[u][1]) + (t[u][2] + t[u][3]);
      q += (t[u][0] * t[u][0] + t[u][1] * t[u][1]) + (t[u][2] * t[u][2] + t[u][3] * t[u][3]);
    }
  }
  double sd = (double)s, qd = (double)q;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
    qd += __shfl_xor_sync(0xffffffffu, qd, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = sd; red[1][w] = qd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ss = 0.0, qq = 0.0;
    for (int i = 0; i < GN_THREADS / 32; ++i) { ss += red[0][i]; qq += red[1][i]; }
    part[0] = ss; part[1] = qq;
  }
  if (CS > 1) gn_cluster_sync(); else __syncthreads();      // every CTA's partial is visible cluster-wide
  if (threadIdx.x == 0) {
    double ss = 0.0, qq = 0.0;
    if (CS > 1) {
      for (int k = 0; k < CS; ++k) { ss += gn_ld_dsmem_f64(&part[0], k); qq += gn_ld_dsmem_f64(&part[1], k); }
    } else {
      ss = part[0]; qq = part[1];
    }
    const double cnt = (double)R * cpg;
    const double mean = ss / cnt;
    double var = qq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  if (CS > 1) gn_cluster_sync(); else __syncthreads();      // stat is set; nobody exits while a peer may still read its partial
  const float mean = stat[0], rstd = stat[1];
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * GN_THREADS) {
    float t[4][4];
    int rr[4], cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * GN_THREADS;
      rr[u] = i / vpg; cc[u] = i - rr[u] * vpg;
      if (i < total) load4(xb + (int64_t)rr[u] * ldx + cc[u] * 4, t[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * GN_THREADS < total) {
        float gm[4], bt[4], o[4];
        load4(gamma + g * cpg + cc[u] * 4, gm);
        load4(beta + g * cpg + cc[u] * 4, bt);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = rstd * gm[j];
          o[j] = fmaf(t[u][j], a, bt[j] - mean * a);
          if (silu) o[j] = silu_fast<sizeof(TY_) == 2>(o[j]);
        }
        store4_fast(yb + (int64_t)rr[u] * ldy + cc[u] * 4, o);
      }
    }
  }
}

// rows per CTA of the apply pass: ~6 CTAs per SM over the whole launch, at least 16 rows, at most what int32 indexing holds
static inline int gn_apply_rows(int64_t n_inst, int64_t R, int64_t C, int64_t ld_max) {
  int64_t rows = ceil_div(n_inst * R, (int64_t)148 * 6);
  if (rows < 16) rows = 16;
  if (rows > R) rows = R;
  const int64_t cap = ((int64_t)1 << 30) / (ld_max / 4 + 1);      // 32-bit vector offsets inside a CTA
  if (rows > cap) rows = cap;
  (void)C;
  return (int)rows;
}

}  // namespace pb

extern "C" {

size_t pb_groupnorm_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int64_t C) {
  if (n_inst <= 0 || rows_per_inst <= 0 || C <= 0) return 0;
  return (size_t)n_inst * pb::gn_splits(n_inst, rows_per_inst) * 32 * 2 * sizeof(float);
}

int pb_groupnorm_stats(const void* x, int32_t x_dtype, int64_t ldx, int64_t n_inst,
                       int64_t rows_per_inst, int64_t C, int32_t groups, float eps, const float* gamma,
                       const float* beta, float* scale_shift, void* ws, size_t ws_bytes, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && scale_shift && ws && gamma && beta, "pb_groupnorm_stats: null pointer");
  PB_REQUIRE(n_inst > 0 && rows_per_inst > 0 && C > 0 && groups > 0, "pb_groupnorm_stats: empty input");
  PB_REQUIRE(C % 4 == 0 && C % groups == 0 && ldx % 4 == 0,
             "pb_groupnorm_stats: need C %% 4 == 0, C %% groups == 0, ldx %% 4 == 0 (C=%lld)", (long long)C);
  PB_REQUIRE(n_inst <= 65535, "pb_groupnorm_stats: n_inst too large");
  PB_REQUIRE(C <= GN_MAX_C && groups <= 32, "pb_groupnorm_stats: C <= %d and groups <= 32 supported", GN_MAX_C);
  const int splits = gn_splits(n_inst, rows_per_inst);
  PB_REQUIRE(ws_bytes >= pb_groupnorm_ws_bytes(n_inst, rows_per_inst, C),
             "pb_groupnorm_stats: workspace too small");
  cudaStream_t st = to_stream(stream);
  int TX, TY;
  gn_shape((int)(C / 4), &TX, &TY);
  dim3 grid(splits, (unsigned)n_inst);
  PB_DISPATCH_DTYPE(x_dtype, T, {
    gn_partial_kernel<T><<<grid, GN_THREADS, 0, st>>>(reinterpret_cast<const T*>(x), ldx, rows_per_inst, (int)C,
                                                      groups, splits, TX, TY, reinterpret_cast<float*>(ws));
  })
  PB_LAUNCH_CHECK("gn_partial_kernel");
  gn_finalize_kernel<<<(unsigned)(n_inst * groups), 128, 0, st>>>(
      reinterpret_cast<const float*>(ws), splits, (int)C, groups, rows_per_inst, eps, gamma, beta, scale_shift);
  PB_LAUNCH_CHECK("gn_finalize_kernel");
  return PB_OK;
}

size_t pb_groupnorm_cols_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int32_t groups) {
  if (n_inst <= 0 || rows_per_inst <= 0 || groups <= 0) return 0;
  const int s = pb::gn_cols_splits(n_inst, rows_per_inst, groups);
  return s > 1 ? (size_t)n_inst * groups * s * 2 * sizeof(double) : 0;
}

int pb_groupnorm_stats_cols(const float* colstats, int64_t ld_colstats, int64_t n_inst, int64_t rows_per_inst,
                            int64_t C, int32_t groups, float eps, const float* gamma, const float* beta,
                            float* scale_shift, void* ws, size_t ws_bytes, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(colstats && scale_shift && gamma && beta, "pb_groupnorm_stats_cols: null pointer");
  PB_REQUIRE(n_inst > 0 && rows_per_inst > 0 && C > 0 && groups > 0 && C % groups == 0,
             "pb_groupnorm_stats_cols: bad sizes");
  PB_REQUIRE(rows_per_inst % 32 == 0, "pb_groupnorm_stats_cols: rows_per_inst must be a multiple of 32 (the slab)");
  PB_REQUIRE(ld_colstats >= 2 * C && ld_colstats % 2 == 0, "pb_groupnorm_stats_cols: bad ld_colstats");
  const int splits = gn_cols_splits(n_inst, rows_per_inst, groups);
  if (splits > 1) {
    PB_REQUIRE(ws && ws_bytes >= pb_groupnorm_cols_ws_bytes(n_inst, rows_per_inst, groups),
               "pb_groupnorm_stats_cols: workspace too small");
    dim3 grid((unsigned)(n_inst * groups), (unsigned)splits);
    gn_cols_partial_kernel<<<grid, 256, 0, to_stream(stream)>>>(colstats, ld_colstats, (int)C, groups, rows_per_inst,
                                                                splits, reinterpret_cast<double*>(ws));
    PB_LAUNCH_CHECK("gn_cols_partial_kernel");
    gn_finalize_f64_kernel<<<(unsigned)(n_inst * groups), 128, 0, to_stream(stream)>>>(
        reinterpret_cast<const double*>(ws), splits, (int)C, groups, rows_per_inst, eps, gamma, beta, scale_shift);
    PB_LAUNCH_CHECK("gn_finalize_f64_kernel");
    return PB_OK;
  }
  gn_finalize_cols_kernel<<<(unsigned)(n_inst * groups), 256, 0, to_stream(stream)>>>(
      colstats, ld_colstats, (int)C, groups, rows_per_inst, eps, gamma, beta, scale_shift);
  PB_LAUNCH_CHECK("gn_finalize_cols_kernel");
  return PB_OK;
}

int pb_groupnorm_apply_cols(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                            const float* colstats, int64_t ld_colstats, int64_t n_inst, int64_t rows_per_inst,
                            int64_t C, int32_t groups, float eps, const float* gamma, const float* beta, int32_t silu,
                            void* ws, size_t ws_bytes, void* x_copy, int64_t ld_copy, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && colstats && gamma && beta && ws, "pb_groupnorm_apply_cols: null pointer");
  PB_REQUIRE(n_inst > 0 && n_inst <= 65535 && rows_per_inst > 0 && C > 0 && groups > 0 && groups <= 32 &&
                 C % groups == 0, "pb_groupnorm_apply_cols: bad sizes");
  PB_REQUIRE(rows_per_inst % 32 == 0, "pb_groupnorm_apply_cols: rows_per_inst must be a multiple of 32 (the slab)");
  PB_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "pb_groupnorm_apply_cols: need C %% 4 == 0, ld %% 4 == 0");
  PB_REQUIRE(ld_colstats >= 2 * C && ld_colstats % 2 == 0, "pb_groupnorm_apply_cols: bad ld_colstats");
  const int splits = gn_cols_splits(n_inst, rows_per_inst, groups);
  PB_REQUIRE(ws_bytes >= (size_t)n_inst * groups * splits * 2 * sizeof(double),
             "pb_groupnorm_apply_cols: workspace too small");
  cudaStream_t st = to_stream(stream);
  dim3 pgrid((unsigned)(n_inst * groups), (unsigned)splits);
  gn_cols_partial_kernel<<<pgrid, 256, 0, st>>>(colstats, ld_colstats, (int)C, groups, rows_per_inst, splits,
                                                reinterpret_cast<double*>(ws));
  PB_LAUNCH_CHECK("gn_cols_partial_kernel");
  const int rows_per_cta = gn_apply_rows(n_inst, rows_per_inst, C, (ldx > ldy ? ldx : ldy) > ld_copy ? (ldx > ldy ? ldx : ldy) : ld_copy);
  dim3 grid((unsigned)ceil_div(rows_per_inst, rows_per_cta), (unsigned)n_inst);
  const size_t smem = (size_t)C * 2 * sizeof(float);
  GnFold fold;
  fold.part = reinterpret_cast<const double*>(ws); fold.gamma = gamma; fold.beta = beta;
  fold.splits = splits; fold.G = groups; fold.count = (double)rows_per_inst * (double)(C / groups); fold.eps = eps;
  PB_DISPATCH_DTYPE(x_dtype, TXT, {
    PB_DISPATCH_DTYPE(y_dtype, TYT, {
      gn_apply_kernel<TXT, TYT><<<grid, GN_THREADS, smem, st>>>(reinterpret_cast<const TXT*>(x), ldx,
                                                                reinterpret_cast<TYT*>(y), ldy, rows_per_inst, (int)C,
                                                                nullptr, silu, rows_per_cta, fold,
                                                                reinterpret_cast<TYT*>(x_copy), ld_copy);
    })
  })
  PB_LAUNCH_CHECK("gn_apply_kernel");
  return PB_OK;
}

size_t pb_groupnorm_apply_cols_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int32_t groups) {
  if (n_inst <= 0 || rows_per_inst <= 0 || groups <= 0) return 0;
  return (size_t)n_inst * groups * pb::gn_cols_splits(n_inst, rows_per_inst, groups) * 2 * sizeof(double);
}

int pb_groupnorm_sums(const void* x, int32_t x_dtype, int64_t ldx, int64_t n_inst, int64_t rows_per_inst,
                      int64_t C, int32_t groups, double* sums, void* ws, size_t ws_bytes, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && sums && ws, "pb_groupnorm_sums: null pointer");
  PB_REQUIRE(n_inst > 0 && n_inst <= 65535 && rows_per_inst > 0 && C > 0 && groups > 0 && groups <= 32,
             "pb_groupnorm_sums: bad sizes");
  PB_REQUIRE(C % 4 == 0 && C % groups == 0 && ldx % 4 == 0 && C <= GN_MAX_C, "pb_groupnorm_sums: bad channel count");
  const int splits = gn_splits(n_inst, rows_per_inst);
  PB_REQUIRE(ws_bytes >= pb_groupnorm_ws_bytes(n_inst, rows_per_inst, C), "pb_groupnorm_sums: workspace too small");
  cudaStream_t st = to_stream(stream);
  int TX, TY;
  gn_shape((int)(C / 4), &TX, &TY);
  dim3 grid(splits, (unsigned)n_inst);
  PB_DISPATCH_DTYPE(x_dtype, T, {
    gn_partial_kernel<T><<<grid, GN_THREADS, 0, st>>>(reinterpret_cast<const T*>(x), ldx, rows_per_inst, (int)C,
                                                      groups, splits, TX, TY, reinterpret_cast<float*>(ws));
  })
  PB_LAUNCH_CHECK("gn_partial_kernel");
  gn_sums_kernel<<<(unsigned)(n_inst * groups), 128, 0, st>>>(reinterpret_cast<const float*>(ws), splits, groups, sums);
  PB_LAUNCH_CHECK("gn_sums_kernel");
  return PB_OK;
}

int pb_groupnorm_affine(const double* sums, int64_t n_inst, int64_t C, int32_t groups, double count, float eps,
                        const float* gamma, const float* beta, float* scale_shift, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(sums && gamma && beta && scale_shift && n_inst > 0 && C > 0 && groups > 0 && count > 0,
             "pb_groupnorm_affine: bad arguments");
  const int total = (int)(n_inst * C);
  gn_affine_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, to_stream(stream)>>>(sums, (int)C, groups, count, eps,
                                                                                 gamma, beta, scale_shift, total);
  PB_LAUNCH_CHECK("gn_affine_kernel");
  return PB_OK;
}

int pb_groupnorm_apply(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype,
                       int64_t ldy, int64_t n_inst, int64_t rows_per_inst, int64_t C,
                       const float* scale_shift, int32_t silu, void* x_copy, int64_t ld_copy, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && scale_shift, "pb_groupnorm_apply: null pointer");
  PB_REQUIRE(n_inst > 0 && n_inst <= 65535 && rows_per_inst > 0 && C > 0, "pb_groupnorm_apply: bad sizes");
  PB_REQUIRE(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "pb_groupnorm_apply: need C %% 4 == 0, ld %% 4 == 0");
  PB_REQUIRE(C <= GN_MAX_C, "pb_groupnorm_apply: C <= %d supported", GN_MAX_C);
  cudaStream_t st = to_stream(stream);
  const int rows_per_cta = gn_apply_rows(n_inst, rows_per_inst, C, (ldx > ldy ? ldx : ldy) > ld_copy ? (ldx > ldy ? ldx : ldy) : ld_copy);
  dim3 grid((unsigned)ceil_div(rows_per_inst, rows_per_cta), (unsigned)n_inst);
  const size_t smem = (size_t)C * 2 * sizeof(float);
  PB_DISPATCH_DTYPE(x_dtype, TXT, {
    PB_DISPATCH_DTYPE(y_dtype, TYT, {
      GnFold nofold;
      memset(&nofold, 0, sizeof(nofold));
      gn_apply_kernel<TXT, TYT><<<grid, GN_THREADS, smem, st>>>(reinterpret_cast<const TXT*>(x), ldx,
                                                                reinterpret_cast<TYT*>(y), ldy, rows_per_inst, (int)C,
                                                                scale_shift, silu, rows_per_cta, nofold,
                                                                reinterpret_cast<TYT*>(x_copy), ld_copy);
    })
  })
  PB_LAUNCH_CHECK("gn_apply_kernel");
  return PB_OK;
}

int pb_groupnorm_small(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                       int64_t n_inst, int64_t rows_per_inst, int64_t C, int32_t groups, float eps, const float* gamma,
                       const float* beta, int32_t silu, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && gamma && beta, "pb_groupnorm_small: null pointer");
  PB_REQUIRE(n_inst > 0 && n_inst <= 65535 && rows_per_inst > 0 && C > 0 && groups > 0 && C % groups == 0,
             "pb_groupnorm_small: bad sizes");
  if ((C / groups) % 4 != 0 || ldx % 4 != 0 || ldy % 4 != 0 || rows_per_inst * (C / groups) > (1 << 22)) {
    set_error("pb_groupnorm_small: needs C/groups %% 4 == 0 and at most 4 Mi elements per (instance, group)");
    return PB_EUNSUPPORTED;
  }
  // a cluster of GN_CLUSTER CTAs per (instance, group) when its share of rows is worth a CTA, else one CTA
  const bool clustered = rows_per_inst >= 64 * GN_CLUSTER;
  dim3 grid((unsigned)(groups * (clustered ? GN_CLUSTER : 1)), (unsigned)n_inst);
  PB_DISPATCH_DTYPE(x_dtype, TXT, {
    PB_DISPATCH_DTYPE(y_dtype, TYT, {
      if (clustered)
        gn_group_kernel<TXT, TYT, GN_CLUSTER><<<grid, GN_THREADS, 0, to_stream(stream)>>>(
            reinterpret_cast<const TXT*>(x), ldx, reinterpret_cast<TYT*>(y), ldy, (int)rows_per_inst, (int)C, groups, eps,
            gamma, beta, silu);
      else
        gn_group_kernel<TXT, TYT, 1><<<grid, GN_THREADS, 0, to_stream(stream)>>>(
            reinterpret_cast<const TXT*>(x), ldx, reinterpret_cast<TYT*>(y), ldy, (int)rows_per_inst, (int)C, groups, eps,
            gamma, beta, silu);
    })
  })
  PB_LAUNCH_CHECK("gn_group_kernel");
  return PB_OK;
}

}  // extern "C"
