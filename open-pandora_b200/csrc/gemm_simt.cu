// gemm_simt.cu — CUDA-core implicit-GEMM (fp32 accumulate) for every contraction on the path.
//
// This is the exact-arithmetic path: fp32 mode of the denoiser (per-step eps rel-L2 <= 1e-3 against
// the fp32 reference needs true fp32 accumulation; TF32 would not pass), the
// tiny-M embedding GEMMs, the 8->320 stem and 320->4 head convolutions, and the on-device checker
// the tcgen05 kernels are validated against.  The bf16 hot path is pb_tc_gemm (gemm_tc.cu).
//
// C[m,n] = sum_k A[m,k] * W[n,k]  (+bias +rowbias +residual), A gathered per pb_gemm_args.a_mode:
//   LINEAR   nn.Linear / 1x1 conv                          attention.py:53-57, openaimodel3d.py:190
//   CONV3X3  k = (ky*3+kx)*cin + c, zero padding 1,        openaimodel3d.py:68,96,157,182
//            stride 1|2, optional folded nearest-x2 upsample (openaimodel3d.py:98-109)
//   CONVT3   k = kt*cin + c, zero padding over frames      openaimodel3d.py:258-269
#include "common.cuh"

namespace pb {

struct GemmP {
  const void* a;
  const void* w;
  void* c;
  const float* bias;
  const float* rowbias;
  const float* residual;
  int64_t M, N, K, lda, ldc, ldr, ldrb, rb_rows;
  int epi;
  int cin, in_h, in_w, out_h, out_w, stride, up;
  int frames, frame_rows;
};

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

// Per-thread description of the A row this thread stages (fixed for the whole K loop).
struct RowInfo {
  int64_t base;   // LINEAR: m*lda ; CONV: image index ; CONVT: (b*T)*HW + p
  int oy, ox;     // CONV3X3 output pixel ; CONVT: oy = t
  bool in_range;  // m < M
};

template <int MODE>
__device__ __forceinline__ RowInfo make_row(const GemmP& p, int64_t m) {
  RowInfo r;
  r.in_range = m < p.M;
  r.base = 0; r.oy = 0; r.ox = 0;
  if (!r.in_range) return r;
  if (MODE == PB_A_LINEAR) {
    r.base = m * p.lda;
  } else if (MODE == PB_A_CONV3X3) {
    int64_t per = (int64_t)p.out_h * p.out_w;
    int64_t n = m / per;
    int rem = (int)(m - n * per);
    r.oy = rem / p.out_w;
    r.ox = rem - r.oy * p.out_w;
    r.base = n;
  } else {
    int64_t bt = m / p.frame_rows;
    int pix = (int)(m - bt * p.frame_rows);
    int64_t b = bt / p.frames;
    r.oy = (int)(bt - b * p.frames);
    r.base = b * p.frames * (int64_t)p.frame_rows + pix;
  }
  return r;
}

// element offset of A[m, k..k+3] or -1 when the window falls into the zero padding
template <int MODE>
__device__ __forceinline__ int64_t a_offset(const GemmP& p, const RowInfo& r, int k) {
  if (MODE == PB_A_LINEAR) return r.base + k;
  int tap = k / p.cin;
  int c = k - tap * p.cin;
  if (MODE == PB_A_CONV3X3) {
    int ky = tap / 3, kx = tap - ky * 3;
    int iy = r.oy * p.stride + ky - 1;
    int ix = r.ox * p.stride + kx - 1;
    int eh = p.up ? 2 * p.in_h : p.in_h, ew = p.up ? 2 * p.in_w : p.in_w;
    if (iy < 0 || iy >= eh || ix < 0 || ix >= ew) return -1;
    if (p.up) { iy >>= 1; ix >>= 1; }
    return ((r.base * p.in_h + iy) * p.in_w + ix) * p.lda + c;
  } else {
    int tt = r.oy + tap - 1;
    if (tt < 0 || tt >= p.frames) return -1;
    return (r.base + (int64_t)tt * p.frame_rows) * p.lda + c;
  }
}

template <typename TA, typename TC, int MODE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmP p) {
  pb::pdl_trigger();
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const TA* A = reinterpret_cast<const TA*>(p.a);
  const TA* W = reinterpret_cast<const TA*>(p.w);

  const int lr = tid >> 2;        // staged row (A) / weight row (B) within the tile
  const int lk = (tid & 3) * 4;   // k offset within the BK chunk
  const RowInfo row = make_row<MODE>(p, m0 + lr);
  const int64_t wn = n0 + lr;
  const bool w_ok = wn < p.N;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = 0; k0 < p.K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = (int)k0 + lk;
    if (k < p.K) {
      if (row.in_range) {
        int64_t off = a_offset<MODE>(p, row, k);
        if (off >= 0) load4(A + off, av);
      }
      if (w_ok) load4(W + wn * p.K + k, bv);
    }
    __syncthreads();  // previous chunk fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Bs[lk + i][lr] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  TC* C = reinterpret_cast<TC*>(p.c);
  const int64_t n = n0 + tx * 4;
  if (n >= p.N) return;  // N % 4 == 0: a 4-group is entirely in or out
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.epi & PB_EPI_BIAS) load4(p.bias + n, bias4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias4[j];
    if (p.epi & PB_EPI_ROWBIAS) {
      float rb[4];
      load4(p.rowbias + (m / p.rb_rows) * p.ldrb + n, rb);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += rb[j];
    }
    if (p.epi & PB_EPI_RESIDUAL) {
      float rs[4];
      load4(p.residual + m * p.ldr + n, rs);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += rs[j];
    }
    store4(C + m * p.ldc + n, v);
  }
}

template <typename TA, typename TC>
static int launch_mode(const GemmP& p, int mode, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.M, BM), (unsigned)ceil_div(p.N, BN));
  if (mode == PB_A_LINEAR) gemm_simt_kernel<TA, TC, PB_A_LINEAR><<<grid, 256, 0, st>>>(p);
  else if (mode == PB_A_CONV3X3) gemm_simt_kernel<TA, TC, PB_A_CONV3X3><<<grid, 256, 0, st>>>(p);
  else gemm_simt_kernel<TA, TC, PB_A_CONVT3><<<grid, 256, 0, st>>>(p);
  PB_LAUNCH_CHECK("pb_gemm launch");
  return PB_OK;
}

int validate_gemm_args(const pb_gemm_args* a, const char* who) {
  PB_REQUIRE(a != nullptr, "%s: null args", who);
  PB_REQUIRE(a->a && a->w && a->c, "%s: null operand pointer", who);
  PB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "%s: empty problem M=%lld N=%lld K=%lld", who,
             (long long)a->M, (long long)a->N, (long long)a->K);
  PB_REQUIRE(a->K % 4 == 0 && a->N % 4 == 0, "%s: K and N must be multiples of 4 (K=%lld N=%lld)",
             who, (long long)a->K, (long long)a->N);
  PB_REQUIRE(a->lda % 4 == 0 && a->ldc % 4 == 0, "%s: lda/ldc must be multiples of 4", who);
  PB_REQUIRE((a->a_dtype == PB_F32 || a->a_dtype == PB_BF16) &&
                 (a->c_dtype == PB_F32 || a->c_dtype == PB_BF16),
             "%s: bad dtype code", who);
  if (a->epilogue & PB_EPI_BIAS) PB_REQUIRE(a->bias, "%s: PB_EPI_BIAS without bias", who);
  if (a->epilogue & PB_EPI_ROWBIAS)
    PB_REQUIRE(a->rowbias && a->rowbias_rows > 0 && a->ld_rowbias % 4 == 0,
               "%s: PB_EPI_ROWBIAS needs rowbias, rowbias_rows > 0, ld_rowbias %% 4 == 0", who);
  if (a->epilogue & PB_EPI_RESIDUAL)
    PB_REQUIRE(a->residual && a->ldr % 4 == 0, "%s: PB_EPI_RESIDUAL needs residual, ldr %% 4 == 0", who);
  if (a->a_mode == PB_A_CONV3X3) {
    PB_REQUIRE(a->cin > 0 && a->cin % 4 == 0 && a->K == 9 * (int64_t)a->cin,
               "%s: conv3x3 needs K == 9*cin, cin %% 4 == 0 (K=%lld cin=%d)", who, (long long)a->K, a->cin);
    PB_REQUIRE(a->stride == 1 || a->stride == 2, "%s: conv3x3 stride must be 1 or 2", who);
    PB_REQUIRE(a->in_h > 0 && a->in_w > 0 && a->out_h > 0 && a->out_w > 0, "%s: conv3x3 bad image size", who);
    int eh = a->upsample ? 2 * a->in_h : a->in_h, ew = a->upsample ? 2 * a->in_w : a->in_w;
    PB_REQUIRE(a->out_h == (eh - 1) / a->stride + 1 && a->out_w == (ew - 1) / a->stride + 1,
               "%s: conv3x3 output size %dx%d inconsistent with input %dx%d stride %d", who, a->out_h,
               a->out_w, eh, ew, a->stride);
    PB_REQUIRE(a->M % ((int64_t)a->out_h * a->out_w) == 0, "%s: conv3x3 M not a whole number of images", who);
  } else if (a->a_mode == PB_A_CONVT3) {
    PB_REQUIRE(a->cin > 0 && a->cin % 4 == 0 && a->K == 3 * (int64_t)a->cin,
               "%s: convt3 needs K == 3*cin, cin %% 4 == 0", who);
    PB_REQUIRE(a->frames > 0 && a->frame_rows > 0 &&
                   a->M % ((int64_t)a->frames * a->frame_rows) == 0,
               "%s: convt3 M must be batch*frames*frame_rows", who);
  } else {
    PB_REQUIRE(a->a_mode == PB_A_LINEAR, "%s: bad a_mode %d", who, a->a_mode);
  }
  return PB_OK;
}

// Few-row fp32 GEMM (M <= 8): the timestep / fps embedding MLPs and the 22 ResBlock emb projections folded into one
// [20160 x 1280] matrix (openaimodel3d.py:222,555,573-581) have 1-2 rows per forward, i.e. they are a read of the
// weight matrix.  One warp per output column streams its weight row with 16-byte loads and keeps the M dot products
// in registers, so the op runs at HBM rate (103 MB per forward) instead of through 64-row tiles that are 97 % padding.
constexpr int GEMV_MAX_M = 8;
__global__ void __launch_bounds__(256)
gemv_rows_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ c, int64_t ldc, int M, int N, int K) {
  pb::pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const float* wr = w + (int64_t)n * K;
  float acc[GEMV_MAX_M];
#pragma unroll
  for (int m = 0; m < GEMV_MAX_M; ++m) acc[m] = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
    for (int m = 0; m < GEMV_MAX_M; ++m) {
      if (m < M) {
        const float4 a4 = *reinterpret_cast<const float4*>(a + (int64_t)m * lda + k);
        acc[m] = fmaf(w4.x, a4.x, acc[m]); acc[m] = fmaf(w4.y, a4.y, acc[m]);
        acc[m] = fmaf(w4.z, a4.z, acc[m]); acc[m] = fmaf(w4.w, a4.w, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < GEMV_MAX_M; ++m) {
    if (m < M) {
      const float v = warp_sum(acc[m]);
      if (lane == 0) c[(int64_t)m * ldc + n] = v + (bias ? bias[n] : 0.f);
    }
  }
}

}  // namespace pb

extern "C" int pb_gemm(const pb_gemm_args* a, pb_stream_t stream) {
  using namespace pb;
  int rc = validate_gemm_args(a, "pb_gemm");
  if (rc != PB_OK) return rc;
  if (a->epilogue & PB_EPI_GEGLU) {
    set_error("pb_gemm: PB_EPI_GEGLU is only fused on the tcgen05 path; use pb_geglu");
    return PB_EUNSUPPORTED;
  }
  GemmP p;
  p.a = a->a; p.w = a->w; p.c = a->c;
  p.bias = a->bias; p.rowbias = a->rowbias; p.residual = a->residual;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.lda = a->lda; p.ldc = a->ldc; p.ldr = a->ldr; p.ldrb = a->ld_rowbias;
  p.rb_rows = a->rowbias_rows > 0 ? a->rowbias_rows : 1;
  p.epi = a->epilogue;
  p.cin = a->cin; p.in_h = a->in_h; p.in_w = a->in_w; p.out_h = a->out_h; p.out_w = a->out_w;
  p.stride = a->stride; p.up = a->upsample;
  p.frames = a->frames; p.frame_rows = a->frame_rows;
  cudaStream_t st = to_stream(stream);
  if (a->a_mode == PB_A_LINEAR && a->M <= GEMV_MAX_M && a->a_dtype == PB_F32 && a->c_dtype == PB_F32 &&
      (a->epilogue & ~PB_EPI_BIAS) == 0 && a->K % 4 == 0 && a->lda % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->w)) & 15) == 0) {
    gemv_rows_kernel<<<(unsigned)ceil_div(a->N, 8), 256, 0, st>>>(
        reinterpret_cast<const float*>(a->a), a->lda, reinterpret_cast<const float*>(a->w),
        (a->epilogue & PB_EPI_BIAS) ? a->bias : nullptr, reinterpret_cast<float*>(a->c), a->ldc, (int)a->M, (int)a->N,
        (int)a->K);
    PB_LAUNCH_CHECK("gemv_rows_kernel");
    return PB_OK;
  }
  if (a->a_dtype == PB_F32 && a->c_dtype == PB_F32) return launch_mode<float, float>(p, a->a_mode, st);
  if (a->a_dtype == PB_F32 && a->c_dtype == PB_BF16) return launch_mode<float, bf16>(p, a->a_mode, st);
  if (a->a_dtype == PB_BF16 && a->c_dtype == PB_F32) return launch_mode<bf16, float>(p, a->a_mode, st);
  return launch_mode<bf16, bf16>(p, a->a_mode, st);
}
