// elementwise.cu — streaming kernels: GEGLU, SiLU, add, strided row copy/cast, sinusoidal embedding,
// video<->rows layout change at the U-Net boundary, and the fused CFG + DDIM update.
#include <math.h>

#include "common.cuh"

namespace pb {

static inline unsigned grid_for(int64_t n, int threads = 256, int64_t cap = 148 * 32) {
  int64_t b = ceil_div(n, threads);
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

// ------------------------------------------------------------------------------------------ GEGLU
template <typename TU, typename TY>
__global__ void geglu_kernel(const TU* __restrict__ u, int64_t ldu, TY* __restrict__ y, int64_t ldy,
                             int64_t n_out, int64_t total_vec) {
  pb::pdl_trigger();
  const int64_t cv = n_out >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cv, v = i - row * cv;
    float a[4], g[4], o[4];
    load4(u + row * ldu + v * 4, a);
    load4(u + row * ldu + n_out + v * 4, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = a[j] * gelu_f(g[j]);
    store4(y + row * ldy + v * 4, o);
  }
}

__global__ void silu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  pb::pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v / (1.0f + expf(-v));
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  pb::pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] + b[i];
}

template <typename TS, typename TD>
__global__ void copy_rows_kernel(const TS* __restrict__ s, int64_t lds, TD* __restrict__ d, int64_t ldd,
                                 int64_t cols, int64_t total_vec) {
  pb::pdl_trigger();
  const int64_t cv = cols >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / cv, v = i - row * cv;
    float t[4];
    load4(s + row * lds + v * 4, t);
    store4(d + row * ldd + v * 4, t);
  }
}

// nearest-neighbour x2 upsample of NHWC rows (openaimodel3d.py:98-109), with dtype conversion
template <typename TS, typename TD>
__global__ void upsample2x_kernel(const TS* __restrict__ s, int64_t lds, TD* __restrict__ d, int64_t ldd,
                                  int H, int W, int64_t cols, int64_t total_vec) {
  pb::pdl_trigger();
  const int64_t cv = cols >> 2;
  const int W2 = 2 * W, H2 = 2 * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t orow = i / cv, v = i - orow * cv;
    const int ox = (int)(orow % W2);
    const int64_t t = orow / W2;
    const int oy = (int)(t % H2);
    const int64_t n = t / H2;
    const int64_t irow = (n * H + (oy >> 1)) * W + (ox >> 1);
    float x[4];
    load4(s + irow * lds + v * 4, x);
    store4(d + orow * ldd + v * 4, x);
  }
}

// out[b, j] = cos(t_b * f_j), out[b, half + j] = sin(t_b * f_j); precise cosf/sinf (arguments reach ~1e3)
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ out, int64_t batch, int64_t half) {
  pb::pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * half) return;
  const int64_t b = i / half, j = i - b * half;
  const float arg = (float)t[b] * freqs[j];
  out[b * 2 * half + j] = cosf(arg);
  out[b * 2 * half + half + j] = sinf(arg);
}

// (b, c, t, hw) x2 -> rows (b, t, p) x (c1 + c2), fp32
template <typename T>
__global__ void video_to_rows_kernel(const T* __restrict__ x1, int c1, const T* __restrict__ x2, int c2,
                                     float* __restrict__ rows, int64_t ld, int64_t b, int64_t t, int64_t hw) {
  pb::pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over (b, t, p)
  if (i >= b * t * hw) return;
  const int64_t p = i % hw, bt = i / hw, tt = bt % t, bb = bt / t;
  float* r = rows + i * ld;
  for (int c = 0; c < c1; ++c) r[c] = to_f(x1[((bb * c1 + c) * t + tt) * hw + p]);
  for (int c = 0; c < c2; ++c) r[c1 + c] = to_f(x2[((bb * c2 + c) * t + tt) * hw + p]);
}

template <typename T>
__global__ void rows_to_video_kernel(const float* __restrict__ rows, int64_t ld, T* __restrict__ out,
                                     int64_t b, int c, int64_t t, int64_t hw) {
  pb::pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b * t * hw) return;
  const int64_t p = i % hw, bt = i / hw, tt = bt % t, bb = bt / t;
  const float* r = rows + i * ld;
  for (int cc = 0; cc < c; ++cc) out[((bb * c + cc) * t + tt) * hw + p] = from_f<T>(r[cc]);
}

// --------------------------------------------------------------------------------- CFG + DDIM step
// ws layout (double): [batch][4] = sum(v_cond), sumsq(v_cond), sum(v_cfg), sumsq(v_cfg)
__global__ void ddim_zero_ws_kernel(double* ws, int n) {
  pb::pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ws[i] = 0.0;
}

// The guided prediction, rounded op by op like the reference's tensor expression:
//   two branches   v_u + s*(v_c - v_u)                                            ddim.py:238
//   three branches v_u + s_img*(v_ui - v_u) + s*(v_c - v_ui)                      ddim_multiplecond.py:234
__device__ __forceinline__ float guided2(float c, float u, float s) {
  return __fadd_rn(u, __fmul_rn(s, __fsub_rn(c, u)));
}
__device__ __forceinline__ float guided3(float c, float u, float ui, float s, float s_img) {
  return __fadd_rn(__fadd_rn(u, __fmul_rn(s_img, __fsub_rn(ui, u))), __fmul_rn(s, __fsub_rn(c, ui)));
}

__global__ void __launch_bounds__(256)
ddim_stats_kernel(const float* __restrict__ vc, const float* __restrict__ vu, const float* __restrict__ vi,
                  int64_t n, float s, float s_img, double* __restrict__ ws) {
  pb::pdl_trigger();
  const int b = blockIdx.y;
  const float* c = vc + (int64_t)b * n;
  const float* u = vu + (int64_t)b * n;
  const float* ui = vi ? vi + (int64_t)b * n : nullptr;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = c[i], y = u[i];
    const float g = ui ? guided3(x, y, ui[i], s, s_img) : guided2(x, y, s);
    a0 += x; a1 += (double)x * x; a2 += g; a3 += (double)g * g;
  }
  __shared__ double red[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][w] = a0; red[1][w] = a1; red[2][w] = a2; red[3][w] = a3; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(ws + b * 4 + threadIdx.x, t);
  }
}

// Arithmetic mirrors the reference op by op (separately rounded fp32 mul/add, no FMA contraction):
//   v    = guided2 / guided3 above                              ddim.py:238, ddim_multiplecond.py:234
//   v    = phi*(v*(std_c/std)) + (1-phi)*v                      utils_diffusion.py:152-157
//   e    = sa*v + so*x ; x0 = sa*x - so*v                       ddpm3d.py:235-247
//   x0  *= rescale                                              ddim.py:273-277
//   xp   = sqrt(a_prev)*x0 + dir*e + sigma*noise*temperature    ddim.py:282-288
__global__ void __launch_bounds__(256)
ddim_update_kernel(const float* x /* may alias xp: in-place update of the graph's static latent */,
                   const float* __restrict__ vc, const float* __restrict__ vu, const float* __restrict__ vi,
                   const float* __restrict__ noise, float* xp, float* __restrict__ x0o,
                   int64_t n, float s, float s_img, float phi, float temperature, const float* __restrict__ tables,
                   const int32_t* __restrict__ step_index, const double* __restrict__ ws) {
  pb::pdl_trigger();
  const int b = blockIdx.y;
  const float* tb = tables + (int64_t)(*step_index) * PB_DDIM_NCOEF;
  const float sa = tb[PB_DDIM_SQRT_AC], so = tb[PB_DDIM_SQRT_OMAC], sap = tb[PB_DDIM_SQRT_A_PREV];
  const float sigma = tb[PB_DDIM_SIGMA], resc = tb[PB_DDIM_X0_RESCALE], dir = tb[PB_DDIM_DIR_COEF];
  float ratio = 1.f;
  if (phi > 0.f && vu != nullptr) {
    const double cnt = (double)n;
    const double mc = ws[b * 4 + 0] / cnt, mg = ws[b * 4 + 2] / cnt;
    double var_c = (ws[b * 4 + 1] - cnt * mc * mc) / (cnt - 1.0);
    double var_g = (ws[b * 4 + 3] - cnt * mg * mg) / (cnt - 1.0);
    ratio = (float)sqrt(var_c > 0 ? var_c : 0.0) / (float)sqrt(var_g > 0 ? var_g : 0.0);
  }
  const int64_t off = (int64_t)b * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xi = x[off + i], c = vc[off + i];
    float v = c;
    if (vu != nullptr) {
      const float u = vu[off + i];
      v = vi ? guided3(c, u, vi[off + i], s, s_img) : guided2(c, u, s);
      if (phi > 0.f)
        v = __fadd_rn(__fmul_rn(phi, __fmul_rn(v, ratio)), __fmul_rn(1.0f - phi, v));
    }
    const float e = __fadd_rn(__fmul_rn(sa, v), __fmul_rn(so, xi));
    float x0 = __fsub_rn(__fmul_rn(sa, xi), __fmul_rn(so, v));
    x0 = __fmul_rn(x0, resc);
    const float nz = __fmul_rn(__fmul_rn(sigma, noise[off + i]), temperature);
    xp[off + i] = __fadd_rn(__fadd_rn(__fmul_rn(sap, x0), __fmul_rn(dir, e)), nz);
    x0o[off + i] = x0;
  }
}

__global__ void ddim_advance_kernel(int32_t* step_index) { *step_index += 1; }

}  // namespace pb

extern "C" {

int pb_geglu(const void* u, int32_t u_dtype, int64_t ldu, void* y, int32_t y_dtype, int64_t ldy,
             int64_t rows, int64_t n_out, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(u && y && rows > 0 && n_out > 0, "pb_geglu: bad arguments");
  PB_REQUIRE(n_out % 4 == 0 && ldu % 4 == 0 && ldy % 4 == 0, "pb_geglu: sizes must be multiples of 4");
  const int64_t total = rows * (n_out / 4);
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(u_dtype, TU, {
    PB_DISPATCH_DTYPE(y_dtype, TY, {
      geglu_kernel<TU, TY><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const TU*>(u), ldu,
                                                            reinterpret_cast<TY*>(y), ldy, n_out, total);
    })
  })
  PB_LAUNCH_CHECK("geglu_kernel");
  return PB_OK;
}

int pb_silu(const float* x, float* y, int64_t n, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && y && n > 0, "pb_silu: bad arguments");
  silu_kernel<<<grid_for(n), 256, 0, to_stream(stream)>>>(x, y, n);
  PB_LAUNCH_CHECK("silu_kernel");
  return PB_OK;
}

int pb_add(const float* a, const float* b, float* y, int64_t n, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(a && b && y && n > 0, "pb_add: bad arguments");
  add_kernel<<<grid_for(n), 256, 0, to_stream(stream)>>>(a, b, y, n);
  PB_LAUNCH_CHECK("add_kernel");
  return PB_OK;
}

int pb_copy_rows(const void* src, int32_t s_dtype, int64_t lds, void* dst, int32_t d_dtype,
                 int64_t ldd, int64_t rows, int64_t cols, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(src && dst && rows > 0 && cols > 0, "pb_copy_rows: bad arguments");
  PB_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "pb_copy_rows: sizes must be multiples of 4");
  const int64_t total = rows * (cols / 4);
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(s_dtype, TS, {
    PB_DISPATCH_DTYPE(d_dtype, TD, {
      copy_rows_kernel<TS, TD><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const TS*>(src), lds,
                                                                reinterpret_cast<TD*>(dst), ldd, cols, total);
    })
  })
  PB_LAUNCH_CHECK("copy_rows_kernel");
  return PB_OK;
}

int pb_upsample2x_rows(const void* src, int32_t s_dtype, int64_t lds, void* dst, int32_t d_dtype, int64_t ldd,
                       int64_t n_img, int64_t H, int64_t W, int64_t cols, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(src && dst && n_img > 0 && H > 0 && W > 0 && cols > 0, "pb_upsample2x_rows: bad arguments");
  PB_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "pb_upsample2x_rows: sizes must be multiples of 4");
  const int64_t total = n_img * 4 * H * W * (cols / 4);
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(s_dtype, TS, {
    PB_DISPATCH_DTYPE(d_dtype, TD, {
      upsample2x_kernel<TS, TD><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const TS*>(src), lds,
                                                                 reinterpret_cast<TD*>(dst), ldd, (int)H, (int)W, cols, total);
    })
  })
  PB_LAUNCH_CHECK("upsample2x_kernel");
  return PB_OK;
}

int pb_timestep_embedding(const int64_t* t, const float* freqs, float* out, int64_t batch,
                          int64_t half, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(t && freqs && out && batch > 0 && half > 0, "pb_timestep_embedding: bad arguments");
  const int64_t n = batch * half;
  timestep_embedding_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, to_stream(stream)>>>(t, freqs, out, batch, half);
  PB_LAUNCH_CHECK("timestep_embedding_kernel");
  return PB_OK;
}

int pb_video_to_rows(const void* x1, int64_t c1, const void* x2, int64_t c2, int32_t in_dtype,
                     float* rows, int64_t ld, int64_t b, int64_t t, int64_t hw, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x1 && rows && c1 > 0 && b > 0 && t > 0 && hw > 0, "pb_video_to_rows: bad arguments");
  PB_REQUIRE(c2 == 0 || x2, "pb_video_to_rows: c2 > 0 needs x2");
  PB_REQUIRE(ld >= c1 + c2, "pb_video_to_rows: ld too small");
  const int64_t n = b * t * hw;
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(in_dtype, T, {
    video_to_rows_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(
        reinterpret_cast<const T*>(x1), (int)c1, reinterpret_cast<const T*>(x2), (int)c2, rows, ld, b, t, hw);
  })
  PB_LAUNCH_CHECK("video_to_rows_kernel");
  return PB_OK;
}

int pb_rows_to_video(const float* rows, int64_t ld, void* out, int32_t out_dtype, int64_t b,
                     int64_t c, int64_t t, int64_t hw, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(rows && out && c > 0 && b > 0 && t > 0 && hw > 0 && ld >= c, "pb_rows_to_video: bad arguments");
  const int64_t n = b * t * hw;
  cudaStream_t st = to_stream(stream);
  PB_DISPATCH_DTYPE(out_dtype, T, {
    rows_to_video_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(rows, ld, reinterpret_cast<T*>(out), b,
                                                                         (int)c, t, hw);
  })
  PB_LAUNCH_CHECK("rows_to_video_kernel");
  return PB_OK;
}

size_t pb_ddim_ws_bytes(int64_t batch) { return batch > 0 ? (size_t)batch * 4 * sizeof(double) : 0; }

int pb_ddim_cfg3_step(const float* x, const float* v_cond, const float* v_uncond, const float* v_uncond_img,
                      const float* noise, float* x_prev, float* pred_x0, int64_t batch, int64_t n,
                      float cfg_scale, float cfg_img, float guidance_rescale, float temperature,
                      const float* tables, int32_t* step_index, int32_t advance, void* ws, size_t ws_bytes,
                      pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(x && v_cond && noise && x_prev && pred_x0 && tables && step_index, "pb_ddim_cfg_step: null pointer");
  PB_REQUIRE(batch > 0 && batch <= 65535 && n > 0, "pb_ddim_cfg_step: bad sizes");
  PB_REQUIRE(v_uncond || !v_uncond_img, "pb_ddim_cfg3_step: the image-only branch needs the unconditional one");
  cudaStream_t st = to_stream(stream);
  const bool rescale = guidance_rescale > 0.f && v_uncond != nullptr;
  unsigned gx = grid_for(n, 256, 148 * 4);
  if (rescale) {
    PB_REQUIRE(ws && ws_bytes >= pb_ddim_ws_bytes(batch), "pb_ddim_cfg_step: workspace too small");
    double* w = reinterpret_cast<double*>(ws);
    ddim_zero_ws_kernel<<<(unsigned)ceil_div(batch * 4, 128), 128, 0, st>>>(w, (int)(batch * 4));
    PB_LAUNCH_CHECK("ddim_zero_ws_kernel");
    ddim_stats_kernel<<<dim3(gx, (unsigned)batch), 256, 0, st>>>(v_cond, v_uncond, v_uncond_img, n, cfg_scale, cfg_img, w);
    PB_LAUNCH_CHECK("ddim_stats_kernel");
  }
  ddim_update_kernel<<<dim3(gx, (unsigned)batch), 256, 0, st>>>(
      x, v_cond, v_uncond, v_uncond_img, noise, x_prev, pred_x0, n, cfg_scale, cfg_img,
      rescale ? guidance_rescale : 0.f, temperature, tables, step_index, reinterpret_cast<const double*>(ws));
  PB_LAUNCH_CHECK("ddim_update_kernel");
  if (advance) {
    ddim_advance_kernel<<<1, 1, 0, st>>>(step_index);
    PB_LAUNCH_CHECK("ddim_advance_kernel");
  }
  return PB_OK;
}

int pb_ddim_cfg_step(const float* x, const float* v_cond, const float* v_uncond, const float* noise,
                     float* x_prev, float* pred_x0, int64_t batch, int64_t n, float cfg_scale,
                     float guidance_rescale, float temperature, const float* tables,
                     int32_t* step_index, int32_t advance, void* ws, size_t ws_bytes,
                     pb_stream_t stream) {
  return pb_ddim_cfg3_step(x, v_cond, v_uncond, nullptr, noise, x_prev, pred_x0, batch, n, cfg_scale, 0.f,
                           guidance_rescale, temperature, tables, step_index, advance, ws, ws_bytes, stream);
}

}  // extern "C"
