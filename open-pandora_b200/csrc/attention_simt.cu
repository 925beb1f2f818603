// attention_simt.cu — CUDA-core softmax attention, head_dim 64, fp32 math (exact-arithmetic path).
//
// One thread owns one query row: q and the output accumulator live in registers, K/V tiles are
// staged in shared memory and read as warp-wide broadcasts, online softmax in base 2.  It takes any
// (n_q, n_kv) and the generic batch addressing of pb_attn_operand, so the same kernel serves
// spatial self-attention, text / image cross-attention (attention.py:101-135) and temporal
// attention read in place from (b, t, p)-ordered rows.  The bf16 hot path is pb_tc_attention.
#include <math.h>

#include "common.cuh"

namespace pb {

constexpr int AD = 64;   // head dim
constexpr int KT = 32;   // keys per shared-memory tile

struct Opnd {
  const void* ptr;
  int64_t bdiv, bso, bsi, rs;
};

struct AttnP {
  Opnd q, k, v;
  void* out;
  int64_t o_bdiv, o_bso, o_bsi, o_rs;
  int64_t batch, heads, nq, nkv;
  int qtiles;
  float scale_log2;
  int accumulate;
};

__device__ __forceinline__ int64_t batch_off(int64_t n, int64_t bdiv, int64_t bso, int64_t bsi) {
  return (n / bdiv) * bso + (n % bdiv) * bsi;
}

template <typename T>
__global__ void __launch_bounds__(128) attn_simt_kernel(AttnP p) {
  pb::pdl_trigger();
  __shared__ __align__(16) float Ks[KT][AD];
  __shared__ __align__(16) float Vs[KT][AD];
  const int64_t n = blockIdx.x / p.qtiles;
  const int qt = blockIdx.x % p.qtiles;
  const int h = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t qi = (int64_t)qt * nthr + tid;
  const bool q_ok = qi < p.nq;

  const T* Q = reinterpret_cast<const T*>(p.q.ptr) + batch_off(n, p.q.bdiv, p.q.bso, p.q.bsi) + h * AD;
  const T* K = reinterpret_cast<const T*>(p.k.ptr) + batch_off(n, p.k.bdiv, p.k.bso, p.k.bsi) + h * AD;
  const T* V = reinterpret_cast<const T*>(p.v.ptr) + batch_off(n, p.v.bdiv, p.v.bso, p.v.bsi) + h * AD;

  float q[AD], acc[AD];
#pragma unroll
  for (int d = 0; d < AD; d += 4) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (q_ok) load4(Q + qi * p.q.rs + d, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { q[d + j] = t[j] * p.scale_log2; acc[d + j] = 0.f; }
  }
  float m = -INFINITY, l = 0.f;

  for (int64_t k0 = 0; k0 < p.nkv; k0 += KT) {
    __syncthreads();
    for (int e = tid; e < KT * AD / 4; e += nthr) {
      const int j = e / (AD / 4), d = (e % (AD / 4)) * 4;
      float tk[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + j < p.nkv) {
        load4(K + (k0 + j) * p.k.rs + d, tk);
        load4(V + (k0 + j) * p.v.rs + d, tv);
      }
      *reinterpret_cast<float4*>(&Ks[j][d]) = make_float4(tk[0], tk[1], tk[2], tk[3]);
      *reinterpret_cast<float4*>(&Vs[j][d]) = make_float4(tv[0], tv[1], tv[2], tv[3]);
    }
    __syncthreads();
    const int jn = (int)((p.nkv - k0) < KT ? (p.nkv - k0) : KT);
    for (int j = 0; j < jn; ++j) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < AD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s);
        s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
      }
      const float m_new = fmaxf(m, s);
      const float corr = exp2f(m - m_new);   // exp2f(-inf) = 0 on the first key
      const float pj = exp2f(s - m_new);
      l = l * corr + pj;
      m = m_new;
#pragma unroll
      for (int d = 0; d < AD; d += 4) {
        float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        acc[d] = fmaf(acc[d], corr, pj * vv.x);
        acc[d + 1] = fmaf(acc[d + 1], corr, pj * vv.y);
        acc[d + 2] = fmaf(acc[d + 2], corr, pj * vv.z);
        acc[d + 3] = fmaf(acc[d + 3], corr, pj * vv.w);
      }
    }
  }
  if (!q_ok) return;
  const float inv = 1.0f / l;
  T* O = reinterpret_cast<T*>(p.out) + batch_off(n, p.o_bdiv, p.o_bso, p.o_bsi) + h * AD + qi * p.o_rs;
#pragma unroll
  for (int d = 0; d < AD; d += 4) {
    float o[4] = {acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv};
    if (p.accumulate) {
      float prev[4];
      load4(O + d, prev);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += prev[j];
    }
    store4(O + d, o);
  }
}

int validate_attn_args(const pb_attn_args* a, const char* who) {
  PB_REQUIRE(a != nullptr, "%s: null args", who);
  PB_REQUIRE(a->q.ptr && a->k.ptr && a->v.ptr && a->out, "%s: null operand", who);
  PB_REQUIRE(a->batch > 0 && a->heads > 0 && a->n_q > 0 && a->n_kv > 0,
             "%s: empty problem (batch=%lld heads=%lld n_q=%lld n_kv=%lld)", who, (long long)a->batch,
             (long long)a->heads, (long long)a->n_q, (long long)a->n_kv);
  PB_REQUIRE(a->q.bdiv > 0 && a->k.bdiv > 0 && a->v.bdiv > 0 && a->o_bdiv > 0, "%s: bdiv must be > 0", who);
  PB_REQUIRE(a->dtype == PB_F32 || a->dtype == PB_BF16, "%s: bad dtype", who);
  PB_REQUIRE(a->q.rs % 4 == 0 && a->k.rs % 4 == 0 && a->v.rs % 4 == 0 && a->o_rs % 4 == 0 &&
                 a->q.bs_outer % 4 == 0 && a->q.bs_inner % 4 == 0 && a->k.bs_outer % 4 == 0 &&
                 a->k.bs_inner % 4 == 0 && a->v.bs_outer % 4 == 0 && a->v.bs_inner % 4 == 0 &&
                 a->o_bs_outer % 4 == 0 && a->o_bs_inner % 4 == 0,
             "%s: strides must be multiples of 4 elements", who);
  PB_REQUIRE(a->heads <= 65535, "%s: too many heads", who);
  return PB_OK;
}

}  // namespace pb

namespace pb { int launch_small_attention(const pb_attn_args* a, cudaStream_t st); }

extern "C" int pb_attention(const pb_attn_args* a, pb_stream_t stream) {
  using namespace pb;
  int rc = validate_attn_args(a, "pb_attention");
  if (rc != PB_OK) return rc;
  if (a->n_q == a->n_kv && a->n_q <= 16) return launch_small_attention(a, to_stream(stream));  // temporal
  AttnP p;
  p.q = {a->q.ptr, a->q.bdiv, a->q.bs_outer, a->q.bs_inner, a->q.rs};
  p.k = {a->k.ptr, a->k.bdiv, a->k.bs_outer, a->k.bs_inner, a->k.rs};
  p.v = {a->v.ptr, a->v.bdiv, a->v.bs_outer, a->v.bs_inner, a->v.rs};
  p.out = a->out;
  p.o_bdiv = a->o_bdiv; p.o_bso = a->o_bs_outer; p.o_bsi = a->o_bs_inner; p.o_rs = a->o_rs;
  p.batch = a->batch; p.heads = a->heads; p.nq = a->n_q; p.nkv = a->n_kv;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  int threads = a->n_q >= 128 ? 128 : (int)(((a->n_q + 31) / 32) * 32);
  p.qtiles = (int)ceil_div(a->n_q, threads);
  const int64_t gx = a->batch * p.qtiles;
  PB_REQUIRE(gx <= 2147483647LL, "pb_attention: grid too large");
  dim3 grid((unsigned)gx, (unsigned)a->heads);
  cudaStream_t st = to_stream(stream);
  if (a->dtype == PB_F32) attn_simt_kernel<float><<<grid, threads, 0, st>>>(p);
  else attn_simt_kernel<bf16><<<grid, threads, 0, st>>>(p);
  PB_LAUNCH_CHECK("attn_simt_kernel");
  return PB_OK;
}
