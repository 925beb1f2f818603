// attention_temporal.cu — softmax attention over a short sequence (<= 16 tokens, head_dim 64): the temporal
// self-attention of TemporalTransformer (attention.py:387-391; seq = 16 frames, batch = pixels x heads).
//
// HBM-bound (4 KB in, 2 KB out, 64 KFLOP per (pixel, head) in bf16), so the design goal is traffic = algorithmic bytes:
// one WARP per (pixel, head); lane l owns query (l % 16) and one 32-wide half of the head dimension; K and V
// (16 x 64 each) are staged once in shared memory and read as broadcasts; the two halves of every q.k dot are
// joined with one shuffle; scores stay in registers (no online softmax needed for 16 keys).
// Rows are read in place from the (b, t, p)-ordered activation through the generic pb_attn_operand strides.
#include <math.h>

#include "common.cuh"

namespace pb {

constexpr int TS_MAX = 16;      // tokens
constexpr int TD = 64;          // head dim
constexpr int T_WARPS = 4;      // (pixel, head) groups per CTA

struct SmallAttnP {
  const void *q, *k, *v;
  void* out;
  int64_t q_bdiv, q_bso, q_bsi, q_rs;
  int64_t k_bdiv, k_bso, k_bsi, k_rs;
  int64_t v_bdiv, v_bso, v_bsi, v_rs;
  int64_t o_bdiv, o_bso, o_bsi, o_rs;
  int64_t batch, heads;
  int n;
  float scale_log2;
  int accumulate;
};

template <typename T>
__global__ void __launch_bounds__(32 * T_WARPS) attn_small_kernel(SmallAttnP p) {
  __shared__ __align__(16) float Ks[T_WARPS][TS_MAX][TD];
  __shared__ __align__(16) float Vs[T_WARPS][TS_MAX][TD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t grp = (int64_t)blockIdx.x * T_WARPS + warp;       // (batch, head) index
  if (grp >= p.batch * p.heads) return;                           // whole warp exits together
  const int64_t n = grp / p.heads;
  const int h = (int)(grp - n * p.heads);
  const T* Q = reinterpret_cast<const T*>(p.q) + (n / p.q_bdiv) * p.q_bso + (n % p.q_bdiv) * p.q_bsi + h * TD;
  const T* K = reinterpret_cast<const T*>(p.k) + (n / p.k_bdiv) * p.k_bso + (n % p.k_bdiv) * p.k_bsi + h * TD;
  const T* V = reinterpret_cast<const T*>(p.v) + (n / p.v_bdiv) * p.v_bso + (n % p.v_bdiv) * p.v_bsi + h * TD;

  // stage K, V: 16 rows x 64 = 256 4-element vectors each, 8 per lane
#pragma unroll 2
  for (int e = lane; e < TS_MAX * TD / 4; e += 32) {
    const int j = e >> 4, d = (e & 15) * 4;
    float tk[4] = {0.f, 0.f, 0.f, 0.f}, tv[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < p.n) {
      load4(K + j * p.k_rs + d, tk);
      load4(V + j * p.v_rs + d, tv);
    }
    *reinterpret_cast<float4*>(&Ks[warp][j][d]) = make_float4(tk[0], tk[1], tk[2], tk[3]);
    *reinterpret_cast<float4*>(&Vs[warp][j][d]) = make_float4(tv[0], tv[1], tv[2], tv[3]);
  }
  const int qi = lane & 15, half = lane >> 4;
  const bool q_ok = qi < p.n;
  float q[32];
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (q_ok) load4(Q + qi * p.q_rs + half * 32 + d, t);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[d + i] = t[i] * p.scale_log2;
  }
  __syncwarp();

  float s[TS_MAX];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < TS_MAX; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 32; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&Ks[warp][j][half * 32 + d]);
      a = fmaf(q[d], kk.x, a); a = fmaf(q[d + 1], kk.y, a); a = fmaf(q[d + 2], kk.z, a); a = fmaf(q[d + 3], kk.w, a);
    }
    a += __shfl_xor_sync(0xffffffffu, a, 16);       // join the two 32-wide halves of the dot product
    s[j] = (j < p.n) ? a : -INFINITY;
    mx = fmaxf(mx, s[j]);
    asm volatile("" ::: "memory");   // keep the 16 x 8 shared loads from being hoisted into registers at once
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TS_MAX; ++j) { s[j] = exp2f(s[j] - mx); l += s[j]; }
  const float inv = 1.0f / l;
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < TS_MAX; ++j) {
#pragma unroll
    for (int d = 0; d < 32; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(&Vs[warp][j][half * 32 + d]);
      o[d] = fmaf(s[j], vv.x, o[d]); o[d + 1] = fmaf(s[j], vv.y, o[d + 1]);
      o[d + 2] = fmaf(s[j], vv.z, o[d + 2]); o[d + 3] = fmaf(s[j], vv.w, o[d + 3]);
    }
    asm volatile("" ::: "memory");
  }
  if (!q_ok) return;
  T* O = reinterpret_cast<T*>(p.out) + (n / p.o_bdiv) * p.o_bso + (n % p.o_bdiv) * p.o_bsi + h * TD + qi * p.o_rs +
         half * 32;
#pragma unroll
  for (int d = 0; d < 32; d += 4) {
    float r[4] = {o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv};
    if (p.accumulate) {
      float prev[4];
      load4(O + d, prev);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] += prev[i];
    }
    store4(O + d, r);
  }
}

// called by pb_attention for n_q == n_kv <= 16
int launch_small_attention(const pb_attn_args* a, cudaStream_t st) {
  SmallAttnP p;
  p.q = a->q.ptr; p.k = a->k.ptr; p.v = a->v.ptr; p.out = a->out;
  p.q_bdiv = a->q.bdiv; p.q_bso = a->q.bs_outer; p.q_bsi = a->q.bs_inner; p.q_rs = a->q.rs;
  p.k_bdiv = a->k.bdiv; p.k_bso = a->k.bs_outer; p.k_bsi = a->k.bs_inner; p.k_rs = a->k.rs;
  p.v_bdiv = a->v.bdiv; p.v_bso = a->v.bs_outer; p.v_bsi = a->v.bs_inner; p.v_rs = a->v.rs;
  p.o_bdiv = a->o_bdiv; p.o_bso = a->o_bs_outer; p.o_bsi = a->o_bs_inner; p.o_rs = a->o_rs;
  p.batch = a->batch; p.heads = a->heads; p.n = (int)a->n_q;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  const int64_t groups = a->batch * a->heads;
  const int64_t blocks = ceil_div(groups, T_WARPS);
  PB_REQUIRE(blocks <= 2147483647LL, "pb_attention: grid too large");
  if (a->dtype == PB_F32) attn_small_kernel<float><<<(unsigned)blocks, 32 * T_WARPS, 0, st>>>(p);
  else attn_small_kernel<bf16><<<(unsigned)blocks, 32 * T_WARPS, 0, st>>>(p);
  PB_LAUNCH_CHECK("attn_small_kernel");
  return PB_OK;
}

}  // namespace pb
