// attention_temporal.cu — softmax attention over a short sequence (<= 16 tokens, head_dim 64): the temporal
// self-attention of TemporalTransformer (attention.py:387-391; seq = 16 frames, batch = pixels x heads).
//
// HBM-bound (4 KB in, 2 KB out, 64 KFLOP per (pixel, head) in bf16), so the design goal is traffic = algorithmic bytes:
// one WARP per (pixel, head); lane l owns query (l % 16) and one 32-wide half of the head dimension; K and V
// (16 x 64 each) are staged once in shared memory and read as broadcasts; the two halves of every q.k dot are
// joined with one shuffle; scores stay in registers (no online softmax needed for 16 keys).
// Rows are read in place from the (b, t, p)-ordered activation through the generic pb_attn_operand strides.
#include <math.h>
#include <stdlib.h>

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
  pb::pdl_trigger();
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

// ---------------------------------------------------------------------------------------------------
// bf16 fast path: the same warp-per-(pixel, head) decomposition, but Q / K / V (16 x 64 bf16 = 2 KB each) are
// staged with cp.async (16 B per lane, all 12 copies of a warp in flight at once, no register staging), and the
// two 16x16x64 / 16x64x16 products run on mma.sync m16n8k16 (a 16-token sequence is far below the 128-row
// tcgen05 tile; the op is HBM-bound and the point of the tensor-core path is to take the 1024 FMAs per lane
// off the issue slots so that the loads set the pace).  Scores never leave registers: the S accumulator
// fragments are exactly the A fragments of the P.V product.  O is staged through the (dead) Q tile so the
// global stores are 16 B per lane.
constexpr int TM_WARPS = 8;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;      // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(32 * TM_WARPS) attn_small_mma_kernel(SmallAttnP p) {
  pb::pdl_trigger();
  // per warp: Q, K, V tiles [16 rows][64 bf16] = 128 B rows, 16-byte chunk c of row j stored at chunk c ^ (j & 7)
  __shared__ __align__(128) uint8_t tiles[TM_WARPS][3][TS_MAX * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t grp = (int64_t)blockIdx.x * TM_WARPS + warp;
  if (grp >= p.batch * p.heads) return;
  const int64_t n = grp / p.heads;
  const int h = (int)(grp - n * p.heads);
  const bf16* Q = reinterpret_cast<const bf16*>(p.q) + (n / p.q_bdiv) * p.q_bso + (n % p.q_bdiv) * p.q_bsi + h * TD;
  const bf16* K = reinterpret_cast<const bf16*>(p.k) + (n / p.k_bdiv) * p.k_bso + (n % p.k_bdiv) * p.k_bsi + h * TD;
  const bf16* V = reinterpret_cast<const bf16*>(p.v) + (n / p.v_bdiv) * p.v_bso + (n % p.v_bdiv) * p.v_bsi + h * TD;
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(&tiles[warp][0][0]);
  const uint32_t sK = sQ + TS_MAX * 128, sV = sK + TS_MAX * 128;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = lane + 32 * i, j = e >> 3, c = e & 7;
    const uint32_t off = (uint32_t)(j * 128 + ((c ^ (j & 7)) << 4));
    const bool ok = j < p.n;
    cp_async16(sQ + off, Q + (int64_t)j * p.q_rs + c * 8, ok);
    cp_async16(sK + off, K + (int64_t)j * p.k_rs + c * 8, ok);
    cp_async16(sV + off, V + (int64_t)j * p.v_rs + c * 8, ok);
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  // ---- S = Q K^T : two n-tiles (keys 0-7, 8-15), four k-steps over d
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4], b[4];
    {
      const int row = lane & 15, c = 2 * ks + (lane >> 4);
      ldsm_x4(sQ + (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)), a);
    }
    {
      const int key = ((lane >> 4) << 3) + (lane & 7), c = 2 * ks + ((lane >> 3) & 1);
      ldsm_x4(sK + (uint32_t)(key * 128 + ((c ^ (key & 7)) << 4)), b);
    }
    mma_bf16_16816(s0, a, b[0], b[1]);
    mma_bf16_16816(s1, a, b[2], b[3]);
  }
  // ---- softmax over the 16 keys of rows g = lane/4 and g + 8; a thread holds keys {2t, 2t+1, 8+2t, 9+2t}
  const int t2 = (lane & 3) * 2;
  const float NEG = -INFINITY;
  float v0[4] = {s0[0], s0[1], s1[0], s1[1]};      // row g
  float v1[4] = {s0[2], s0[3], s1[2], s1[3]};      // row g + 8
  const int keys[4] = {t2, t2 + 1, 8 + t2, 9 + t2};
  float m0 = NEG, m1 = NEG;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v0[i] = keys[i] < p.n ? v0[i] * p.scale_log2 : NEG;
    v1[i] = keys[i] < p.n ? v1[i] * p.scale_log2 : NEG;
    m0 = fmaxf(m0, v0[i]); m1 = fmaxf(m1, v1[i]);
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v0[i] = exp2f(v0[i] - m0); l0 += v0[i];
    v1[i] = exp2f(v1[i] - m1); l1 += v1[i];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  // P as the A fragment of the second product (normalised first, so O needs no rescale)
  uint32_t pa[4] = {pack_bf16x2(v0[0] * inv0, v0[1] * inv0), pack_bf16x2(v1[0] * inv1, v1[1] * inv1),
                    pack_bf16x2(v0[2] * inv0, v0[3] * inv0), pack_bf16x2(v1[2] * inv1, v1[3] * inv1)};
  // ---- O = P V : eight n-tiles of 8 head-dim columns, one k-step (16 keys); V is [key][d] -> transposed ldmatrix
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    uint32_t b[4];
    const int key = (lane & 7) + (((lane >> 3) & 1) << 3), c = 2 * np + (lane >> 4);
    ldsm_x4_t(sV + (uint32_t)(key * 128 + ((c ^ (key & 7)) << 4)), b);
    mma_bf16_16816(o[2 * np], pa, b[0], b[1]);
    mma_bf16_16816(o[2 * np + 1], pa, b[2], b[3]);
  }
  // ---- stage O (bf16) through the Q tile, then 16-byte coalesced stores
  __syncwarp();
  const int g = lane >> 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    // columns nt*8 + t2, +1 -> chunk nt, byte offset t2*2 inside the chunk
    const uint32_t a0 = sQ + (uint32_t)(g * 128 + ((nt ^ (g & 7)) << 4) + t2 * 2);
    const uint32_t a1 = sQ + (uint32_t)((g + 8) * 128 + ((nt ^ ((g + 8) & 7)) << 4) + t2 * 2);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(pack_bf16x2(o[nt][0], o[nt][1])) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(pack_bf16x2(o[nt][2], o[nt][3])) : "memory");
  }
  __syncwarp();
  bf16* O = reinterpret_cast<bf16*>(p.out) + (n / p.o_bdiv) * p.o_bso + (n % p.o_bdiv) * p.o_bsi + h * TD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = lane + 32 * i, j = e >> 3, c = e & 7;
    if (j >= p.n) continue;
    uint4 val;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                 : "r"(sQ + (uint32_t)(j * 128 + ((c ^ (j & 7)) << 4))));
    uint4* dst = reinterpret_cast<uint4*>(O + (int64_t)j * p.o_rs + c * 8);
    if (p.accumulate) {
      const uint4 prev = *dst;
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&val);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&prev);
      uint32_t r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        r[q] = pack_bf16x2(__low2float(a2[q]) + __low2float(b2[q]), __high2float(a2[q]) + __high2float(b2[q]));
      val = make_uint4(r[0], r[1], r[2], r[3]);
    }
    *dst = val;
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
  if (a->dtype == PB_F32) {
    attn_small_kernel<float><<<(unsigned)blocks, 32 * T_WARPS, 0, st>>>(p);
  } else {
    // tensor-core path needs 16-byte addressable rows (every stride a multiple of 8 bf16, 16-byte aligned bases)
    const int64_t strides = a->q.bs_outer | a->q.bs_inner | a->q.rs | a->k.bs_outer | a->k.bs_inner | a->k.rs |
                            a->v.bs_outer | a->v.bs_inner | a->v.rs | a->o_bs_outer | a->o_bs_inner | a->o_rs;
    const uintptr_t ptrs = reinterpret_cast<uintptr_t>(a->q.ptr) | reinterpret_cast<uintptr_t>(a->k.ptr) |
                           reinterpret_cast<uintptr_t>(a->v.ptr) | reinterpret_cast<uintptr_t>(a->out);
    static const bool mma_on = []() { const char* e = getenv("PB_TEMPORAL_MMA"); return !(e && e[0] == '0'); }();
    if (mma_on && strides % 8 == 0 && (ptrs & 15) == 0) {
      const int64_t blocks8 = ceil_div(groups, TM_WARPS);
      attn_small_mma_kernel<<<(unsigned)blocks8, 32 * TM_WARPS, 0, st>>>(p);
    } else {
      attn_small_kernel<bf16><<<(unsigned)blocks, 32 * T_WARPS, 0, st>>>(p);
    }
  }
  PB_LAUNCH_CHECK("attn_small_kernel");
  return PB_OK;
}

}  // namespace pb
