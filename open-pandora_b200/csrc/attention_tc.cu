// attention_tc.cu — tcgen05 flash attention, head_dim 64, bf16 operands, fp32 softmax / accumulation in tensor memory.
//
// Two kernels (details above each):
//   tc_attention3_kernel   several key tiles (spatial self-attention, 160 ... 9216 tokens): one CTA per (frame, head,
//                          PAIR of 128-query tiles), K/V tiles through a 4-stage TMA ring, S / P / O in tensor memory,
//                          two softmax warpgroups + a polling MMA issuer.
//   tc_attention_x_kernel  one key tile (text + image cross-attention with two softmax segments, the smallest
//                          self-attention levels): K|V resident, query tiles streamed, bulk-tensor-store output.
// Key masking handles n_kv that is not a multiple of 128 (77 text / 16 image tokens, 160, 40, ...).
// kv_split > 0 (n_kv <= 128): keys [0, split) and [split, n_kv) are two independent softmaxes whose outputs are
// summed — text + image cross-attention (attention.py:128-142) in one pass over one concatenated K/V tile.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pb {
namespace tc {

constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 64;
constexpr int UMMA_K = 16;
constexpr int TILE_BYTES = 128 * 64 * 2;      // 16 KiB: one [128 x 64] bf16 tile

struct AttP {
  void* out;
  int64_t o_bdiv, o_bso, o_bsi, o_rs;
  int q_bdiv, k_bdiv, v_bdiv;
  int nq, nkv, kv_tiles;
  float scale_log2;
  int accumulate;
  int split;             // > 0: two softmax segments [0, split) and [split, nkv) inside ONE key tile
  int q_chunk;           // single-key-tile kernel: query tiles per CTA
  int o_tma;             // single-key-tile kernel: the output leaves as bulk tensor stores through tmO
};

#ifdef PB_ATT_TRACE
// Profiling build only (tools/attn_timeline.py): per-phase SM clock stamps of CTA (0,0,0) of tc_attention3_kernel.
// Layout (int64): [role 0..2][ATT_TRACE_MAX]; role 0 = issuer, 1 / 2 = warp 0 of the softmax group of tile 0 / 1.
constexpr int ATT_TRACE_MAX = 4096;
__device__ long long* g_att_trace = nullptr;
#define ATT_TRACE_DECL(role)                                                                                     \
  long long* tr_ = (g_att_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0) ? g_att_trace + (role) * ATT_TRACE_MAX : nullptr; \
  int tr_n_ = 0;
#define ATT_TRACE(tag)                                                                          \
  do {                                                                                          \
    if (tr_ && tr_n_ < ATT_TRACE_MAX - 1) tr_[tr_n_++] = (clock64() << 8) | (long long)(tag);  \
  } while (0)
#else
#define ATT_TRACE_DECL(role)
#define ATT_TRACE(tag)
#endif

// MN-major B operand tile [k rows][64 contiguous n], 128-byte swizzle: 8-row K groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;   // LBO: stride between 64-wide N blocks (single block here)
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO: stride between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// MUFU ex2 (approximate, flush-to-zero): ~2 ulp, the standard flash-attention choice
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// The same packing on the integer pipe (round half up: +0x8000, keep the high halves) for finite non-negative
// inputs — the softmax probabilities.  F2FP shares the quarter-rate XU pipe with MUFU ex2 (measured: XU 65 % busy with
// 96 ex2 + 64 F2FP per thread and key tile), and that pipe is what bounds d = 64 attention.
__device__ __forceinline__ uint32_t pack_bf16_alu(float a, float b) {
  return __byte_perm(__float_as_uint(a) + 0x8000u, __float_as_uint(b) + 0x8000u, 0x7632);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32f(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),
      "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]),
      "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])
      : "memory");
}

// packed fp32x2 arithmetic (sm_100): one instruction for two lanes of the scale-and-shift and of the row sum
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %4};\nmov.b64 rc, {%5, %5};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1) {
  asm("{\n.reg .b64 ra, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rd, {%0, %1};\nadd.rn.f32x2 rd, rd, ra;\nmov.b64 {%0, %1}, rd;\n}"
      : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1));
}

// 2^e for a PAIR of scores on the FMA pipe (no MUFU): e = n + f with n = round(e), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P), and n goes
// straight into the exponent field.  Packed f32x2 adds / FMAs: 10 issue slots per pair against 2 MUFU.EX2 that
// occupy the 4-lane MUFU pipe for 16 cycles.  Used for a fixed fraction of the pairs so that both pipes are busy.
__device__ __forceinline__ void exp2_poly2(float e0, float e1, float& p0, float& p1) {
  e0 = fmaxf(e0, -126.0f);
  e1 = fmaxf(e1, -126.0f);
  float r0, r1, q0, q1;
  asm("{\n.reg .b64 a, m, nm, r, n, f, q, c;\n"
      "mov.b64 a, {%4, %5};\n"
      "mov.b64 m, {%6, %6};\n"
      "mov.b64 nm, {%7, %7};\n"
      "add.rn.f32x2 r, a, m;\n"          // low mantissa bits = round(e) (two's complement)
      "add.rn.f32x2 n, r, nm;\n"         // n = round(e) as a float
      "mov.b64 c, {%8, %8};\n"
      "fma.rn.f32x2 f, n, c, a;\n"       // f = e - n
      "mov.b64 q, {%9, %9};\n"
      "mov.b64 c, {%10, %10};\n"
      "fma.rn.f32x2 q, q, f, c;\n"
      "mov.b64 c, {%11, %11};\n"
      "fma.rn.f32x2 q, q, f, c;\n"
      "mov.b64 c, {%12, %12};\n"
      "fma.rn.f32x2 q, q, f, c;\n"
      "mov.b64 {%0, %1}, r;\n"
      "mov.b64 {%2, %3}, q;\n}"
      : "=f"(r0), "=f"(r1), "=f"(q0), "=f"(q1)
      : "f"(e0), "f"(e1), "f"(12582912.0f), "f"(-12582912.0f), "f"(-1.0f), "f"(0.055171672f), "f"(0.24261114f),
        "f"(0.69326097f), "f"(0.99992806f));
  p0 = __int_as_float((__float_as_int(r0) << 23) + __float_as_int(q0));
  p1 = __int_as_float((__float_as_int(r1) << 23) + __float_as_int(q1));
}

// ================================================================================================
// Single key tile (n_kv <= 128): text + image cross-attention (93 keys per frame, two softmax segments), and the
// self-attention of the smallest levels.  The work is HBM-shaped (a 128-query tile reads 16 KiB of Q and writes
// 16 KiB of O for 2 x 128 x n_kv x 64 MACs), so the kernel is built around not re-doing per-tile setup: a CTA owns
// a CHUNK of consecutive query tiles of one (frame, head); K and V are fetched once and stay in shared memory, tensor
// memory and the barriers are set up once, Q tiles stream through a 2-deep TMA ring, and the output tile leaves as
// ONE bulk tensor store from a swizzled staging buffer (full 128-byte lines).  80 KiB of shared memory and 256 TMEM
// columns per CTA -> two CTAs per SM overlap each other's softmax / MMA / store phases.
//   warp 0  TMA producer   K, V once; Q tiles into the ring
//   warp 1  MMA issuer     S_i = Q_i K^T (UMMA 128 x n_kv16 x 64) -> TMEM [0,128);  O_i = P_i V (A = P from TMEM, which
//                          aliases the first 64 columns of S; B = V MN-major) -> TMEM [128,192).  S_i+1 is issued right
//                          behind PV_i, so it is computed while the softmax warps are still storing tile i.
//   warps 2..5  softmax + output: one thread per query row; the score row is read ONCE into registers, both segment
//                          softmaxes are normalised there, P goes back to tensor memory as bf16.
// (The kernel this replaces launched one CTA per query tile — 3200 CTAs at level 0, each re-loading the same 93-row
// K|V tile and paying TMEM allocation + barrier setup for one tile of work: 565 GB/s, 0.09 of the HBM roofline.)
constexpr int X_QSTAGES = 2;
constexpr int X_SMEM = TILE_BYTES * (2 + X_QSTAGES + 1);          // K, V, Q ring, O staging = 80 KiB
constexpr int X_NBARS = 1 + 2 * X_QSTAGES + 4;
constexpr int X_O_COL0 = 128;

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// NCH = 32-key chunks of the score row held in registers: 3 (n_kv <= 96: the 77 + 16 cross-attention keys; 2 CTAs per
// SM) or 4 (n_kv <= 128; one CTA per SM, the 128-value row does not fit 168 registers).
// CSP / CNK > 0: the segment split and key count are compile-time constants (the Open-Pandora context is always 77 text
// + 16 image tokens per frame): every per-key segment test folds away; 0 = runtime values.
template <int NCH, int CSP, int CNK>
__global__ void __launch_bounds__(192, NCH == 3 ? 2 : 1)
tc_attention_x_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttP p) {
  pb::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = base + TILE_BYTES;
  auto sQ = [&](int s) { return base + TILE_BYTES * (2 + s); };
  const uint32_t sO = base + TILE_BYTES * (2 + X_QSTAGES);
  const uint32_t bars = base + X_SMEM;
  const uint32_t kv_full = bars;
  auto q_full = [&](int s) { return bars + 8u * (1 + s); };
  auto q_empty = [&](int s) { return bars + 8u * (1 + X_QSTAGES + s); };
  const uint32_t s_full = bars + 8u * (1 + 2 * X_QSTAGES), p_full = s_full + 8, o_full = s_full + 16, o_free = s_full + 24;
  const uint32_t tmem_slot = bars + 8u * X_NBARS;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y, n = blockIdx.z;
  const int q_tiles = (p.nq + ATT_BM - 1) / ATT_BM;
  const int i0 = blockIdx.x * p.q_chunk;
  const int nt = (q_tiles - i0 < p.q_chunk) ? q_tiles - i0 : p.q_chunk;       // query tiles of this CTA (>= 1)
  const int nkv16 = (p.nkv + 15) & ~15;                                        // keys the products run over

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < X_QSTAGES; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    mbar_init(s_full, 1); mbar_init(o_full, 1);
    mbar_init(p_full, 4); mbar_init(o_free, 4);                     // one arrival per softmax warp
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // Producer and issuer loops run on the whole warp in uniform control flow with one elected lane issuing (see the
  // note at the issuer of tc_attention3_kernel: under `if (lane == 0)` every issue sits in an election loop).
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);       // rows past n_kv are zero-filled by the copy engine
      tma_load_3d(sK, &tmK, kv_full, head * ATT_D, 0, n / p.k_bdiv);
      tma_load_3d(sV, &tmV, kv_full, head * ATT_D, 0, n / p.v_bdiv);
    }
    __syncwarp();
    for (int it = 0; it < nt; ++it) {
      const int s = it % X_QSTAGES;
      mbar_wait(q_empty(s), ((it / X_QSTAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(q_full(s), TILE_BYTES);
        tma_load_3d(sQ(s), &tmQ, q_full(s), head * ATT_D, (i0 + it) * ATT_BM, n / p.q_bdiv);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc_bf16(ATT_BM, (uint32_t)nkv16);
    const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);     // B (= V) MN-major
    const uint64_t k_desc = make_smem_desc_sw128(sK);
    const uint64_t v_desc = make_smem_desc_mn_sw128(sV);
    auto issue_s = [&](int it) {
      const int s = it % X_QSTAGES;
      mbar_wait(q_full(s), (it / X_QSTAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t q_desc = make_smem_desc_sw128(sQ(s));
#pragma unroll
        for (int k = 0; k < ATT_D / UMMA_K; ++k) umma_bf16(tmem_base, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(q_empty(s));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_s(0);
    const int ksteps = nkv16 / UMMA_K;
    for (int it = 0; it < nt; ++it) {
      mbar_wait(p_full, it & 1);                          // P_i is in tensor memory (S_i has been consumed)
      if (it > 0) mbar_wait(o_free, (it - 1) & 1);        // O_i-1 has been read out
      tc_fence_after();
      if (elect_one()) {
        for (int k = 0; k < ksteps; ++k)                  // 16 keys = 2048 B = 128 descriptor units
          umma_bf16_ts(tmem_base + X_O_COL0, tmem_base + (uint32_t)(k * 8), v_desc + (uint64_t)(k * 128), idesc_o, k != 0);
        umma_commit(o_full);
      }
      __syncwarp();
      if (it + 1 < nt) issue_s(it + 1);                   // executes behind PV_i: P_i is consumed before S overwrites it
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t o_addr = lane_addr + X_O_COL0;
    const int sp = CNK > 0 ? CSP : (p.split > 0 ? p.split : p.nkv), nk = CNK > 0 ? CNK : p.nkv;
    const int nchunk = (nk + 31) >> 5;          // <= NCH
    const bool tma_out = p.o_tma != 0;
    const bool leader = warp == 2 && lane == 0;
    for (int it = 0; it < nt; ++it) {
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      float v[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if (c < nchunk) tmem_ld32(lane_addr + c * 32, v[c]);
      tmem_ld_wait();
      // four independent chains per segment (a single fmaxf / add chain over 93 values is 93 dependent instructions)
      float m1[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, m2[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c < nchunk) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c * 32 + i;
            if (col < sp) m1[i & 3] = fmaxf(m1[i & 3], v[c][i]);
            else if (col < nk) m2[i & 3] = fmaxf(m2[i & 3], v[c][i]);
          }
        }
      }
      const float mx1 = fmaxf(fmaxf(m1[0], m1[1]), fmaxf(m1[2], m1[3])) * p.scale_log2;
      const float mx2 = fmaxf(fmaxf(m2[0], m2[1]), fmaxf(m2[2], m2[3])) * p.scale_log2;
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c < nchunk) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c * 32 + i;
            float e = 0.f;
            if (col < sp) { e = fast_exp2(fmaf(v[c][i], p.scale_log2, -mx1)); s1[i & 3] += e; }
            else if (col < nk) { e = fast_exp2(fmaf(v[c][i], p.scale_log2, -mx2)); s2[i & 3] += e; }
            v[c][i] = e;
          }
        }
      }
      const float l1 = (s1[0] + s1[1]) + (s1[2] + s1[3]), l2 = (s2[0] + s2[1]) + (s2[2] + s2[3]);
      const float i1 = 1.0f / l1, i2 = l2 > 0.f ? 1.0f / l2 : 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c < nchunk) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int col = c * 32 + i;
            w[i >> 1] = pack_bf16_alu(v[c][i] * (col < sp ? i1 : i2), v[c][i + 1] * (col + 1 < sp ? i1 : i2));
          }
          tmem_st16(lane_addr + (uint32_t)(c * 16), w);       // P aliases the head of S: 32 keys = 16 packed columns
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- output of this tile
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      float o0[32], o1[32];
      tmem_ld32(o_addr, o0);
      tmem_ld32(o_addr + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (tma_out) {
        // the staging buffer is free once the previous tile's bulk store has read it
        if (it > 0) {
          if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const uint32_t row = sO + (uint32_t)r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float* src = c < 4 ? &o0[c * 8] : &o1[(c - 4) * 8];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (uint32_t)((c ^ (r & 7)) << 4)),
                       "r"(pack_bf16(src[0], src[1])), "r"(pack_bf16(src[2], src[3])), "r"(pack_bf16(src[4], src[5])),
                       "r"(pack_bf16(src[6], src[7])) : "memory");
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (leader) {      // rows past n_q are clipped by the copy engine
          tma_store_3d(&tmO, sO, head * ATT_D, (i0 + it) * ATT_BM, n);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        const int64_t qi = (int64_t)(i0 + it) * ATT_BM + r;
        if (qi < p.nq) {
          bf16* O = reinterpret_cast<bf16*>(p.out) + (n / p.o_bdiv) * p.o_bso + (n % p.o_bdiv) * p.o_bsi + qi * p.o_rs +
                    head * ATT_D;
#pragma unroll
          for (int c = 0; c < ATT_D; c += 8) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = c < 32 ? o0[c + i] : o1[c - 32 + i];
            uint4* dst = reinterpret_cast<uint4*>(O + c);
            if (p.accumulate) {
              uint4 prev = *dst;
              const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
              for (int i = 0; i < 4; ++i) { o[2 * i] += __low2float(pv[i]); o[2 * i + 1] += __high2float(pv[i]); }
            }
            uint4 val;
            val.x = pack_bf16(o[0], o[1]); val.y = pack_bf16(o[2], o[3]);
            val.z = pack_bf16(o[4], o[5]); val.w = pack_bf16(o[6], o[7]);
            *dst = val;
          }
        }
      }
    }
    if (tma_out && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // smem must outlive the last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ================================================================================================
// Self-attention over several key tiles: the FlashAttention-4 data flow, with the score and probability tiles
// DECOUPLED so that neither side ever waits on a round trip through the other.
//   * S_t,j (128 x 128 fp32) is read from tensor memory ONCE into registers; the softmax warps then hand the S
//     buffer back (s_free) at the START of their work, and the issuer puts S_t,j+1 = Q_t K_j+1^T on the tensor pipe
//     while the exponentials of tile j are still being computed: when the softmax of tile j ends, the next score
//     tile is already there.
//   * P_t,j (bf16) has its own 64 columns of tensor memory (tcgen05.st) and is the A operand of the PV product
//     straight from TMEM (no shared-memory round trip, no proxy fence).
//   * O_t accumulates in tensor memory across key tiles; it is rescaled only when a row maximum grows by more than
//     2^8 (lazy rescale, decided per warp), otherwise the stale maximum is kept — the probabilities then stay below
//     2^8, exact in fp32 and harmless in bf16 — so there is no per-tile fold of O.
//   * the issuer is an event loop: it polls s_free / p_full of both query tiles and issues whichever product has its
//     operands ready, so the two query tiles drift into whatever phase keeps the MUFU and the tensor pipe busiest.
//     The whole warp runs that loop in uniform control flow with one elected lane issuing (see the note there).
//   * one eighth of the exponentials is evaluated on the FMA pipe (exp2_poly2) to take load off the MUFU.
// Measured alternatives that were NOT better (tools/attn_timeline.py, same box): strict ping-pong of the two
// tiles' exponential phases through named barriers (each phase alone keeps the MUFU only 70 % busy: -3 %), one
// straight-line issuer warp per tile instead of the polling loop (-5 %), 2/8 .. 4/8 polynomial share (0 .. -10 %).
// TMEM columns: S_0 [0,128) S_1 [128,256) P_0 [256,320) P_1 [320,384) O_0 [384,448) O_1 [448,512).
// (Measured on the version this replaces, where S_t,j+1 could only be issued after PV_t,j because P aliased S:
// 39 % of the softmax warps' samples were spent waiting for the next score tile, tensor pipe 26 % active.)
constexpr int V3_STAGES = 4;
constexpr int V3_SMEM = TILE_BYTES * (2 + 2 * V3_STAGES);      // Q x2 + (K, V) x4 = 160 KiB
constexpr float V3_RESCALE_THRESHOLD = 8.0f;                    // log2 units
constexpr int V3_P_COL0 = 256, V3_O_COL0 = 384;
constexpr int V3_POLY_EIGHTHS = 1;      // eighths of the exponentials evaluated on the FMA pipe (exp2_poly2)
constexpr int V3_NBARS = 1 + 2 * V3_STAGES + 8;                 // q_full, kv_full/empty, {s_full, s_free, p_full, o_full} x 2

// 384 threads = three warpgroups: {TMA producer, MMA issuer, two idle warps} | softmax of query tile 0 | softmax of
// query tile 1.  setmaxnreg moves registers from the first warpgroup (88 each) to the softmax groups (208 each), so a
// thread holds its whole 128-wide score row.
__global__ void __launch_bounds__(384, 1)
tc_attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttP p) {
  pb::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int t) { return base + TILE_BYTES * t; };
  auto sK = [&](int s) { return base + TILE_BYTES * (2 + s); };
  auto sV = [&](int s) { return base + TILE_BYTES * (2 + V3_STAGES + s); };
  const uint32_t bars = base + V3_SMEM;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + V3_STAGES + s); };
  auto s_full = [&](int t) { return bars + 8u * (1 + 2 * V3_STAGES + t); };
  auto s_free = [&](int t) { return bars + 8u * (3 + 2 * V3_STAGES + t); };
  auto p_full = [&](int t) { return bars + 8u * (5 + 2 * V3_STAGES + t); };
  auto o_full = [&](int t) { return bars + 8u * (7 + 2 * V3_STAGES + t); };
  const uint32_t tmem_slot = bars + 8u * V3_NBARS;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int q_tile0 = blockIdx.x * 2, head = blockIdx.y, n = blockIdx.z;
  const int n_tiles = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < V3_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1); mbar_init(o_full(t), 1);
      mbar_init(s_free(t), 4); mbar_init(p_full(t), 4);          // one arrival per softmax warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  }
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        tma_load_3d(sQ(t), &tmQ, q_full, head * ATT_D, (q_tile0 + t) * ATT_BM, n / p.q_bdiv);
    }
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % V3_STAGES;
      mbar_wait(kv_empty(s), ((j / V3_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(kv_full(s), 2 * TILE_BYTES);
        tma_load_3d(sK(s), &tmK, kv_full(s), head * ATT_D, j * ATT_BN, n / p.k_bdiv);
        tma_load_3d(sV(s), &tmV, kv_full(s), head * ATT_D, j * ATT_BN, n / p.v_bdiv);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // The WHOLE warp runs the event loop in uniform control flow (`warp` comes from a shuffle, so the compiler
    // knows it is warp-uniform) and one elected lane issues.  With the loop under `if (lane == 0)` every
    // tcgen05.mma sat in a compiler-generated election loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~100 clk per
    // MMA, measured with tools/attn_timeline.py): 24 MMAs per key tile made the ISSUER the bottleneck of the
    // kernel (3500 clk per key tile against 2050 clk of exponentials).
    const uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);
    const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);     // B (= V) MN-major
    int s_iss0 = 0, s_iss1 = 0, pv_iss0 = 0, pv_iss1 = 0, released = 0;
    uint32_t idle = 0;
    ATT_TRACE_DECL(0)
    mbar_wait(q_full, 0);
    ATT_TRACE(0);
    const uint64_t q_desc0 = make_smem_desc_sw128(sQ(0)), q_desc1 = make_smem_desc_sw128(sQ(1));
    while (pv_iss0 < n_tiles || pv_iss1 < n_tiles) {
      bool did = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        int& s_iss = t ? s_iss1 : s_iss0;
        int& pv_iss = t ? pv_iss1 : pv_iss0;
        if (s_iss < n_tiles) {
          // S_t,j overwrites S_t,j-1: the softmax warps must have it in registers; K_j must have landed
          const int j = s_iss, st = j % V3_STAGES;
          if ((j == 0 || mbar_test(s_free(t), (j - 1) & 1)) && mbar_test(kv_full(st), (j / V3_STAGES) & 1)) {
            tc_fence_after();
            ATT_TRACE(5);
            if (elect_one()) {
              const uint64_t q_desc = t ? q_desc1 : q_desc0;
              const uint64_t k_desc = make_smem_desc_sw128(sK(st));
              const uint32_t d = tmem_base + (uint32_t)t * ATT_BN;
#pragma unroll
              for (int k = 0; k < ATT_D / UMMA_K; ++k) umma_bf16(d, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
              umma_commit(s_full(t));
            }
            __syncwarp();
            ATT_TRACE(1 + t);
            ++s_iss;
            did = true;
          }
        }
        if (pv_iss < s_iss) {
          const int j = pv_iss, st = j % V3_STAGES;
          if (mbar_test(p_full(t), j & 1)) {   // P_t,j is in tensor memory, O_t has been rescaled if it had to be
            tc_fence_after();
            ATT_TRACE(6);
            ++pv_iss;
            // K_j / V_j are free once both query tiles have issued every product on them
            const int lo = pv_iss0 < pv_iss1 ? pv_iss0 : pv_iss1;
            if (elect_one()) {
              const uint32_t d = tmem_base + V3_O_COL0 + (uint32_t)t * ATT_D;
              const uint32_t a_tm = tmem_base + V3_P_COL0 + (uint32_t)t * 64;     // bf16 pairs: 8 columns per k-step
              const uint64_t b_desc = make_smem_desc_mn_sw128(sV(st));
#pragma unroll
              for (int k = 0; k < ATT_BN / UMMA_K; ++k)                            // 16 keys = 2048 B = 128 desc units
                umma_bf16_ts(d, a_tm + (uint32_t)(k * 8), b_desc + (uint64_t)(k * 128), idesc_o, (j | k) != 0);
              umma_commit(o_full(t));
              for (int r = released; r < lo; ++r) umma_commit(kv_empty(r % V3_STAGES));
            }
            __syncwarp();
            released = lo > released ? lo : released;
            ATT_TRACE(3 + t);
            did = true;
          }
        }
      }
      if (did) { idle = 0; continue; }
      if (++idle > (1u << 24)) {
        if (lane == 0)
          printf("pandora_b200: attention issuer timeout (block %d,%d,%d s %d/%d pv %d/%d)\n", blockIdx.x, blockIdx.y,
                 blockIdx.z, s_iss0, s_iss1, pv_iss0, pv_iss1);
        __trap();
      }
    }
  } else if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)t * ATT_BN;
    const uint32_t p_addr = lane_addr + V3_P_COL0 + (uint32_t)t * 64;
    const uint32_t o_addr = lane_addr + V3_O_COL0 + (uint32_t)t * ATT_D;
    float m_run = -INFINITY, l = 0.f;
#ifdef PB_ATT_TRACE
    long long* tr_ = (g_att_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && quarter == 0 && lane == 0)
                         ? g_att_trace + (1 + t) * ATT_TRACE_MAX : nullptr;
    int tr_n_ = 0;
#endif

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = p.nkv - j * ATT_BN;
      const bool full = valid >= ATT_BN;
      ATT_TRACE(10);
      mbar_wait(s_full(t), j & 1);
      tc_fence_after();
      ATT_TRACE(11);
      float v[4][32];
      tmem_ld32(s_addr, v[0]);
      tmem_ld32(s_addr + 32, v[1]);
      tmem_ld32(s_addr + 64, v[2]);
      tmem_ld32(s_addr + 96, v[3]);
      tmem_ld_wait();
      // the score tile is in registers: hand the buffer back so that S_t,j+1 is computed under this tile's softmax
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(t));
      ATT_TRACE(12);
      float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], v[c][i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], (c * 32 + i < valid) ? v[c][i] : -INFINITY);
        }
      }
      const float m_tile = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])) * p.scale_log2;
      float corr = 1.0f;
      bool rescale = false;
      if (j == 0) {
        m_run = m_tile;
      } else {
        const bool need = m_tile > m_run + V3_RESCALE_THRESHOLD;
        rescale = __any_sync(0xffffffffu, need);          // warp-uniform: tcgen05.ld / st are warp-collective
        if (need) { corr = fast_exp2(m_run - m_tile); m_run = m_tile; l *= corr; }
      }
      float lq[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_run;
      uint32_t w[4][16];
      ATT_TRACE(13);
      ATT_TRACE(14);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (full) {                                       // no per-element key masking on all but the last tile
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1;
            fma2(e0, e1, v[c][i], v[c][i + 1], p.scale_log2, neg_m);
            float p0, p1;
            if (((c * 16 + (i >> 1)) & 7) < V3_POLY_EIGHTHS) exp2_poly2(e0, e1, p0, p1);
            else { p0 = fast_exp2(e0); p1 = fast_exp2(e1); }
            add2(lq[(i & 2)], lq[(i & 2) + 1], p0, p1);
            w[c][i >> 1] = pack_bf16_alu(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c * 32 + i < valid) ? fast_exp2(fmaf(v[c][i], p.scale_log2, neg_m)) : 0.f;
            const float p1 = (c * 32 + i + 1 < valid) ? fast_exp2(fmaf(v[c][i + 1], p.scale_log2, neg_m)) : 0.f;
            lq[(i >> 1) & 3] += p0 + p1;
            w[c][i >> 1] = pack_bf16_alu(p0, p1);
          }
        }
      }
      ATT_TRACE(15);
      l += (lq[0] + lq[1]) + (lq[2] + lq[3]);
      if (j > 0) {
        // PV_t,j-1 must be complete before P is overwritten (and before O may be rescaled); it was issued a whole
        // softmax period ago
        mbar_wait(o_full(t), (j - 1) & 1);
        tc_fence_after();
        ATT_TRACE(16);
        if (rescale) {
#pragma unroll 1
          for (int c = 0; c < ATT_D; c += 32) {
            float o0[32];
            tmem_ld32(o_addr + c, o0);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o0[i] *= corr;
            tmem_st32f(o_addr + c, o0);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(p_addr + (uint32_t)(c * 16), w[c]);       // 32 keys = 16 packed columns
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(t));
      ATT_TRACE(17);
    }
    // ---- epilogue: O_t / l
    mbar_wait(o_full(t), (n_tiles - 1) & 1);
    tc_fence_after();
    const int64_t qi = (int64_t)(q_tile0 + t) * ATT_BM + r;
    float o0[32], o1[32];
    tmem_ld32(o_addr, o0);
    tmem_ld32(o_addr + 32, o1);
    tmem_ld_wait();
    if (qi < p.nq) {
      const float inv = 1.0f / l;
      bf16* O = reinterpret_cast<bf16*>(p.out) + (n / p.o_bdiv) * p.o_bso + (n % p.o_bdiv) * p.o_bsi + qi * p.o_rs +
                head * ATT_D;
#pragma unroll
      for (int c = 0; c < ATT_D; c += 8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (c < 32 ? o0[c + i] : o1[c - 32 + i]) * inv;
        uint4* dst = reinterpret_cast<uint4*>(O + c);
        if (p.accumulate) {
          uint4 prev = *dst;
          const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&prev);
#pragma unroll
          for (int i = 0; i < 4; ++i) { o[2 * i] += __low2float(pv[i]); o[2 * i + 1] += __high2float(pv[i]); }
        }
        uint4 val;
        val.x = pack_bf16(o[0], o[1]); val.y = pack_bf16(o[2], o[3]);
        val.z = pack_bf16(o[4], o[5]); val.w = pack_bf16(o[6], o[7]);
        *dst = val;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static int encode_rows3d(CUtensorMap* m, const pb_attn_operand& o, int64_t cols, int64_t rows, int64_t n_batch) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)rows, (uint64_t)(n_batch < 1 ? 1 : n_batch)};
  uint64_t bs = o.bs_outer > 0 ? (uint64_t)o.bs_outer : (uint64_t)rows * o.rs;
  uint64_t str[2] = {(uint64_t)o.rs * 2, bs * 2};
  uint32_t box[3] = {64, 128, 1};
  return encode_tmap_bf16(m, o.ptr, 3, dims, str, box);
}

}  // namespace tc
}  // namespace pb

#ifdef PB_ATT_TRACE
extern "C" int pb_debug_attn_trace(void* device_buffer) {
  long long* ptr = reinterpret_cast<long long*>(device_buffer);
  return cudaMemcpyToSymbol(pb::tc::g_att_trace, &ptr, sizeof(ptr)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int pb_tc_attention(const pb_attn_args* a, pb_stream_t stream) {
  using namespace pb;
  using namespace pb::tc;
  int rc = validate_attn_args(a, "pb_tc_attention");
  if (rc != PB_OK) return rc;
#define ATT_UNSUPPORTED(...)   \
  do {                         \
    set_error(__VA_ARGS__);    \
    return PB_EUNSUPPORTED;    \
  } while (0)
  if (a->dtype != PB_BF16) ATT_UNSUPPORTED("pb_tc_attention: bf16 only");
  if (a->q.bs_inner != 0 || a->k.bs_inner != 0 || a->v.bs_inner != 0)
    ATT_UNSUPPORTED("pb_tc_attention: strided (temporal) batches use the dedicated kernel");
  if (a->n_q < 64) ATT_UNSUPPORTED("pb_tc_attention: n_q < 64 is served by the CUDA-core kernel");
  if ((a->q.rs | a->k.rs | a->v.rs | a->o_rs) % 8 != 0 || (a->q.bs_outer | a->k.bs_outer | a->v.bs_outer) % 8 != 0)
    ATT_UNSUPPORTED("pb_tc_attention: strides must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a->q.ptr) | reinterpret_cast<uintptr_t>(a->k.ptr) |
       reinterpret_cast<uintptr_t>(a->v.ptr) | reinterpret_cast<uintptr_t>(a->out)) & 15)
    ATT_UNSUPPORTED("pb_tc_attention: operands must be 16-byte aligned");
  if (a->batch > 65535 || a->heads > 65535) ATT_UNSUPPORTED("pb_tc_attention: grid too large");

  CUtensorMap tq, tk, tv;
  const int64_t cols = a->heads * 64;
  if ((rc = encode_rows3d(&tq, a->q, cols, a->n_q, ceil_div(a->batch, a->q.bdiv))) != PB_OK) return rc;
  if ((rc = encode_rows3d(&tk, a->k, cols, a->n_kv, ceil_div(a->batch, a->k.bdiv))) != PB_OK) return rc;
  if ((rc = encode_rows3d(&tv, a->v, cols, a->n_kv, ceil_div(a->batch, a->v.bdiv))) != PB_OK) return rc;

  AttP p;
  memset(&p, 0, sizeof(p));
  p.out = a->out;
  p.o_bdiv = a->o_bdiv; p.o_bso = a->o_bs_outer; p.o_bsi = a->o_bs_inner; p.o_rs = a->o_rs;
  p.q_bdiv = (int)a->q.bdiv; p.k_bdiv = (int)a->k.bdiv; p.v_bdiv = (int)a->v.bdiv;
  p.nq = (int)a->n_q; p.nkv = (int)a->n_kv;
  p.kv_tiles = (int)ceil_div(a->n_kv, ATT_BN);
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.accumulate = a->accumulate;
  p.split = a->kv_split;

  if (a->kv_split < 0 || (a->kv_split > 0 && (a->n_kv > ATT_BN || a->kv_split >= a->n_kv)))
    ATT_UNSUPPORTED("pb_tc_attention: kv_split needs 0 < kv_split < n_kv <= 128");
  static bool configured = false;
  const size_t smem_x = X_SMEM + 1024 + 8 * X_NBARS + 16;
  const size_t smem_v3 = V3_SMEM + 1024 + 8 * V3_NBARS + 16;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_attention_x_kernel<3, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_attention_x_kernel<3, 77, 93>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_attention_x_kernel<4, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_attention3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v3);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_attention kernels)");
    configured = true;
  }
  if (p.kv_tiles == 1) {
    // query tiles per CTA: amortise the per-CTA setup (K|V fetch, TMEM, barriers ~ 2 tiles of work) without leaving
    // the tail wave half empty; 2 CTAs per SM
    const int q_tiles = (int)ceil_div(a->n_q, ATT_BM);
    const bool narrow = a->n_kv <= 96;
    const long units = (long)a->heads * a->batch, slots = (narrow ? 2L : 1L) * 148;
    int best = 1;
    double best_cost = 1e30;
    for (int qc = 1; qc <= 10 && qc <= q_tiles; ++qc) {
      const long ctas = units * ((q_tiles + qc - 1) / qc);
      const double cost = (double)((ctas + slots - 1) / slots) * (qc + 2.0);
      if (cost < best_cost - 1e-9) { best_cost = cost; best = qc; }
    }
    p.q_chunk = best;
    p.o_tma = (!a->accumulate && a->o_bdiv == 1 && a->o_bs_inner == 0 && a->o_bs_outer % 8 == 0) ? 1 : 0;
    CUtensorMap to = tq;
    if (p.o_tma) {
      pb_attn_operand oo;
      oo.ptr = a->out; oo.bdiv = 1; oo.bs_outer = a->o_bs_outer; oo.bs_inner = 0; oo.rs = a->o_rs;
      if ((rc = encode_rows3d(&to, oo, cols, a->n_q, a->batch)) != PB_OK) return rc;
    }
    dim3 grid((unsigned)ceil_div(q_tiles, best), (unsigned)a->heads, (unsigned)a->batch);
    if (a->kv_split == 77 && a->n_kv == 93)
      tc_attention_x_kernel<3, 77, 93><<<grid, 192, smem_x, to_stream(stream)>>>(tq, tk, tv, to, p);
    else if (narrow) tc_attention_x_kernel<3, 0, 0><<<grid, 192, smem_x, to_stream(stream)>>>(tq, tk, tv, to, p);
    else tc_attention_x_kernel<4, 0, 0><<<grid, 192, smem_x, to_stream(stream)>>>(tq, tk, tv, to, p);
  } else {
    dim3 grid((unsigned)ceil_div(a->n_q, ATT_BM * 2), (unsigned)a->heads, (unsigned)a->batch);
    tc_attention3_kernel<<<grid, 384, smem_v3, to_stream(stream)>>>(tq, tk, tv, p);
  }
  PB_LAUNCH_CHECK("tc_attention_kernel");
  return PB_OK;
#undef ATT_UNSUPPORTED
}
