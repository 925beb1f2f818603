// attention_tc.cu — tcgen05 flash attention, head_dim 64, bf16 operands, fp32 softmax/accumulation.
//
// One CTA per (batch, head, PAIR of 128-query tiles); 320 threads, warp-specialised:
//   warp 0      TMA producer : the two Q tiles once, then K_j / V_j tiles (128 keys x 64) through a 3-stage ring;
//                              K/V are fetched once for both query tiles
//   warp 1      MMA issuer   : S_t,j = Q_t K_j^T (UMMA 128x128x64, both operands K-major, -> TMEM, one buffer per
//                              query tile) and O_t,j = P_t,j V_j (UMMA 128x64x128, A = P from shared memory,
//                              B = V_j used MN-major straight from its row-major [key][d] tile, -> TMEM).
//                              Issue order S_0,j+1 | PV_1,j | S_1,j+1 | PV_0,j+1 ... : while one query tile is in
//                              its softmax the tensor pipe works for the other one (ping-pong).
//   warps 2..5 / 6..9  softmax of query tile 0 / 1: one thread per query row (a TMEM lane): row max over S in
//                              TMEM, fold of the previous O partial into a register accumulator with the
//                              online-softmax correction, then p = exp2(s*scale - m) written to shared memory as
//                              the swizzled bf16 A operand of the PV product.
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
// Two instantiations: <QT=2, ST=3> (two query tiles per CTA in ping-pong, 3-stage K/V ring, 512 TMEM columns, one CTA
// per SM) for self-attention, and <QT=1, ST=1> (80 KB of shared memory, 256 TMEM columns -> two CTAs per SM) for
// launches with a single key tile (cross-attention: per-CTA latency, not throughput, is what matters there)
constexpr int UMMA_K = 16;
constexpr int TILE_BYTES = 128 * 64 * 2;      // 16 KiB: one [128 x 64] bf16 tile
constexpr int P_BYTES = 2 * TILE_BYTES;       // [128 x 128] bf16 as two K-blocks
__host__ __device__ constexpr int att_smem(int QT, int ST) { return TILE_BYTES * (QT + 2 * ST) + QT * P_BYTES; }   // 192 / 80 KiB
constexpr int S_COL0 = 0;                     // TMEM columns: S of query tile t at t*128, O at QT*128 + t*64

struct AttP {
  void* out;
  int64_t o_bdiv, o_bso, o_bsi, o_rs;
  int q_bdiv, k_bdiv, v_bdiv;
  int nq, nkv, kv_tiles;
  float scale_log2;
  int accumulate;
  int split;             // > 0: two softmax segments [0, split) and [split, nkv) inside ONE key tile
};

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

template <int ATT_QT, int KV_STAGES, int TMEM_NCOLS>
__global__ void __launch_bounds__(64 + 128 * ATT_QT, ATT_QT == 1 ? 2 : 1)
tc_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int t) { return base + TILE_BYTES * t; };
  auto sK = [&](int s) { return base + TILE_BYTES * (ATT_QT + s); };
  auto sV = [&](int s) { return base + TILE_BYTES * (ATT_QT + KV_STAGES + s); };
  auto sP = [&](int t) { return base + TILE_BYTES * (ATT_QT + 2 * KV_STAGES) + P_BYTES * t; };
  constexpr int ATT_SMEM = att_smem(ATT_QT, KV_STAGES);
  constexpr int O_COL0 = ATT_QT * ATT_BN;
  const uint32_t bars = base + ATT_SMEM;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + KV_STAGES + s); };
  auto s_full = [&](int t) { return bars + 8u * (1 + 2 * KV_STAGES + t); };
  auto p_full = [&](int t) { return bars + 8u * (3 + 2 * KV_STAGES + t); };
  auto o_full = [&](int t) { return bars + 8u * (5 + 2 * KV_STAGES + t); };
  const uint32_t tmem_slot = bars + 8u * (7 + 2 * KV_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile0 = blockIdx.x * ATT_QT, head = blockIdx.y, n = blockIdx.z;
  const int n_tiles = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int t = 0; t < ATT_QT; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 128); mbar_init(o_full(t), 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_NCOLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_QT * TILE_BYTES);
      for (int t = 0; t < ATT_QT; ++t)      // a tile past the last query row is zero-filled by the copy engine
        tma_load_3d(sQ(t), &tmQ, q_full, head * ATT_D, (q_tile0 + t) * ATT_BM, n / p.q_bdiv);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % KV_STAGES;
        mbar_wait(kv_empty(s), ((j / KV_STAGES) & 1) ^ 1);
        mbar_expect_tx(kv_full(s), 2 * TILE_BYTES);
        tma_load_3d(sK(s), &tmK, kv_full(s), head * ATT_D, j * ATT_BN, n / p.k_bdiv);
        tma_load_3d(sV(s), &tmV, kv_full(s), head * ATT_D, j * ATT_BN, n / p.v_bdiv);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);                 // K-major A and B
      const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);     // B (= V) MN-major
      auto issue_s = [&](int t, int j) {
        const int s = j % KV_STAGES;
        mbar_wait(kv_full(s), (j / KV_STAGES) & 1);
        tc_fence_after();
        const uint64_t q_desc = make_smem_desc_sw128(sQ(t));
        const uint64_t k_desc = make_smem_desc_sw128(sK(s));
        const uint32_t d = tmem_base + S_COL0 + (uint32_t)t * ATT_BN;
#pragma unroll
        for (int k = 0; k < ATT_D / UMMA_K; ++k) umma_bf16(d, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full(t));
      };
      mbar_wait(q_full, 0);
      issue_s(0, 0);
      if (ATT_QT > 1) issue_s(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % KV_STAGES;
#pragma unroll
        for (int t = 0; t < ATT_QT; ++t) {
          // P_t,j is in shared memory (this also means S_t,j has been consumed and O_t,j-1 has been folded)
          mbar_wait(p_full(t), j & 1);
          tc_fence_after();
          const uint32_t d = tmem_base + O_COL0 + (uint32_t)t * ATT_D;
#pragma unroll
          for (int k = 0; k < ATT_BN / UMMA_K; ++k) {
            // A: P block k/4 (16 KiB apart), 32 B per k-step inside the swizzled row; B: 16 key rows = 2 KiB per step
            const uint64_t a_desc = make_smem_desc_sw128(sP(t) + (k >> 2) * TILE_BYTES) + 2 * (k & 3);
            const uint64_t b_desc = make_smem_desc_mn_sw128(sV(s) + k * 16 * 128);
            umma_bf16(d, a_desc, b_desc, idesc_o, k != 0);
          }
          umma_commit(o_full(t));
          if (t == ATT_QT - 1) umma_commit(kv_empty(s));     // K_j / V_j are free once every product on them is done
          if (j + 1 < n_tiles) issue_s(t, j + 1);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + output (warps 2..5: tile 0, 6..9: tile 1)
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + S_COL0 + (uint32_t)t * ATT_BN;
    const uint32_t o_addr = lane_addr + O_COL0 + (uint32_t)t * ATT_D;
    const uint32_t p_row = sP(t) + (uint32_t)r * 128;
    float acc[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) acc[i] = 0.f;
    float m_run = -INFINITY, m_acc = -INFINITY, l = 0.f;

    // TMEM loads are asynchronous until tcgen05.wait::ld: every loop below issues the load of the next chunk
    // before it works on the current one, so the tensor-memory round trip hides behind the ALU / MUFU work
    auto fold_o = [&](int jt, float m_t) {
      mbar_wait(o_full(t), jt & 1);
      tc_fence_after();
      const float corr = fast_exp2(m_acc - m_t);   // 0 on the first tile (m_acc = -inf)
      m_acc = m_t;
      float o0[32], o1[32];
      tmem_ld32(o_addr, o0);
      tmem_ld_wait();
      tmem_ld32(o_addr + 32, o1);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], corr, o0[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[32 + i] = fmaf(acc[32 + i], corr, o1[i]);
    };
    // 32 probabilities (keys c .. c+31 of the tile) -> the swizzled bf16 A operand of the PV product
    auto store_p = [&](int c, const uint32_t (&w)[16]) {
      const uint32_t blk = p_row + (uint32_t)(c >> 6) * TILE_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t chunk = (uint32_t)(((c & 63) >> 3) + q);
        const uint32_t addr = blk + ((chunk ^ (uint32_t)(r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * q]), "r"(w[4 * q + 1]),
                     "r"(w[4 * q + 2]), "r"(w[4 * q + 3])
                     : "memory");
      }
    };

    if (p.split > 0) {
      // ---- two softmax segments in one key tile (text | image tokens): normalised probabilities go into P, the
      // product is the final output
      mbar_wait(s_full(t), 0);
      tc_fence_after();
      const int sp = p.split, nk = p.nkv;
      float mx1 = -INFINITY, mx2 = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATT_BN; c += 32) {
        float v[32];
        tmem_ld32(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mx1 = fmaxf(mx1, (c + i < sp) ? v[i] : -INFINITY);
          mx2 = fmaxf(mx2, (c + i >= sp && c + i < nk) ? v[i] : -INFINITY);
        }
      }
      mx1 *= p.scale_log2; mx2 *= p.scale_log2;
      float l1 = 0.f, l2 = 0.f;
#pragma unroll
      for (int c = 0; c < ATT_BN; c += 32) {
        float v[32];
        tmem_ld32(s_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (c + i < sp) l1 += fast_exp2(fmaf(v[i], p.scale_log2, -mx1));
          else if (c + i < nk) l2 += fast_exp2(fmaf(v[i], p.scale_log2, -mx2));
        }
      }
      const float i1 = 1.0f / l1, i2 = 1.0f / l2;
#pragma unroll
      for (int c = 0; c < ATT_BN; c += 32) {
        float v[32];
        tmem_ld32(s_addr + c, v);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pr[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = c + i + u;
            pr[u] = col < sp ? fast_exp2(fmaf(v[i + u], p.scale_log2, -mx1)) * i1
                  : col < nk ? fast_exp2(fmaf(v[i + u], p.scale_log2, -mx2)) * i2 : 0.f;
          }
          w[i >> 1] = pack_bf16(pr[0], pr[1]);
        }
        store_p(c, w);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full(t));
      m_acc = 0.f;
      fold_o(0, 0.f);           // corr = exp2(0) = 1 on a zero accumulator
      l = 1.f;
    } else {
      float m_prev_tile = -INFINITY;
      for (int j = 0; j < n_tiles; ++j) {
        const int valid = p.nkv - j * ATT_BN;       // keys of this tile that exist (>= 1)
        mbar_wait(s_full(t), j & 1);
        tc_fence_after();
        const bool full = valid >= ATT_BN;          // no key masking needed (all but the last tile)
        float mx;
        {
          // four independent running maxima: a single fmaxf chain over 128 values is 128 dependent instructions
          float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          float v0[32], v1[32];
          auto take = [&](int c, const float (&v)[32]) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], (c + i < valid) ? v[i] : -INFINITY);
            }
          };
          tmem_ld32(s_addr, v0);
          tmem_ld_wait();
          tmem_ld32(s_addr + 32, v1);
          take(0, v0);
          tmem_ld_wait();
          tmem_ld32(s_addr + 64, v0);
          take(32, v1);
          tmem_ld_wait();
          tmem_ld32(s_addr + 96, v1);
          take(64, v0);
          tmem_ld_wait();
          take(96, v1);
          mx = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
        }
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        l *= fast_exp2(m_run - m_new);
        m_run = m_new;
        // the previous product of this query tile is complete: fold it, and P may be overwritten
        if (j > 0) fold_o(j - 1, m_prev_tile);
        float lsum = 0.f;
        {
          float va[32], vb[32];
          float lq[4] = {0.f, 0.f, 0.f, 0.f};       // independent partial sums (no 64-long dependent add chain)
          auto emit = [&](int c, const float (&v)[32]) {
            uint32_t w[16];
            if (full) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = fast_exp2(fmaf(v[i], p.scale_log2, -m_new));
                const float p1 = fast_exp2(fmaf(v[i + 1], p.scale_log2, -m_new));
                lq[(i >> 1) & 3] += p0 + p1;
                w[i >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float p0 = (c + i < valid) ? fast_exp2(fmaf(v[i], p.scale_log2, -m_new)) : 0.f;
                const float p1 = (c + i + 1 < valid) ? fast_exp2(fmaf(v[i + 1], p.scale_log2, -m_new)) : 0.f;
                lq[(i >> 1) & 3] += p0 + p1;
                w[i >> 1] = pack_bf16(p0, p1);
              }
            }
            store_p(c, w);
          };
          tmem_ld32(s_addr, va);
          tmem_ld_wait();
          tmem_ld32(s_addr + 32, vb);
          emit(0, va);
          tmem_ld_wait();
          tmem_ld32(s_addr + 64, va);
          emit(32, vb);
          tmem_ld_wait();
          tmem_ld32(s_addr + 96, vb);
          emit(64, va);
          tmem_ld_wait();
          emit(96, vb);
          lsum = (lq[0] + lq[1]) + (lq[2] + lq[3]);
        }
        l += lsum;
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
        tc_fence_before();
        mbar_arrive(p_full(t));
        m_prev_tile = m_new;
      }
      fold_o(n_tiles - 1, m_prev_tile);
    }

    const int64_t qi = (int64_t)(q_tile0 + t) * ATT_BM + r;
    if (qi < p.nq) {
      const float inv = 1.0f / l;
      bf16* O = reinterpret_cast<bf16*>(p.out) + (n / p.o_bdiv) * p.o_bso + (n % p.o_bdiv) * p.o_bsi + qi * p.o_rs +
                head * ATT_D;
#pragma unroll
      for (int c = 0; c < ATT_D; c += 8) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = acc[c + i] * inv;
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
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_NCOLS);
}

// ================================================================================================
// Version 2 of the self-attention kernel (multi-tile K/V): the FlashAttention-4 data flow.
//   * S_t,j is read from tensor memory ONCE into registers (128 fp32 per thread);
//   * P_t,j goes back into tensor memory (bf16, aliasing the first 64 columns of S_t) with tcgen05.st and is the
//     A operand of the PV product straight from TMEM (no shared-memory round trip, no proxy fence);
//   * O_t accumulates in tensor memory across key tiles (the PV UMMAs accumulate); it is rescaled only when a row
//     maximum grows by more than 2^8 (lazy rescale, decided per warp), otherwise the stale maximum is kept — the
//     probabilities then stay below 2^8, exact in fp32 and harmless in bf16 — so the per-tile fold of O disappears.
// Tensor-pipe ordering does the rest: S_t,j+1 is issued after PV_t,j by the same thread, so P_t,j is consumed before
// the next S overwrites it, and PV_t,j-1 has completed before S_t,j (hence before the softmax of tile j starts).
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

constexpr int V2_STAGES = 4;
constexpr int V2_SMEM = TILE_BYTES * (2 + 2 * V2_STAGES);      // Q x2 + (K, V) x4 = 160 KiB
constexpr float V2_RESCALE_THRESHOLD = 8.0f;                    // log2 units

// 384 threads = three warpgroups: {TMA producer, MMA issuer, two idle warps} | softmax of query tile 0 | softmax of
// query tile 1.  setmaxnreg moves registers from the first warpgroup (88 each) to the softmax groups (208 each), so a
// thread can hold its whole 128-wide score row.
__global__ void __launch_bounds__(384, 1)
tc_attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto sQ = [&](int t) { return base + TILE_BYTES * t; };
  auto sK = [&](int s) { return base + TILE_BYTES * (2 + s); };
  auto sV = [&](int s) { return base + TILE_BYTES * (2 + V2_STAGES + s); };
  const uint32_t bars = base + V2_SMEM;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + V2_STAGES + s); };
  auto s_full = [&](int t) { return bars + 8u * (1 + 2 * V2_STAGES + t); };
  auto p_full = [&](int t) { return bars + 8u * (3 + 2 * V2_STAGES + t); };
  auto o_full = [&](int t) { return bars + 8u * (5 + 2 * V2_STAGES + t); };
  const uint32_t tmem_slot = bars + 8u * (7 + 2 * V2_STAGES);
  constexpr int O_COL0 = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile0 = blockIdx.x * 2, head = blockIdx.y, n = blockIdx.z;
  const int n_tiles = p.kv_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < V2_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int t = 0; t < 2; ++t) { mbar_init(s_full(t), 1); mbar_init(p_full(t), 128); mbar_init(o_full(t), 1); }
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
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        tma_load_3d(sQ(t), &tmQ, q_full, head * ATT_D, (q_tile0 + t) * ATT_BM, n / p.q_bdiv);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % V2_STAGES;
        mbar_wait(kv_empty(s), ((j / V2_STAGES) & 1) ^ 1);
        mbar_expect_tx(kv_full(s), 2 * TILE_BYTES);
        tma_load_3d(sK(s), &tmK, kv_full(s), head * ATT_D, j * ATT_BN, n / p.k_bdiv);
        tma_load_3d(sV(s), &tmV, kv_full(s), head * ATT_D, j * ATT_BN, n / p.v_bdiv);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN);
      const uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D) | (1u << 16);     // B (= V) MN-major
      auto issue_s = [&](int t, int j) {
        const int s = j % V2_STAGES;
        mbar_wait(kv_full(s), (j / V2_STAGES) & 1);
        tc_fence_after();
        const uint64_t q_desc = make_smem_desc_sw128(sQ(t));
        const uint64_t k_desc = make_smem_desc_sw128(sK(s));
        const uint32_t d = tmem_base + (uint32_t)t * ATT_BN;
#pragma unroll
        for (int k = 0; k < ATT_D / UMMA_K; ++k) umma_bf16(d, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(s_full(t));
      };
      mbar_wait(q_full, 0);
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % V2_STAGES;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(p_full(t), j & 1);         // P_t,j is in tensor memory, O_t has been rescaled if it had to be
          tc_fence_after();
          const uint32_t d = tmem_base + O_COL0 + (uint32_t)t * ATT_D;
          const uint32_t a_tm = tmem_base + (uint32_t)t * ATT_BN;      // P aliases the head of S_t: 8 columns per k-step
#pragma unroll
          for (int k = 0; k < ATT_BN / UMMA_K; ++k) {
            const uint64_t b_desc = make_smem_desc_mn_sw128(sV(s) + k * 16 * 128);
            umma_bf16_ts(d, a_tm + (uint32_t)(k * 8), b_desc, idesc_o, (j | k) != 0);
          }
          umma_commit(o_full(t));
          if (t == 1) umma_commit(kv_empty(s));
          if (j + 1 < n_tiles) issue_s(t, j + 1);
        }
      }
    }
  } else if (warp >= 4) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t s_addr = lane_addr + (uint32_t)t * ATT_BN;
    const uint32_t o_addr = lane_addr + O_COL0 + (uint32_t)t * ATT_D;
    float m_run = -INFINITY, l = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = p.nkv - j * ATT_BN;
      const bool full = valid >= ATT_BN;
      mbar_wait(s_full(t), j & 1);
      tc_fence_after();
      // keys 0..63 stay in registers; keys 64..127 are read again for the exponentials (the register file of a
      // 384-thread CTA holds 168 per thread; P only overwrites columns 0..63 of S, so the second half stays intact)
      float v[2][32];
      float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      auto take = [&](int c, const float (&x)[32]) {
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], x[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mq[i & 3] = fmaxf(mq[i & 3], (c + i < valid) ? x[i] : -INFINITY);
        }
      };
      {
        float u[32];
        tmem_ld32(s_addr + 64, u);
        tmem_ld32(s_addr, v[0]);
        tmem_ld_wait();
        take(64, u);
        tmem_ld32(s_addr + 96, u);
        tmem_ld32(s_addr + 32, v[1]);
        take(0, v[0]);
        tmem_ld_wait();
        take(96, u);
        take(32, v[1]);
      }
      const float m_tile = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3])) * p.scale_log2;
      if (j == 0) {
        m_run = m_tile;
      } else {
        const bool need = m_tile > m_run + V2_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {        // warp-uniform: tcgen05.ld / st are warp-collective
          mbar_wait(o_full(t), (j - 1) & 1);        // PV_t,j-1 is complete (it precedes S_t,j in the pipe)
          tc_fence_after();
          const float corr = need ? fast_exp2(m_run - m_tile) : 1.0f;
          if (need) { m_run = m_tile; l *= corr; }
#pragma unroll 1
          for (int c = 0; c < ATT_D; c += 32) {      // 32 columns at a time: the score row is live in registers
            float o0[32];
            tmem_ld32(o_addr + c, o0);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o0[i] *= corr;
            tmem_st32f(o_addr + c, o0);
          }
        }
      }
      float lq[4] = {0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m_run;
      auto emit = [&](int c, const float (&x)[32]) {
        uint32_t w[16];
        if (full) {                                       // no per-element key masking on all but the last tile
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float e0, e1;
            fma2(e0, e1, x[i], x[i + 1], p.scale_log2, neg_m);
            const float p0 = fast_exp2(e0), p1 = fast_exp2(e1);
            add2(lq[(i & 2)], lq[(i & 2) + 1], p0, p1);
            w[i >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (c + i < valid) ? fast_exp2(fmaf(x[i], p.scale_log2, neg_m)) : 0.f;
            const float p1 = (c + i + 1 < valid) ? fast_exp2(fmaf(x[i + 1], p.scale_log2, neg_m)) : 0.f;
            lq[(i >> 1) & 3] += p0 + p1;
            w[i >> 1] = pack_bf16(p0, p1);
          }
        }
        tmem_st16(s_addr + (uint32_t)(c >> 1), w);       // 32 keys = 16 packed columns
      };
      {
        float u[32];
        tmem_ld32(s_addr + 64, u);        // in flight while the first half is exponentiated
        emit(0, v[0]);
        emit(32, v[1]);
        tmem_ld_wait();
        float u2[32];
        tmem_ld32(s_addr + 96, u2);
        emit(64, u);
        tmem_ld_wait();
        emit(96, u2);
      }
      l += (lq[0] + lq[1]) + (lq[2] + lq[3]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(t));
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
  auto k_self = tc_attention_kernel<2, 3, 512>;
  auto k_one = tc_attention_kernel<1, 1, 256>;
  const size_t smem_self = att_smem(2, 3) + 1024 + 8 * (8 + 2 * 3) + 16;
  const size_t smem_one = att_smem(1, 1) + 1024 + 8 * (8 + 2 * 1) + 16;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_self, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_self);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_one);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_attention_kernel)");
    configured = true;
  }
  static const bool v2_on = []() { const char* e = getenv("PB_ATT_V2"); return !(e && e[0] == '0'); }();
  static bool v2_configured = false;
  const size_t smem_v2 = V2_SMEM + 1024 + 8 * (8 + 2 * V2_STAGES) + 16;
  if (v2_on && !v2_configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v2);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_attention2_kernel)");
    v2_configured = true;
  }
  if (p.kv_tiles == 1) {
    dim3 grid((unsigned)ceil_div(a->n_q, ATT_BM), (unsigned)a->heads, (unsigned)a->batch);
    k_one<<<grid, 192, smem_one, to_stream(stream)>>>(tq, tk, tv, p);
  } else if (v2_on) {
    dim3 grid((unsigned)ceil_div(a->n_q, ATT_BM * 2), (unsigned)a->heads, (unsigned)a->batch);
    tc_attention2_kernel<<<grid, 384, smem_v2, to_stream(stream)>>>(tq, tk, tv, p);
  } else {
    dim3 grid((unsigned)ceil_div(a->n_q, ATT_BM * 2), (unsigned)a->heads, (unsigned)a->batch);
    k_self<<<grid, 320, smem_self, to_stream(stream)>>>(tq, tk, tv, p);
  }
  PB_LAUNCH_CHECK("tc_attention_kernel");
  return PB_OK;
#undef ATT_UNSUPPORTED
}
