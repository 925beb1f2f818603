// gemm_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM for the bf16 hot path (sm_100a).
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      A, W bf16; fp32 accumulation in tensor memory
//
// One persistent CTA per SM, 320 threads, warp-specialised:
//   warp 0      TMA producer   : cp.async.bulk.tensor tiles of A (128 rows x 64 k) and W (block_n x 64 k)
//                                into a ring of 128B-swizzled shared-memory stages (mbarrier full/empty)
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (M=128, N=block_n, K=16) x4 per
//                                stage; tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..9  epilogue       : two teams of four warps (a team covers the four TMEM lane quarters); a team
//                                takes every other 32-column chunk of the 128 x block_n fp32 accumulator:
//                                tcgen05.ld (thread = row) -> +bias / GEGLU / +row-bias / +residual -> swizzled
//                                shared-memory slot -> ONE bulk tensor store per chunk.  The residual tile
//                                is fetched by a bulk tensor LOAD into the same slot one chunk ahead, so the
//                                epilogue issues no per-thread global loads or stores at all: the K=320
//                                layers are bound by HBM, and memory-level parallelism comes from the copy
//                                engine instead of 128 threads x 8 loads.
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// The A operand and the C / residual tiles are addressed purely by TMA coordinates, so convolution needs
// no im2col buffer and no row-index arithmetic:
//   LINEAR   A: 2-D map [M][K]                       C: [splits][M][N]
//   CONV3X3  A: 4-D map [n][y][x][c]; tap (ky,kx) = the same box shifted by (ky-1, kx-1); out-of-bounds
//            rows/columns are zero-filled by the TMA unit, which IS the conv's zero padding
//                                                    C: [n][y][x][cout], same brick
//   CONVT3   A: 4-D map [b][t][p][c]; tap kt = box shifted by kt-1 frames (zero-filled outside the clip)
// A 128-row tile is a BW x BH x BN (x, y, image) or BP x BT (pixel, frame) brick chosen so the brick tiles
// the tensor exactly.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pb {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
// epilogue teams of four warps (a team covers the four TMEM lane quarters): two, or four for the GEGLU kernels whose
// epilogue evaluates an erf per accumulator element (with K = 320 that, not the tensor pipe, set the pace)
constexpr int MAX_TEAMS = 4;
__host__ __device__ constexpr int num_teams(bool geglu) { return geglu ? 4 : 2; }
__host__ __device__ constexpr int num_threads(bool geglu) { return 64 + 128 * num_teams(geglu); }     // 320 / 576
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 8;
constexpr int EPI_CHUNK = 32;                           // accumulator columns per epilogue chunk
constexpr int EPI_SLOTS = 2;                            // staging slots per team (default)
constexpr int MAX_SLOTS = 3;                            // short-K tiles with a residual: 3 slots, residual loads 2 chunks ahead
constexpr int EPI_SLOT_F32 = BLOCK_M * EPI_CHUNK * 4;   // 16 KiB: [128][32] fp32, 128B-swizzled rows
constexpr int EPI_SLOT_BF16 = BLOCK_M * EPI_CHUNK * 2;  //  8 KiB: [128][32] bf16,  64B-swizzled rows
constexpr int BAR_BYTES = 8 * (2 * MAX_STAGES + 4 + MAX_TEAMS * MAX_SLOTS) + 16;
constexpr int SMEM_TOTAL = 227 * 1024;

struct TcParams {
  const float* bias;
  const float* rowbias;
  float* colstats;       // optional [ceil(M/32)][ld_cs] fp32: (sum, sumsq) per output column over each 32-row slab
  int64_t ld_cs;
  int64_t M, ldrb, rb_rows;
  int N;                 // accumulator columns in total (weight rows)
  int block_n;           // accumulator columns per tile
  int n_tiles, m_tiles;
  int k_chunks;          // Cin / 64 (per tap) or K / 64
  int taps;              // 1, 9 or 3
  int stages;
  int epi;
  int geglu_half;        // N / 2 when GEGLU
  int slot_bytes;        // epilogue staging slot size
  int epi_slots;         // staging slots per epilogue team (2, or 1 for long-K tiles: the bytes go to the TMA ring)
  // brick geometry (conv modes)
  int bw, bh, bn_img;    // CONV3X3: brick x, y, images ; CONVT3: bw = pixels, bh = frames
  int W, H;              // CONV3X3: image size ; CONVT3: W = HW (pixels per frame), H = T
  int tiles_x, tiles_y;  // CONV3X3: W/bw, H/bh ; CONVT3: HW/bw, T/bh
  int n_outer;           // images (CONV3X3) / samples (CONVT3): the C map's outermost extent per split
  int cstride;           // CONV3X3 spatial stride (1 or 2); W/H above are OUTPUT sizes
  int splits;            // split-K ways (1 = off); split s covers k-iterations [s*kps, (s+1)*kps)
  int kps;               // k-iterations per split
  int pair;              // 1 = CTA-pair kernel (cta_group::2): block_n is the width of the 256-row pair tile
};

template <int MODE>
__device__ __forceinline__ int64_t out_row(const TcParams& p, int m_tile, int r) {
  if (MODE == PB_A_LINEAR) return (int64_t)m_tile * BLOCK_M + r;
  const int tx = m_tile % p.tiles_x;
  const int ty = (m_tile / p.tiles_x) % p.tiles_y;
  const int tn = m_tile / (p.tiles_x * p.tiles_y);
  const int bx = r % p.bw;
  const int by = (r / p.bw) % p.bh;
  const int bi = r / (p.bw * p.bh);
  // CONV3X3: ((img*H + y)*W + x) ; CONVT3 (bn_img = 1): ((b*T + t)*HW + pix)
  return ((int64_t)(tn * p.bn_img + bi) * p.H + (ty * p.bh + by)) * p.W + (tx * p.bw + bx);
}

// ---- bulk tensor store / group management (epilogue)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void team_bar_sync(int team) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// value * gelu(gate) for a PAIR of outputs, gelu(g) = g * Phi(g) with the exact (erf) Phi of F.gelu:
// Phi(-t) = 2^q(t) for t = min(|g|, 9), q a degree-6 polynomial fitted to log2(erfc(t / sqrt 2) / 2) — max |error| of
// Phi 3.4e-7, of gelu 1.3e-7 (the Abramowitz-Stegun form this replaces: 1.5e-7) — and Phi(g) = 1 - Phi(-g) for g > 0.
// One MUFU.EX2 and six packed FMAs per pair instead of RCP + EX2 + 19 scalar instructions per element: the GEGLU
// epilogue at K = 320 was issue-bound (ncu: issue slots 70 % busy).  Results are rounded to bf16 right away.
__device__ __forceinline__ void geglu_pair(float a0, float a1, float g0, float g1, float& o0, float& o1) {
  const float t0 = fminf(fabsf(g0), 9.0f), t1 = fminf(fabsf(g1), 9.0f);
  float q0, q1;
  asm("{\n.reg .b64 t, q, c;\n"
      "mov.b64 t, {%2, %3};\n"
      "mov.b64 q, {%4, %4};\n"
      "mov.b64 c, {%5, %5};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 c, {%6, %6};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 c, {%7, %7};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 c, {%8, %8};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 c, {%9, %9};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 c, {%10, %10};\n"
      "fma.rn.f32x2 q, q, t, c;\n"
      "mov.b64 {%0, %1}, q;\n}"
      : "=f"(q0), "=f"(q1)
      : "f"(t0), "f"(t1), "f"(2.905618202930782e-05f), "f"(-0.0007324367179535329f), "f"(0.007954133674502373f),
        "f"(-0.053205594420433044f), "f"(-0.4589345157146454f), "f"(-1.1511449813842773f), "f"(-0.9999991059303284f));
  float p0, p1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(q1));
  const float phi0 = g0 < 0.f ? p0 : 1.0f - p0, phi1 = g1 < 0.f ? p1 : 1.0f - p1;
  o0 = a0 * (g0 * phi0);
  o1 = a1 * (g1 * phi1);
}

// ================================================================================================ epilogue
// Runs in warps 2..9.  first/step/total describe this CTA's work list (tiles x splits, or pair tiles);
// `rank` is the CTA rank inside a pair (0 otherwise).  tempty_addr(acc) is the (possibly remote) barrier
// that hands the accumulator buffer back to the MMA issuer.
template <int MODE, bool GEGLU, typename TC, bool PAIR>
__device__ __forceinline__ void epilogue_loop(const CUtensorMap* tmC, const CUtensorMap* tmR, const TcParams& p,
                                              uint32_t tmem_base, uint32_t epi_base, uint32_t tfull0, uint32_t tempty0,
                                              uint32_t rfull0, int first, int step, int total, int rank) {
  constexpr int NT = num_teams(GEGLU);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int team = (warp - 2) >> 2;
  const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
  const int r = quarter * 32 + lane;       // accumulator row owned by this thread
  const bool leader = (warp - 2) == team * 4 && lane == 0;
  const bool RES = (p.epi & PB_EPI_RESIDUAL) != 0;
  constexpr bool OUT_BF16 = sizeof(TC) == 2;
  const int out_cols = GEGLU ? (p.block_n >> 1) : p.block_n;
  const int n_chunks = (out_cols + EPI_CHUNK - 1) / EPI_CHUNK;
  const int n_slots = p.epi_slots;
  const uint32_t slot_base = epi_base + (uint32_t)(team * n_slots * p.slot_bytes);
  auto rfull = [&](int slot) { return rfull0 + 8u * (uint32_t)(team * MAX_SLOTS + slot); };
  // no residual: slot cnt % S is rewritten once the store of chunk cnt - S has read it (S - 1 newer stores may be pending)
  auto wait_slot_drained = [&]() {
    if (n_slots == 1) bulk_wait_read<0>(); else if (n_slots == 2) bulk_wait_read<1>(); else bulk_wait_read<2>();
  };
  // (work, chunk) -> tensor coordinates of the 128 x 32 output box
  auto coords = [&](int work, int ci, int (&c)[4], int& m_tile, int& n_tile) {
    int split = 0;
    if (PAIR) {          // work = (pair m-tile, n-tile); this CTA owns rows [rank*128, rank*128 + 128) of the pair tile
      n_tile = work % p.n_tiles;
      m_tile = 2 * (work / p.n_tiles) + rank;
    } else {
      const int tile = work / p.splits;
      split = work - tile * p.splits;
      n_tile = tile % p.n_tiles;
      m_tile = tile / p.n_tiles;
    }
    c[0] = n_tile * out_cols + ci * EPI_CHUNK;
    if (MODE == PB_A_LINEAR) {
      c[1] = m_tile * BLOCK_M; c[2] = split; c[3] = 0;
    } else {
      c[1] = (m_tile % p.tiles_x) * p.bw;
      c[2] = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
      c[3] = (m_tile / (p.tiles_x * p.tiles_y)) * p.bn_img + split * p.n_outer;
    }
  };

  // residual: the leader runs a prefetch stream `pf_dist` chunks ahead of the chunk being processed (2 with three slots:
  // ncu showed the epilogue warps of the HBM-bound K = 320 layers waiting on a one-chunk-ahead load 17 % of the time).
  // The slot of chunk cnt + D last held chunk cnt + D - S, so S - D newer stores (= 1 for S = 2, 3) may still be pending.
  const int pf_dist = n_slots >= 3 ? 2 : 1;
  int pf_work = first, pf_ci = team;
  uint32_t pf_cnt = 0;
  // (The leader is ONE divergent lane on purpose: run warp-uniform with an elected lane, as the producer / issuer
  // loops are, the other 31 lanes of the leader warp wait for its bulk-group drains instead of starting the next
  // chunk's tensor-memory loads — measured 3-5 % slower on the Linear / GEGLU family.)
  auto pf_issue = [&]() {
    if (pf_work >= total || team >= n_chunks) return;
    int nc[4], nmt, nnt;
    coords(pf_work, pf_ci, nc, nmt, nnt);
    const int sl = (int)(pf_cnt % (uint32_t)n_slots);
    mbar_expect_tx(rfull(sl), (uint32_t)EPI_SLOT_F32);
    tma_load_4d(slot_base + (uint32_t)(sl * p.slot_bytes), tmR, rfull(sl), nc[0], nc[1], nc[2], nc[3]);
    ++pf_cnt;
    pf_ci += NT;
    if (pf_ci >= n_chunks) { pf_ci = team; pf_work += step; }
  };

  int acc = 0, acc_phase = 0;
  uint32_t cnt = 0;          // chunks this team has processed: slot = cnt % SLOTS, parity = (cnt / SLOTS) & 1
  if (RES && leader) {
    for (int i = 0; i < pf_dist; ++i) pf_issue();
  }
  for (int work = first; work < total; work += step) {
    int c[4], m_tile, n_tile;
    coords(work, 0, c, m_tile, n_tile);
    const int64_t n0 = (int64_t)n_tile * out_cols;       // first output column of the tile
    const float* rb_ptr = nullptr;
    if (p.epi & PB_EPI_ROWBIAS) {
      int64_t m = out_row<MODE>(p, m_tile, r);
      if (m >= p.M) m = p.M - 1;
      rb_ptr = p.rowbias + (m / p.rb_rows) * p.ldrb + n0;
    }
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
    for (int ci = team; ci < n_chunks; ci += NT) {
      const int c0 = ci * EPI_CHUNK;
      const int cw = out_cols - c0 < EPI_CHUNK ? out_cols - c0 : EPI_CHUNK;   // 32, or 16 on a ragged last chunk
      const int slot = (int)(cnt % (uint32_t)n_slots);
      const uint32_t sphase = (cnt / (uint32_t)n_slots) & 1u;
      const uint32_t sbuf = slot_base + (uint32_t)(slot * p.slot_bytes);
      // ---- (1) TMEM -> registers (+bias, GEGLU)
      float v[EPI_CHUNK];
#pragma unroll
      for (int h = 0; h < EPI_CHUNK; h += 16) {
        if (h < cw) {
          float a16[16];
          tmem_ld16(t_row + (uint32_t)(c0 + h), a16);
          float g16[16];
          if (GEGLU) tmem_ld16(t_row + (uint32_t)(out_cols + c0 + h), g16);
          tmem_ld_wait();
          if (p.epi & PB_EPI_BIAS) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + h + j));
              a16[j] += b4.x; a16[j + 1] += b4.y; a16[j + 2] += b4.z; a16[j + 3] += b4.w;
              if (GEGLU) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.bias + p.geglu_half + n0 + c0 + h + j));
                g16[j] += g4.x; g16[j + 1] += g4.y; g16[j + 2] += g4.z; g16[j + 3] += g4.w;
              }
            }
          }
          if (GEGLU) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) geglu_pair(a16[j], a16[j + 1], g16[j], g16[j + 1], v[h + j], v[h + j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[h + j] = a16[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[h + j] = 0.f;
        }
      }
      if (ci + NT >= n_chunks) {
        // this team's last read of the accumulator buffer: hand it back to the MMA warp early
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((tempty0 + 8u * acc) & 0xFEFFFFFFu) : "memory");
          else mbar_arrive(tempty0 + 8u * acc);
        }
      }
      if (rb_ptr != nullptr) {
#pragma unroll
        for (int j = 0; j < EPI_CHUNK; j += 4) {
          if (j < cw) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(rb_ptr + c0 + j));
            v[j] += u.x; v[j + 1] += u.y; v[j + 2] += u.z; v[j + 3] += u.w;
          }
        }
      }
      // ---- (2) the slot is free (no residual) / holds the residual tile (bulk load issued one chunk ago)
      if (RES) {
        mbar_wait(rfull(slot), sphase);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t4[4];
          lds128(sbuf + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)), t4);
          v[4 * j] += t4[0]; v[4 * j + 1] += t4[1]; v[4 * j + 2] += t4[2]; v[4 * j + 3] += t4[3];
        }
        if (OUT_BF16) team_bar_sync(team);   // rows are repacked to 64 B: all reads before any write
      } else {
        if (leader) wait_slot_drained();
        team_bar_sync(team);
      }
      // ---- (3) registers -> swizzled staging slot
      if (OUT_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(sbuf + (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)), pack2_bf16(v[8 * j], v[8 * j + 1]),
                  pack2_bf16(v[8 * j + 2], v[8 * j + 3]), pack2_bf16(v[8 * j + 4], v[8 * j + 5]),
                  pack2_bf16(v[8 * j + 6], v[8 * j + 7]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(sbuf + (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      // ---- (4) one bulk tensor store per chunk; prefetch the residual of this team's next chunk
      fence_proxy_async();
      team_bar_sync(team);
      // ---- (4b) GroupNorm column statistics of the tile just staged: warp w of the team owns the 32-row slab w,
      // lane = column; rows of a swizzled slot are read conflict-free.  The bulk store reads the slot concurrently;
      // the slot is not rewritten before this team's next barrier.  For a bf16 output the statistics are those of
      // the ROUNDED values (what the norm that consumes the tensor will actually see).
      if (p.colstats != nullptr) {
        const int w4 = (warp - 2) & 3;
        const int64_t m0 = out_row<MODE>(p, m_tile, w4 * 32);
        float cs_s = 0.f, cs_q = 0.f;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          const int row = w4 * 32 + i;
          float t;
          if (OUT_BF16) {
            unsigned short h;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h)
                         : "r"(sbuf + (uint32_t)(row * 64 + ((((lane >> 3) ^ ((row >> 1) & 3))) << 4) + ((lane & 7) << 1))));
            t = __uint_as_float((uint32_t)h << 16);
          } else {
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t)
                         : "r"(sbuf + (uint32_t)(row * 128 + ((((lane >> 2) ^ (row & 7))) << 4) + ((lane & 3) << 2))));
          }
          cs_s += t;
          cs_q = fmaf(t, t, cs_q);
        }
        if (lane < cw && m0 < p.M)
          *reinterpret_cast<float2*>(p.colstats + (m0 >> 5) * p.ld_cs + 2 * (n0 + c0 + lane)) = make_float2(cs_s, cs_q);
      }
      if (leader) {
        coords(work, ci, c, m_tile, n_tile);
        tma_store_4d(tmC, sbuf, c[0], c[1], c[2], c[3]);
        bulk_commit();
        if (RES && pf_work < total) {
          if (n_slots == 1) bulk_wait_read<0>(); else bulk_wait_read<1>();   // the store that last used that slot has drained it
          pf_issue();
        }
      }
      ++cnt;
    }
    if (team >= n_chunks) {      // single-chunk tiles: the idle team still frees the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"((tempty0 + 8u * acc) & 0xFEFFFFFFu) : "memory");
        else mbar_arrive(tempty0 + 8u * acc);
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
  if (leader) bulk_wait_all();     // shared memory must outlive the last bulk stores
}

template <int MODE, bool GEGLU, typename TC>
__global__ void __launch_bounds__(num_threads(GEGLU), 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  pb::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-byte alignment (swizzle-128B atoms)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const uint32_t epi_base = smem_base + p.stages * stage_bytes;          // teams x epi_slots staging slots
  const uint32_t bar_base = epi_base + num_teams(GEGLU) * p.epi_slots * p.slot_bytes;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2], rfull[2*SLOTS]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t rfull0 = bar_base + 8u * (2 * MAX_STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4 + MAX_TEAMS * MAX_SLOTS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (p.epi & PB_EPI_RESIDUAL) prefetch_tmap(&tmR);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * num_teams(GEGLU)); }
    for (int s = 0; s < MAX_TEAMS * MAX_SLOTS; ++s) mbar_init(rfull0 + 8u * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();               // everything above ran under the predecessor's tail; global memory is touched from here on

  if (warp == 0) {
    // ================================================================== TMA producer (uniform loop, elected issue)
    {
      int stage = 0, phase = 0;
      for (int work = blockIdx.x; work < total_tiles * p.splits; work += gridDim.x) {
        const int tile = work / p.splits, split = work - tile * p.splits;
        const int it0 = split * p.kps, it1 = (it0 + p.kps < k_iters) ? it0 + p.kps : k_iters;
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        int cx = 0, cy = 0, cn = 0;
        if (MODE != PB_A_LINEAR) {
          cx = (m_tile % p.tiles_x) * p.bw;
          cy = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
          cn = (m_tile / (p.tiles_x * p.tiles_y)) * p.bn_img;
        }
        for (int it = it0; it < it1; ++it) {
          const int tap = it / p.k_chunks, kc = it - tap * p.k_chunks;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + A_STAGE_BYTES;
            mbar_expect_tx(full_bar(stage), (uint32_t)stage_bytes);
            if (MODE == PB_A_LINEAR) {
              tma_load_2d(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, m_tile * BLOCK_M);
            } else if (MODE == PB_A_CONV3X3) {
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, cx * p.cstride + kx - 1, cy * p.cstride + ky - 1, cn);
            } else {
              tma_load_4d(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, cx, cy + tap - 1, cn);
            }
            const int kcol = it * BLOCK_K;   // weight column = tap*Cin + kc*64
            if (GEGLU) {
              const int half = p.block_n >> 1;
              tma_load_2d(b_dst, &tmB, full_bar(stage), kcol, n_tile * half);
              tma_load_2d(b_dst + half * BLOCK_K * 2, &tmB, full_bar(stage), kcol, p.geglu_half + n_tile * half);
            } else {
              tma_load_2d(b_dst, &tmB, full_bar(stage), kcol, n_tile * p.block_n);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    // The whole warp walks the loop in uniform control flow (`warp` comes from a shuffle) and one elected lane
    // issues: under `if (lane == 0)` ptxas wraps EVERY tcgen05.mma in an election loop (ELECT / R2UR.BROADCAST /
    // BRA.U.ANY, ~100 clk per MMA measured in the attention kernel), which paces the tensor pipe by the issuer.
    {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, (uint32_t)p.block_n);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int work = blockIdx.x; work < total_tiles * p.splits; work += gridDim.x) {
        const int split = work % p.splits;
        const int it0 = split * p.kps, it1 = (it0 + p.kps < k_iters) ? it0 + p.kps : k_iters;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int it = it0; it < it1; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + stage * stage_bytes;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 16 elements (32 bytes) along K inside the swizzled row: +2 in the (addr >> 4) field
              umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, ((it - it0) | k) != 0);
            }
            umma_commit(empty_bar(stage));
            if (it + 1 == it1) umma_commit(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    epilogue_loop<MODE, GEGLU, TC, false>(&tmC, &tmR, p, tmem_base, epi_base, tfull_bar(0), tempty_bar(0), rfull0,
                                          (int)blockIdx.x, (int)gridDim.x, total_tiles * p.splits, 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================
// CTA-pair variant (cta_group::2): two SMs of a TPC compute one 256 x block_n tile.  Each CTA stages ITS
// 128 rows of A and ITS half (block_n/2 rows) of W, the leader CTA issues one UMMA M=256 that reads both
// shared memories, and each CTA's tensor memory receives its own 128 accumulator rows.  Per SM this moves
// 16 KiB + block_n*64 B per k-step instead of 16 KiB + block_n*128 B, so the same shared-memory ring covers
// 1.5x more FLOPs of latency and L2 traffic per FLOP drops by a third.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z0 = 0, z1 = 0, z2 = 0, z3 = 0, z4 = 0, z5 = 0, z6 = 0, z7 = 0;   // disable-output-lane masks
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z0), "r"(z1), "r"(z2), "r"(z3),
      "r"(z4), "r"(z5), "r"(z6), "r"(z7)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {   // arrives on `bar` of every CTA in mask
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int MODE, bool GEGLU, typename TC>
__global__ void __launch_bounds__(num_threads(GEGLU), 1)      // cluster of 2 CTAs (launch attribute)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const TcParams p) {
  pb::pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int half_n = p.block_n >> 1;                        // W rows staged by this CTA
  const int stage_bytes = A_STAGE_BYTES + half_n * BLOCK_K * 2;
  const uint32_t epi_base = smem_base + p.stages * stage_bytes;
  const uint32_t bar_base = epi_base + num_teams(GEGLU) * p.epi_slots * p.slot_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };       // one per CTA (multicast commit)
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };   // one per CTA (multicast commit)
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };  // leader only, 16 arrivals
  const uint32_t rfull0 = bar_base + 8u * (2 * MAX_STAGES + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4 + MAX_TEAMS * MAX_SLOTS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);     // warp-uniform for ptxas
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // CTA inside its pair: which 128 of the 256 rows
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int pair_m_tiles = (p.m_tiles + 1) >> 1;
  const int total_work = pair_m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (p.epi & PB_EPI_RESIDUAL) prefetch_tmap(&tmR);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8 * num_teams(GEGLU)); }
    for (int s = 0; s < MAX_TEAMS * MAX_SLOTS; ++s) mbar_init(rfull0 + 8u * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();               // everything above ran under the predecessor's tail; global memory is touched from here on

  if (warp == 0) {
    // ================================================================== TMA producer (both CTAs; uniform loop, elected issue)
    {
      int stage = 0, phase = 0;
      for (int work = pair_id; work < total_work; work += num_pairs) {
        const int n_tile = work % p.n_tiles;
        const int m_tile = 2 * (work / p.n_tiles) + (int)rank;
        int cx = 0, cy = 0, cn = 0;
        if (MODE != PB_A_LINEAR) {
          cx = (m_tile % p.tiles_x) * p.bw;
          cy = ((m_tile / p.tiles_x) % p.tiles_y) * p.bh;
          cn = (m_tile / (p.tiles_x * p.tiles_y)) * p.bn_img;
        }
        for (int it = 0; it < k_iters; ++it) {
          const int tap = it / p.k_chunks, kc = it - tap * p.k_chunks;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (elect_one()) {
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + A_STAGE_BYTES;
            if (leader) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)stage_bytes);   // bytes of BOTH CTAs
            if (MODE == PB_A_LINEAR) {
              tma_load_2d_pair(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, m_tile * BLOCK_M);
            } else if (MODE == PB_A_CONV3X3) {
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d_pair(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, cx * p.cstride + kx - 1,
                               cy * p.cstride + ky - 1, cn);
            } else {
              tma_load_4d_pair(a_dst, &tmA, full_bar(stage), kc * BLOCK_K, cx, cy + tap - 1, cn);
            }
            const int kcol = it * BLOCK_K;
            // GEGLU: CTA 0 stages the value rows, CTA 1 the gate rows of the same output columns
            const int wrow = GEGLU ? (int)rank * p.geglu_half + n_tile * half_n : n_tile * p.block_n + (int)rank * half_n;
            tma_load_2d_pair(b_dst, &tmB, full_bar(stage), kcol, wrow);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader CTA only)
    if (leader) {                                      // uniform: whole warp loops, one elected lane issues
      const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, (uint32_t)p.block_n);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int work = pair_id; work < total_work; work += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + stage * stage_bytes;
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (it | k) != 0);
            umma_commit_pair(empty_bar(stage), 3);      // both CTAs of the pair may refill this stage
            if (it + 1 == k_iters) umma_commit_pair(tfull_bar(acc), 3);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================== epilogue (warps 2..9, both CTAs)
    epilogue_loop<MODE, GEGLU, TC, true>(&tmC, &tmR, p, tmem_base, epi_base, tfull_bar(0), tempty_bar(0), rfull0,
                                         pair_id, num_pairs, total_work, (int)rank);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody frees tensor memory / exits while the peer may still touch it
  if (warp == 1) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

// split-K fold: C = sum_s partial[s] (+bias +rowbias +residual), deterministic order
template <typename TC>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t M, int N, TC* __restrict__ c, int64_t ldc,
                     const float* __restrict__ bias, const float* __restrict__ rowbias, int64_t ldrb, int64_t rb_rows,
                     const float* __restrict__ residual, int64_t ldr) {
  pb::pdl_trigger();
  const int64_t nv = N >> 2, total = M * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / nv;
    const int n = (int)(i - m * nv) * 4;
    float4 a = *reinterpret_cast<const float4*>(ws + m * N + n);
    for (int s = 1; s < splits; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(ws + ((int64_t)s * M + m) * N + n);
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    if (bias) { const float4 t = __ldg(reinterpret_cast<const float4*>(bias + n)); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    if (rowbias) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(rowbias + (m / rb_rows) * ldrb + n));
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    if (residual) { const float4 t = *reinterpret_cast<const float4*>(residual + m * ldr + n); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
    const float o[4] = {a.x, a.y, a.z, a.w};
    store4(c + m * ldc + n, o);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// Encoded tensor maps are cached per thread, keyed by everything that goes into the encoding (base pointer, shape,
// strides, box, element strides, dtype, swizzle): the caching allocator hands the same addresses back forward after
// forward, so an eager forward pays cuTensorMapEncodeTiled once per distinct operand instead of 4x per GEMM and 3-4x
// per attention launch.  Direct-mapped, 1024 entries; a miss simply re-encodes.
namespace {
struct TmapKey {
  const void* base;
  uint64_t dims[5], strides[4];
  uint32_t box[5], estr[5];
  int32_t dtype, swizzle, rank, pad;
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int TMAP_CACHE = 1024;
}  // namespace

int encode_tmap(CUtensorMap* out, int dtype, int swizzle_bytes, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  static thread_local TmapEntry* cache = nullptr;
  if (cache == nullptr) cache = static_cast<TmapEntry*>(calloc(TMAP_CACHE, sizeof(TmapEntry)));
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.dtype = dtype; key.swizzle = swizzle_bytes; key.rank = rank;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; key.estr[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  uint64_t h = 1469598103934665603ull;
  const uint64_t* kw = reinterpret_cast<const uint64_t*>(&key);
  for (size_t i = 0; i < sizeof(key) / 8; ++i) { h ^= kw[i]; h *= 1099511628211ull; }
  TmapEntry* slot = cache ? &cache[(h ^ (h >> 29)) & (TMAP_CACHE - 1)] : nullptr;
  if (slot && slot->valid && memcmp(&slot->key, &key, sizeof(key)) == 0) {
    *out = slot->map;
    return PB_OK;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PB_ECUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, dtype == PB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return PB_ECUDA;
  }
  if (slot) { slot->key = key; slot->map = *out; slot->valid = true; }
  return PB_OK;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return encode_tmap(out, PB_BF16, 128, base, rank, dims, strides_bytes, box, elem_strides);
}

// 128-row brick bw x bh x bn (bw*bh*bn = 128, each extent a power of two no larger than the tensor's).  Exact
// tilings (bw | W, bh | H, bn | N) are preferred, widest bw first (long contiguous runs for the TMA unit).  When
// none exists — 9x16 images with 4 frames per rank, 36 pixels per rank under sequence parallelism — the brick that
// needs the fewest tiles is taken and the ragged edge tiles hang over the tensor: the TMA unit zero-fills the rows it
// loads from outside and clips the rows it stores there, so the tile arithmetic is unchanged (the hang-over is wasted
// tensor work, still an order of magnitude faster than the CUDA-core path).
static bool pick_brick(int W, int H, int N, int* bw, int* bh, int* bn, bool* exact) {
  long best_tiles = -1;
  bool best_exact = false;
  for (int w = BLOCK_M; w >= 1; w >>= 1) {
    if (w > W) continue;
    for (int h = BLOCK_M / w; h >= 1; h >>= 1) {
      if (h > H) continue;
      const int n = BLOCK_M / (w * h);
      if (n > N) continue;
      const bool ex = W % w == 0 && H % h == 0 && N % n == 0;
      const long tiles = (long)((W + w - 1) / w) * ((H + h - 1) / h) * ((N + n - 1) / n);
      // exact beats inexact; then fewer tiles; then (loop order) wider w, taller h
      if (best_tiles < 0 || (ex && !best_exact) || (ex == best_exact && tiles < best_tiles)) {
        best_tiles = tiles; best_exact = ex; *bw = w; *bh = h; *bn = n;
      }
    }
  }
  if (exact) *exact = best_exact;
  return best_tiles > 0;
}

static int num_sms();

// Tile width: a multiple of 32 (the epilogue's chunk) dividing n_cols, <= cap: the widest tile unless a narrower
// one removes a whole wave over the SMs; cost = waves x (width + 360) in "column" units.  (The constant dates from
// round 1, when an election loop around every tcgen05.mma made a k-step cost the same for every width; with the
// uniform issuer a sweep of the constant over 48 .. 360 changes the choice on two shapes only and the total of the
// 24 GEMM-family shapes of a forward by < 2 %: the widths that divide 320 / 640 / 1280 leave little to choose.)
// A width that is only a multiple of 16 is legal when it is the whole output (its ragged last chunk is clipped
// at the tensor edge by the bulk store).
static int pick_block_n(int n_cols, int cap, int m_tiles) {
  if (n_cols <= cap && n_cols % 16 == 0 && n_cols % 32 != 0) return n_cols;
#ifdef PB_TC_TUNING          // profiling build only: tile-width sweeps (tools/tile_sweep.py)
  if (const char* f = getenv("PB_TC_FORCE_BN")) {
    const int bn = atoi(f);
    if (bn >= 32 && bn <= cap && n_cols % bn == 0) return bn;
  }
#endif
  int best = 0;
  long best_cost = 0;
  for (int bn = cap - cap % 32; bn >= 64 || (best == 0 && bn >= 32); bn -= 32) {
    if (n_cols % bn != 0) continue;
    const long tiles = (long)m_tiles * (n_cols / bn);
    const long waves = (tiles + num_sms() - 1) / num_sms();
    const long cost = waves * (bn + 360);
    if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

// split-K ways for launches that would leave most SMs idle (H/8 level: 5 m-tiles)
static int pick_splits(int tiles, int k_iters) {
  if (tiles * 2 > num_sms() || k_iters < 8) return 1;
  int s = num_sms() / tiles;
  if (s > k_iters / 4) s = k_iters / 4;
  return s < 1 ? 1 : s;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE, bool GEGLU, typename TC>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                  const TcParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;
  auto kern = tc_gemm_kernel<MODE, GEGLU, TC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_kernel)");
    configured = true;
  }
  int grid = p.m_tiles * p.n_tiles * p.splits;
  if (grid > num_sms()) grid = num_sms();
#ifdef PB_PDL
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(num_threads(GEGLU)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_gemm_kernel)");
#else
  kern<<<grid, num_threads(GEGLU), smem, st>>>(tmA, tmB, tmC, tmR, p);
#endif
  PB_LAUNCH_CHECK("tc_gemm_kernel");
  return PB_OK;
}

template <int MODE, bool GEGLU, typename TC>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                       const TcParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;
  static int max_clusters = 0;                      // co-resident CTA pairs
  auto kern = tc_gemm_pair_kernel<MODE, GEGLU, TC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_pair_kernel)");
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(num_threads(GEGLU)); cfg.dynamicSmemBytes = smem; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef PB_PDL
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 2;
#endif
  if (max_clusters == 0) {
    cfg.gridDim = dim3((unsigned)(num_sms() / 2 * 2));
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = num_sms() / 2; }
    max_clusters = n;
  }
  int clusters = ((p.m_tiles + 1) / 2) * p.n_tiles;
  if (clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, p);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(tc_gemm_pair_kernel)");
  PB_LAUNCH_CHECK("tc_gemm_pair_kernel");
  return PB_OK;
}

template <int MODE, bool GEGLU>
static int launch_c(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& r,
                    const TcParams& p, size_t smem, cudaStream_t st, int c_dtype) {
  if (p.pair) {
    if (c_dtype == PB_F32) return launch_pair<MODE, GEGLU, float>(a, b, c, r, p, smem, st);
    return launch_pair<MODE, GEGLU, bf16>(a, b, c, r, p, smem, st);
  }
  if (c_dtype == PB_F32) return launch<MODE, GEGLU, float>(a, b, c, r, p, smem, st);
  return launch<MODE, GEGLU, bf16>(a, b, c, r, p, smem, st);
}

}  // namespace tc
}  // namespace pb

namespace pb {
namespace tc {

// Everything pb_tc_gemm decides before it touches the device: tile geometry, split-K, CTA pairing, where C goes.
struct TcPlan {
  TcParams p;
  int a_rank;
  uint64_t adims[4], astr[3];
  uint32_t abox[4], aest[4];
  bool a_strided;
  uint64_t odims[3];      // output tensor as the C / residual maps see it: [cols][d1][d2][d3]
  uint32_t obox[3];
  void* c_base;
  int64_t ldc;
  int c_dtype;
  int n_cols;             // output columns
  bool geglu;
  bool colstats;          // the epilogue will write per-32-row column sums
};

#define TC_UNSUPPORTED(...)      \
  do {                           \
    set_error(__VA_ARGS__);      \
    return PB_EUNSUPPORTED;      \
  } while (0)

static int tc_plan(const pb_gemm_args* a, TcPlan* pl) {
  int rc = validate_gemm_args(a, "pb_tc_gemm");
  if (rc != PB_OK) return rc;
  if (a->a_dtype != PB_BF16) TC_UNSUPPORTED("pb_tc_gemm: operands must be bf16");
  const bool geglu = (a->epilogue & PB_EPI_GEGLU) != 0;
  const bool has_res = (a->epilogue & PB_EPI_RESIDUAL) != 0;
  if (a->N % 16 != 0 || (geglu && a->N % 32 != 0)) TC_UNSUPPORTED("pb_tc_gemm: N=%lld not a multiple of 16", (long long)a->N);
  if (a->lda % 8 != 0) TC_UNSUPPORTED("pb_tc_gemm: lda must be a multiple of 8 (16-byte rows)");
  if ((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->w)) & 15)
    TC_UNSUPPORTED("pb_tc_gemm: operands must be 16-byte aligned");
  if (geglu && (a->epilogue & (PB_EPI_ROWBIAS | PB_EPI_RESIDUAL)))
    TC_UNSUPPORTED("pb_tc_gemm: GEGLU cannot be combined with rowbias/residual");
  // C and the residual move as bulk tensor copies: 16-byte aligned bases and row pitches
  const uint64_t c_esz = a->c_dtype == PB_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(a->c) & 15) || (a->ldc * c_esz) % 16)
    TC_UNSUPPORTED("pb_tc_gemm: C must be 16-byte aligned with a 16-byte row pitch");
  if (has_res && ((reinterpret_cast<uintptr_t>(a->residual) & 15) || (a->ldr * 4) % 16))
    TC_UNSUPPORTED("pb_tc_gemm: residual must be 16-byte aligned with a 16-byte row pitch");
  if ((a->epilogue & PB_EPI_ROWBIAS) && ((reinterpret_cast<uintptr_t>(a->rowbias) & 15) || a->ld_rowbias % 4))
    TC_UNSUPPORTED("pb_tc_gemm: rowbias must be 16-byte aligned");

  memset(pl, 0, sizeof(*pl));
  TcParams& p = pl->p;
  p.bias = a->bias; p.rowbias = a->rowbias;
  p.M = a->M; p.ldrb = a->ld_rowbias;
  p.rb_rows = a->rowbias_rows > 0 ? a->rowbias_rows : 1;
  p.N = (int)a->N; p.epi = a->epilogue;
  p.geglu_half = geglu ? (int)(a->N / 2) : 0;
  pl->geglu = geglu;

  const int cin = a->a_mode == PB_A_LINEAR ? (int)a->K : a->cin;
  if (cin % BLOCK_K != 0) TC_UNSUPPORTED("pb_tc_gemm: K per tap (%d) must be a multiple of 64", cin);
  p.k_chunks = cin / BLOCK_K;
  const uint64_t esz = 2;
  pl->odims[0] = 0; pl->odims[1] = 1; pl->odims[2] = 1;
  pl->obox[0] = 1; pl->obox[1] = 1; pl->obox[2] = 1;
  bool slab_ok;   // every 32-row slab of a tile is 32 consecutive output rows starting at a multiple of 32
  if (a->a_mode == PB_A_LINEAR) {
    p.taps = 1;
    p.m_tiles = (int)ceil_div(a->M, BLOCK_M);
    pl->a_rank = 2;
    pl->adims[0] = (uint64_t)a->K; pl->adims[1] = (uint64_t)a->M;
    pl->astr[0] = (uint64_t)a->lda * esz;
    pl->abox[0] = BLOCK_K; pl->abox[1] = BLOCK_M;
    pl->odims[0] = (uint64_t)a->M; pl->obox[0] = BLOCK_M;
    p.n_outer = 1;
    slab_ok = true;
  } else if (a->a_mode == PB_A_CONV3X3) {
    if (a->upsample) TC_UNSUPPORTED("pb_tc_gemm: upsample is folded only on the CUDA-core path (use pb_upsample2x_rows)");
    p.taps = 9;
    // bricks tile the OUTPUT image; for stride 2 the TMA box walks the input with element stride 2, so the
    // strided gather is done by the copy engine (no im2col, no separate kernel)
    const int s2 = a->stride;
    const int W = a->out_w, H = a->out_h, n_img = (int)(a->M / ((int64_t)H * W));
    p.W = W; p.H = H; p.cstride = s2;
    bool exact = false;
    if (!pick_brick(W, H, n_img, &p.bw, &p.bh, &p.bn_img, &exact))
      TC_UNSUPPORTED("pb_tc_gemm: %dx%d x %d images cannot hold a 128-row brick", H, W, n_img);
    if (s2 * p.bw > 256 || s2 * p.bh > 256) TC_UNSUPPORTED("pb_tc_gemm: strided box exceeds the TMA box limit");
    p.tiles_x = (W + p.bw - 1) / p.bw; p.tiles_y = (H + p.bh - 1) / p.bh;
    p.m_tiles = p.tiles_x * p.tiles_y * ((n_img + p.bn_img - 1) / p.bn_img);
    pl->a_rank = 4;
    pl->adims[0] = (uint64_t)cin; pl->adims[1] = (uint64_t)a->in_w; pl->adims[2] = (uint64_t)a->in_h; pl->adims[3] = (uint64_t)n_img;
    pl->astr[0] = (uint64_t)a->lda * esz; pl->astr[1] = pl->astr[0] * a->in_w; pl->astr[2] = pl->astr[1] * a->in_h;
    pl->abox[0] = BLOCK_K; pl->abox[1] = (uint32_t)(s2 * p.bw); pl->abox[2] = (uint32_t)(s2 * p.bh); pl->abox[3] = (uint32_t)p.bn_img;
    pl->aest[0] = 1; pl->aest[1] = (uint32_t)s2; pl->aest[2] = (uint32_t)s2; pl->aest[3] = 1;
    pl->a_strided = s2 != 1;
    pl->odims[0] = (uint64_t)W; pl->odims[1] = (uint64_t)H; pl->odims[2] = (uint64_t)n_img;
    pl->obox[0] = (uint32_t)p.bw; pl->obox[1] = (uint32_t)p.bh; pl->obox[2] = (uint32_t)p.bn_img;
    p.n_outer = n_img;
    slab_ok = exact && ((int64_t)W * H) % 32 == 0 && ((p.bw % 32 == 0) || (p.bw == W && 32 % p.bw == 0 && p.bh % (32 / p.bw) == 0));
  } else {
    p.taps = 3;
    const int HW = a->frame_rows, T = a->frames, nb = (int)(a->M / ((int64_t)HW * T));
    p.W = HW; p.H = T; p.bn_img = 1;
    int one = 1;
    bool exact = false;
    if (!pick_brick(HW, T, 1, &p.bw, &p.bh, &one, &exact))
      TC_UNSUPPORTED("pb_tc_gemm: %d pixels x %d frames cannot hold a 128-row brick", HW, T);
    p.tiles_x = (HW + p.bw - 1) / p.bw; p.tiles_y = (T + p.bh - 1) / p.bh;
    p.m_tiles = p.tiles_x * p.tiles_y * nb;
    pl->a_rank = 4;
    pl->adims[0] = (uint64_t)cin; pl->adims[1] = (uint64_t)HW; pl->adims[2] = (uint64_t)T; pl->adims[3] = (uint64_t)nb;
    pl->astr[0] = (uint64_t)a->lda * esz; pl->astr[1] = pl->astr[0] * HW; pl->astr[2] = pl->astr[1] * T;
    pl->abox[0] = BLOCK_K; pl->abox[1] = (uint32_t)p.bw; pl->abox[2] = (uint32_t)p.bh; pl->abox[3] = 1;
    pl->odims[0] = (uint64_t)HW; pl->odims[1] = (uint64_t)T; pl->odims[2] = (uint64_t)nb;
    pl->obox[0] = (uint32_t)p.bw; pl->obox[1] = (uint32_t)p.bh; pl->obox[2] = 1;
    p.n_outer = nb;
    slab_ok = exact && ((int64_t)HW * T) % 32 == 0 && ((p.bw % 32 == 0) || (p.bw == HW && 32 % p.bw == 0 && p.bh % (32 / p.bw) == 0));
  }
  const int n_cols = geglu ? (int)(a->N / 2) : (int)a->N;   // output columns
  pl->n_cols = n_cols;
  const int bn_out = pick_block_n(n_cols, geglu ? 128 : 256, p.m_tiles);
  if (bn_out == 0) TC_UNSUPPORTED("pb_tc_gemm: no tile width divides N=%d", n_cols);
  p.block_n = geglu ? 2 * bn_out : bn_out;
  p.n_tiles = n_cols / bn_out;
  // split-K (deterministic: fp32 partials in the caller's workspace + a fold kernel) when the grid is tiny
  const int k_iters = p.taps * p.k_chunks;
  p.splits = 1; p.kps = k_iters;
  // split-K partials are consecutive copies of the output: an image-ragged brick would spill into the next copy
  const bool n_ragged = a->a_mode == PB_A_CONV3X3 && p.n_outer % p.bn_img != 0;
  const int want_splits = (geglu || n_ragged) ? 1 : pick_splits(p.m_tiles * p.n_tiles, k_iters);
  pl->c_base = a->c; pl->ldc = a->ldc; pl->c_dtype = a->c_dtype;
  if (want_splits > 1 && a->workspace && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0 &&
      a->workspace_bytes >= (int64_t)want_splits * a->M * a->N * (int64_t)sizeof(float)) {
    p.kps = (k_iters + want_splits - 1) / want_splits;
    p.splits = (k_iters + p.kps - 1) / p.kps;
    pl->c_base = a->workspace; pl->ldc = a->N; pl->c_dtype = PB_F32; p.epi = 0;
    p.bias = nullptr; p.rowbias = nullptr;
  }
  // CTA pairs whenever there are enough 256-row tiles to occupy the 74 SM pairs and no split-K
  p.pair = 0;
  if (p.splits == 1 && p.m_tiles >= 2) {
    const int pm = (p.m_tiles + 1) / 2;
    const int bn_pair = pick_block_n(n_cols, geglu ? 128 : 256, 2 * pm);   // waves counted over 74 pairs ~ 148 / 2
    if (bn_pair >= 32 && (long)pm * (n_cols / bn_pair) * 2 >= num_sms() / 2) {
      p.pair = 1;
      p.block_n = geglu ? 2 * bn_pair : bn_pair;
      p.n_tiles = n_cols / bn_pair;
    }
  }
  // fused GroupNorm column statistics: fp32 output written by this kernel itself, slabs aligned in row space
  pl->colstats = a->colstats != nullptr && p.splits == 1 && !geglu && slab_ok &&
                 !(a->c_dtype == PB_BF16 && (a->epilogue & PB_EPI_RESIDUAL)) &&   // (that staging slot is fp32-shaped)
                 a->ld_colstats >= 2 * (int64_t)n_cols && a->ld_colstats % 2 == 0 &&
                 (reinterpret_cast<uintptr_t>(a->colstats) & 7) == 0;
  if (pl->colstats) { p.colstats = a->colstats; p.ld_cs = a->ld_colstats; }
  const bool res_in_kernel = (p.epi & PB_EPI_RESIDUAL) != 0;
  p.slot_bytes = (pl->c_dtype == PB_F32 || res_in_kernel) ? EPI_SLOT_F32 : EPI_SLOT_BF16;
  return PB_OK;
}

}  // namespace tc
}  // namespace pb

extern "C" int pb_tc_gemm_colstats_ok(const pb_gemm_args* a) {
  pb::tc::TcPlan pl;
  if (pb::tc::tc_plan(a, &pl) != PB_OK) return 0;
  return pl.colstats ? 1 : 0;
}

extern "C" int pb_tc_gemm(const pb_gemm_args* a, pb_stream_t stream) {
  using namespace pb;
  using namespace pb::tc;
  TcPlan pl;
  int rc = tc_plan(a, &pl);
  if (rc != PB_OK) return rc;
  TcParams& p = pl.p;
  const bool geglu = pl.geglu;
  const uint64_t esz = 2;
  CUtensorMap tmA, tmB, tmC, tmR;
  rc = encode_tmap_bf16(&tmA, a->a, pl.a_rank, pl.adims, pl.astr, pl.abox, pl.a_strided ? pl.aest : nullptr);
  if (rc != PB_OK) return rc;
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->K * esz};
    uint32_t box[2] = {BLOCK_K, (uint32_t)(p.block_n / 2)};
    if (!p.pair && !geglu) box[1] = (uint32_t)p.block_n;
    rc = encode_tmap_bf16(&tmB, a->w, 2, dims, str, box, nullptr);
    if (rc != PB_OK) return rc;
  }
  const bool res_in_kernel = (p.epi & PB_EPI_RESIDUAL) != 0;
  {
    // C: [cols][d1][d2][d3 * splits] (split-K partials are consecutive copies of the output in the workspace)
    const uint64_t* od = pl.odims;
    const uint64_t ce = pl.c_dtype == PB_F32 ? 4 : 2;
    uint64_t dims[4] = {(uint64_t)pl.n_cols, od[0], od[1], od[2] * (uint64_t)p.splits};
    uint64_t str[3] = {(uint64_t)pl.ldc * ce, (uint64_t)pl.ldc * ce * od[0], (uint64_t)pl.ldc * ce * od[0] * od[1]};
    uint32_t box[4] = {EPI_CHUNK, pl.obox[0], pl.obox[1], pl.obox[2]};
    if (a->a_mode == PB_A_LINEAR) {   // [cols][M][splits][1]
      dims[2] = (uint64_t)p.splits; dims[3] = 1;
      str[1] = p.splits > 1 ? (uint64_t)a->M * a->N * 4 : str[0] * od[0];   // split pitch (fp32 workspace, ldc = N)
      str[2] = str[1] * dims[2];
    }
    rc = encode_tmap(&tmC, pl.c_dtype, pl.c_dtype == PB_F32 ? 128 : 64, pl.c_base, 4, dims, str, box, nullptr);
    if (rc != PB_OK) return rc;
    tmR = tmC;
    if (res_in_kernel) {
      uint64_t rdims[4] = {(uint64_t)pl.n_cols, od[0], od[1], od[2]};
      uint64_t rstr[3] = {(uint64_t)a->ldr * 4, (uint64_t)a->ldr * 4 * od[0], (uint64_t)a->ldr * 4 * od[0] * od[1]};
      rc = encode_tmap(&tmR, PB_F32, 128, a->residual, 4, rdims, rstr, box, nullptr);
      if (rc != PB_OK) return rc;
    }
  }
  // long-K tiles visit the epilogue rarely: one staging slot per team is enough and the ring gets the bytes
  p.epi_slots = (p.taps * p.k_chunks / p.splits >= 16) ? 1 : EPI_SLOTS;
  // short-K tiles with a residual are bound by the epilogue's HBM traffic: a third slot lets the residual loads run two
  // chunks ahead (the ring only needs as many stages as the tile has k-steps)
  // (measured: the Linear family of a 16x320x512 step 11.93 -> 11.41 ms)
  if (res_in_kernel && !geglu && p.taps * p.k_chunks <= 12) p.epi_slots = MAX_SLOTS;
  const int epi_bytes = num_teams(geglu) * p.epi_slots * p.slot_bytes;
  const int stage_bytes = A_STAGE_BYTES + (p.pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
  p.stages = (SMEM_TOTAL - 1024 - epi_bytes - BAR_BYTES) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) TC_UNSUPPORTED("pb_tc_gemm: tile does not fit shared memory");
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 /*alignment slack*/ + epi_bytes + BAR_BYTES;

  cudaStream_t st = to_stream(stream);
  const int c_dtype = pl.c_dtype;
  if (a->a_mode == PB_A_LINEAR)
    rc = geglu ? launch_c<PB_A_LINEAR, true>(tmA, tmB, tmC, tmR, p, smem, st, c_dtype)
               : launch_c<PB_A_LINEAR, false>(tmA, tmB, tmC, tmR, p, smem, st, c_dtype);
  else if (a->a_mode == PB_A_CONV3X3) rc = launch_c<PB_A_CONV3X3, false>(tmA, tmB, tmC, tmR, p, smem, st, c_dtype);
  else rc = launch_c<PB_A_CONVT3, false>(tmA, tmB, tmC, tmR, p, smem, st, c_dtype);
  if (rc != PB_OK || p.splits == 1) return rc;
  const int64_t total = a->M * (a->N / 4);
  int64_t blocks = ceil_div(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  const float* ws = reinterpret_cast<const float*>(a->workspace);
  const float* bias = (a->epilogue & PB_EPI_BIAS) ? a->bias : nullptr;
  const float* rowb = (a->epilogue & PB_EPI_ROWBIAS) ? a->rowbias : nullptr;
  const float* resid = (a->epilogue & PB_EPI_RESIDUAL) ? a->residual : nullptr;
  const int64_t rbr = a->rowbias_rows > 0 ? a->rowbias_rows : 1;
  if (a->c_dtype == PB_F32)
    splitk_reduce_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(ws, p.splits, a->M, (int)a->N, reinterpret_cast<float*>(a->c),
                                                                  a->ldc, bias, rowb, a->ld_rowbias, rbr, resid, a->ldr);
  else
    splitk_reduce_kernel<bf16><<<(unsigned)blocks, 256, 0, st>>>(ws, p.splits, a->M, (int)a->N, reinterpret_cast<bf16*>(a->c),
                                                                 a->ldc, bias, rowb, a->ld_rowbias, rbr, resid, a->ldr);
  PB_LAUNCH_CHECK("splitk_reduce_kernel");
  return PB_OK;
}
#undef TC_UNSUPPORTED

extern "C" size_t pb_tc_gemm_workspace_bytes(const pb_gemm_args* a) {
  // upper bound for the split-K partials of this problem (0 when split-K will not be used)
  using namespace pb::tc;
  if (!a || a->M <= 0 || a->N <= 0 || (a->epilogue & PB_EPI_GEGLU)) return 0;
  const int m_tiles = (int)((a->M + BLOCK_M - 1) / BLOCK_M);
  if (m_tiles * 2 > num_sms()) return 0;
  const int cin = a->a_mode == PB_A_LINEAR ? (int)a->K : a->cin;
  const int taps = a->a_mode == PB_A_LINEAR ? 1 : (a->a_mode == PB_A_CONV3X3 ? 9 : 3);
  const int k_iters = taps * (cin / BLOCK_K);
  int worst = 1;
  for (int bn = 256; bn >= 16; bn -= 16) {
    if (a->N % bn || (bn % 32 && bn != a->N)) continue;
    const int s = pick_splits(m_tiles * (int)(a->N / bn), k_iters);
    if (s > worst) worst = s;
  }
  return worst > 1 ? (size_t)worst * a->M * a->N * sizeof(float) : 0;
}
