// gemm_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM for the bf16 hot path (sm_100a).
//
//   C[m, n] = epilogue( sum_k A[m, k] * W[n, k] )      A, W bf16; fp32 accumulation in tensor memory
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0      TMA producer   : cp.async.bulk.tensor tiles of A (128 rows x 64 k) and W (block_n x 64 k)
//                                into a ring of 128B-swizzled shared-memory stages (mbarrier full/empty)
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma (M=128, N=block_n, K=16) x4 per
//                                stage; tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..5  epilogue       : tcgen05.ld the 128 x block_n fp32 accumulator in 32-column chunks (thread =
//                                row), + bias / GEGLU, transpose through a padded shared-memory stage, then
//                                row-contiguous (coalesced) row-bias / residual loads and stores
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// The A operand is addressed purely by TMA coordinates, so convolution needs no im2col buffer:
//   LINEAR   2-D map [M][K]
//   CONV3X3  4-D map [n][y][x][c]; tap (ky,kx) = the same box shifted by (ky-1, kx-1); out-of-bounds
//            rows/columns are zero-filled by the TMA unit, which IS the conv's zero padding
//   CONVT3   4-D map [b][t][p][c]; tap kt = box shifted by kt-1 frames (zero-filled outside the clip)
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
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 8;
constexpr int EPI_CHUNK = 32;                           // accumulator columns per epilogue chunk
constexpr int EPI_LD = EPI_CHUNK + 4;                   // padded row (floats): conflict-free v4 access
constexpr int EPI_STAGE_BYTES = BLOCK_M * EPI_LD * 4;   // 18 KiB, x2 buffers
constexpr int BAR_BYTES = 8 * (2 * MAX_STAGES + 4) + 16;
constexpr int SMEM_TOTAL = 227 * 1024;
constexpr int SMEM_PIPE_BUDGET = SMEM_TOTAL - 1024 - 2 * EPI_STAGE_BYTES - BAR_BYTES;

struct TcParams {
  void* c;
  const float* bias;
  const float* rowbias;
  const float* residual;
  int64_t M, ldc, ldr, ldrb, rb_rows;
  int N;                 // accumulator columns in total (weight rows)
  int block_n;           // accumulator columns per tile
  int n_tiles, m_tiles;
  int k_chunks;          // Cin / 64 (per tap) or K / 64
  int taps;              // 1, 9 or 3
  int stages;
  int epi;
  int geglu_half;        // N / 2 when GEGLU
  // brick geometry (conv modes)
  int bw, bh, bn_img;    // CONV3X3: brick x, y, images ; CONVT3: bw = pixels, bh = frames
  int W, H;              // CONV3X3: image size ; CONVT3: W = HW (pixels per frame), H = T
  int tiles_x, tiles_y;  // CONV3X3: W/bw, H/bh ; CONVT3: HW/bw, T/bh
  int cstride;           // CONV3X3 spatial stride (1 or 2); W/H above are OUTPUT sizes
  int splits;            // split-K ways (1 = off); split s covers k-iterations [s*kps, (s+1)*kps)
  int kps;               // k-iterations per split
  int64_t split_stride;  // elements between the partial outputs of consecutive splits
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

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE, bool GEGLU, typename TC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-byte alignment (swizzle-128B atoms)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  const uint32_t epi_base = smem_base + p.stages * stage_bytes;          // 2 x [128][EPI_LD] fp32
  const uint32_t bar_base = epi_base + 2 * EPI_STAGE_BYTES;
  // barriers: full[stages], empty[stages], tmem_full[2], tmem_empty[2]; then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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

  if (warp == 0) {
    // ================================================================== TMA producer
    if (lane == 0) {
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
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
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
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzled row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, ((it - it0) | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================== epilogue (warps 2..5)
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;       // accumulator row owned by this thread in phase A
    const int et = threadIdx.x - 64;         // 0..127
    const int b_row0 = et >> 3;              // phase B: rows b_row0 + 16*i, 4 columns at b_col
    const int b_col = (et & 7) * 4;
    int acc = 0, acc_phase = 0;
    uint32_t chunk_counter = 0;
    for (int work = blockIdx.x; work < total_tiles * p.splits; work += gridDim.x) {
      const int tile = work / p.splits, split = work - tile * p.splits;
      TC* C = reinterpret_cast<TC*>(p.c) + (int64_t)split * p.split_stride;   // split-K: partials, reduced afterwards
      const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
      // rows this thread stores in phase B (fixed for the tile)
      int64_t rows_m[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = out_row<MODE>(p, m_tile, b_row0 + 16 * i);
        rows_m[i] = m < p.M ? m : -1;
      }
      const int out_cols = GEGLU ? (p.block_n >> 1) : p.block_n;
      const int64_t n0 = (int64_t)n_tile * out_cols;       // first output column of the tile
      // residual + row-bias of a chunk are fetched one chunk ahead (registers), so their L2/HBM latency hides
      // behind the TMEM drain and the barrier instead of being exposed after it
      const bool has_add = (p.epi & (PB_EPI_RESIDUAL | PB_EPI_ROWBIAS)) != 0;
      float4 add_cur[8], add_nxt[8];
      auto fetch_add = [&](int c0, float4 (&dst)[8]) {
        const int cw = out_cols - c0 < EPI_CHUNK ? out_cols - c0 : EPI_CHUNK;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          const int64_t m = rows_m[i];
          if (m >= 0 && b_col < cw) {
            const int64_t n = n0 + c0 + b_col;
            if (p.epi & PB_EPI_RESIDUAL) t = *reinterpret_cast<const float4*>(p.residual + m * p.ldr + n);
            if (p.epi & PB_EPI_ROWBIAS) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(p.rowbias + (m / p.rb_rows) * p.ldrb + n));
              t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
          }
          dst[i] = t;
        }
      };
      if (has_add) fetch_add(0, add_cur);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
      for (int c0 = 0; c0 < out_cols; c0 += EPI_CHUNK) {
        const int cw = out_cols - c0 < EPI_CHUNK ? out_cols - c0 : EPI_CHUNK;   // 16 or 32
        const uint32_t buf = epi_base + (chunk_counter & 1u) * EPI_STAGE_BYTES;
        ++chunk_counter;
        if (has_add && c0 + EPI_CHUNK < out_cols) fetch_add(c0 + EPI_CHUNK, add_nxt);
        // ---- phase A: TMEM -> registers (+bias, GEGLU) -> padded smem stage, thread = row
#pragma unroll
        for (int h = 0; h < EPI_CHUNK; h += 16) {
          if (h < cw) {
            float v[16];
            tmem_ld16(t_row + (uint32_t)(c0 + h), v);
            float g[16];
            if (GEGLU) tmem_ld16(t_row + (uint32_t)(out_cols + c0 + h), g);
            tmem_ld_wait();
            if (p.epi & PB_EPI_BIAS) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + h + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                if (GEGLU) {
                  const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.bias + p.geglu_half + n0 + c0 + h + j));
                  g[j] += g4.x; g[j + 1] += g4.y; g[j + 2] += g4.z; g[j + 3] += g4.w;
                }
              }
            }
            if (GEGLU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= gelu_f(g[j]);
            }
            const uint32_t dst = buf + (uint32_t)(r * EPI_LD + h) * 4u;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 4u * j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          }
        }
        if (c0 + EPI_CHUNK >= out_cols) {
          // last chunk read: the accumulator buffer is drained, hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        epi_bar_sync();   // stage written by all 128 epilogue threads (also orders reuse of the other buffer)
        // ---- phase B: row-contiguous: 8 lanes x 16 B cover one 128-byte row segment
        if (b_col < cw) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = rows_m[i];
            if (m < 0) continue;
            float4 val;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w)
                         : "r"(buf + (uint32_t)((b_row0 + 16 * i) * EPI_LD + b_col) * 4u));
            if (has_add) { val.x += add_cur[i].x; val.y += add_cur[i].y; val.z += add_cur[i].z; val.w += add_cur[i].w; }
            const float o[4] = {val.x, val.y, val.z, val.w};
            store4(C + m * p.ldc + n0 + c0 + b_col, o);
          }
        }
        if (has_add) {
#pragma unroll
          for (int i = 0; i < 8; ++i) add_cur[i] = add_nxt[i];
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
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
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on the same barrier of BOTH CTAs
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
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
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int half_n = p.block_n >> 1;                        // W rows staged by this CTA
  const int stage_bytes = A_STAGE_BYTES + half_n * BLOCK_K * 2;
  const uint32_t epi_base = smem_base + p.stages * stage_bytes;
  const uint32_t bar_base = epi_base + 2 * EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };       // one per CTA (multicast commit)
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };   // one per CTA (multicast commit)
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };  // leader only, 8 arrivals
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int pair_m_tiles = (p.m_tiles + 1) >> 1;
  const int total_work = pair_m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.k_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
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

  if (warp == 0) {
    // ================================================================== TMA producer (both CTAs)
    if (lane == 0) {
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
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, (uint32_t)p.block_n);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int work = pair_id; work < total_work; work += num_pairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * stage_bytes;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_pair(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (it | k) != 0);
          umma_commit_pair(empty_bar(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================================================== epilogue (warps 2..5, both CTAs)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;
    const int b_row0 = et >> 3;
    const int b_col = (et & 7) * 4;
    TC* C = reinterpret_cast<TC*>(p.c);
    int acc = 0, acc_phase = 0;
    uint32_t chunk_counter = 0;
    for (int work = pair_id; work < total_work; work += num_pairs) {
      const int n_tile = work % p.n_tiles;
      const int m_tile = 2 * (work / p.n_tiles) + (int)rank;
      int64_t rows_m[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m_tile < p.m_tiles ? out_row<MODE>(p, m_tile, b_row0 + 16 * i) : p.M;
        rows_m[i] = m < p.M ? m : -1;
      }
      const int out_cols = GEGLU ? (p.block_n >> 1) : p.block_n;
      const int64_t n0 = (int64_t)n_tile * out_cols;
      const bool has_add = (p.epi & (PB_EPI_RESIDUAL | PB_EPI_ROWBIAS)) != 0;
      float4 add_cur[8], add_nxt[8];
      auto fetch_add = [&](int c0, float4 (&dst)[8]) {
        const int cw = out_cols - c0 < EPI_CHUNK ? out_cols - c0 : EPI_CHUNK;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          const int64_t m = rows_m[i];
          if (m >= 0 && b_col < cw) {
            const int64_t n = n0 + c0 + b_col;
            if (p.epi & PB_EPI_RESIDUAL) t = *reinterpret_cast<const float4*>(p.residual + m * p.ldr + n);
            if (p.epi & PB_EPI_ROWBIAS) {
              const float4 u = __ldg(reinterpret_cast<const float4*>(p.rowbias + (m / p.rb_rows) * p.ldrb + n));
              t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
          }
          dst[i] = t;
        }
      };
      if (has_add) fetch_add(0, add_cur);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
      for (int c0 = 0; c0 < out_cols; c0 += EPI_CHUNK) {
        const int cw = out_cols - c0 < EPI_CHUNK ? out_cols - c0 : EPI_CHUNK;
        const uint32_t buf = epi_base + (chunk_counter & 1u) * EPI_STAGE_BYTES;
        ++chunk_counter;
        if (has_add && c0 + EPI_CHUNK < out_cols) fetch_add(c0 + EPI_CHUNK, add_nxt);
#pragma unroll
        for (int h = 0; h < EPI_CHUNK; h += 16) {
          if (h < cw) {
            float v[16];
            tmem_ld16(t_row + (uint32_t)(c0 + h), v);
            float g[16];
            if (GEGLU) tmem_ld16(t_row + (uint32_t)(out_cols + c0 + h), g);
            tmem_ld_wait();
            if (p.epi & PB_EPI_BIAS) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + h + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                if (GEGLU) {
                  const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.bias + p.geglu_half + n0 + c0 + h + j));
                  g[j] += g4.x; g[j + 1] += g4.y; g[j + 2] += g4.z; g[j + 3] += g4.w;
                }
              }
            }
            if (GEGLU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] *= gelu_f(g[j]);
            }
            const uint32_t dst = buf + (uint32_t)(r * EPI_LD + h) * 4u;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 4u * j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          }
        }
        if (c0 + EPI_CHUNK >= out_cols) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(tempty_bar(acc));   // 4 warps x 2 CTAs free the accumulator pair
        }
        epi_bar_sync();
        if (b_col < cw) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = rows_m[i];
            if (m < 0) continue;
            float4 val;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(val.x), "=f"(val.y), "=f"(val.z), "=f"(val.w)
                         : "r"(buf + (uint32_t)((b_row0 + 16 * i) * EPI_LD + b_col) * 4u));
            if (has_add) { val.x += add_cur[i].x; val.y += add_cur[i].y; val.z += add_cur[i].z; val.w += add_cur[i].w; }
            const float o[4] = {val.x, val.y, val.z, val.w};
            store4(C + m * p.ldc + n0 + c0 + b_col, o);
          }
        }
        if (has_add) {
#pragma unroll
          for (int i = 0; i < 8; ++i) add_cur[i] = add_nxt[i];
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
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

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PB_ECUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return PB_ECUDA;
  }
  return PB_OK;
}

// Exact 128-row brick bw x bh x bn with bw | W, bh | H, bn | N; prefers wide bricks (long contiguous
// runs for the TMA unit).  Returns false when the tensor cannot be tiled exactly.
static bool pick_brick(int W, int H, int N, int* bw, int* bh, int* bn) {
  for (int w = W < BLOCK_M ? W : BLOCK_M; w >= 1; --w) {
    if (W % w != 0 || BLOCK_M % w != 0) continue;
    const int rest = BLOCK_M / w;
    for (int h = H < rest ? H : rest; h >= 1; --h) {
      if (H % h != 0 || rest % h != 0) continue;
      const int n = rest / h;
      if (n <= N && N % n == 0) { *bw = w; *bh = h; *bn = n; return true; }
    }
  }
  return false;
}

static int num_sms();

// Tile width: a multiple of 16 dividing n_cols, <= cap, minimising (waves over the SMs) x (per-tile cost);
// the +64 charges the A-tile traffic every tile pays regardless of its width, so narrow tiles only win when
// they remove a whole wave.
static int pick_block_n(int n_cols, int cap, int m_tiles) {
  int best = 0;
  long best_cost = 0;
  for (int bn = cap - cap % 16; bn >= 64 || (best == 0 && bn >= 16); bn -= 16) {
    if (n_cols % bn != 0) continue;
    const long tiles = (long)m_tiles * (n_cols / bn);
    const long waves = (tiles + num_sms() - 1) / num_sms();
    const long cost = waves * (bn + 64);
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
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;
  auto kern = tc_gemm_kernel<MODE, GEGLU, TC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_kernel)");
    configured = true;
  }
  int grid = p.m_tiles * p.n_tiles * p.splits;
  if (grid > num_sms()) grid = num_sms();
  kern<<<grid, NUM_THREADS, smem, st>>>(tmA, tmB, p);
  PB_LAUNCH_CHECK("tc_gemm_kernel");
  return PB_OK;
}

template <int MODE, bool GEGLU, typename TC>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, size_t smem, cudaStream_t st) {
  static bool configured = false;
  auto kern = tc_gemm_pair_kernel<MODE, GEGLU, TC>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(tc_gemm_pair_kernel)");
    configured = true;
  }
  int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
  if (pairs > num_sms() / 2) pairs = num_sms() / 2;
  kern<<<2 * pairs, NUM_THREADS, smem, st>>>(tmA, tmB, p);   // __cluster_dims__(2,1,1)
  PB_LAUNCH_CHECK("tc_gemm_pair_kernel");
  return PB_OK;
}

template <int MODE, bool GEGLU>
static int launch_c(const CUtensorMap& a, const CUtensorMap& b, const TcParams& p, size_t smem, cudaStream_t st,
                    int c_dtype) {
  if (p.pair) {
    if (c_dtype == PB_F32) return launch_pair<MODE, GEGLU, float>(a, b, p, smem, st);
    return launch_pair<MODE, GEGLU, bf16>(a, b, p, smem, st);
  }
  if (c_dtype == PB_F32) return launch<MODE, GEGLU, float>(a, b, p, smem, st);
  return launch<MODE, GEGLU, bf16>(a, b, p, smem, st);
}

}  // namespace tc
}  // namespace pb

extern "C" int pb_tc_gemm(const pb_gemm_args* a, pb_stream_t stream) {
  using namespace pb;
  using namespace pb::tc;
  int rc = validate_gemm_args(a, "pb_tc_gemm");
  if (rc != PB_OK) return rc;
#define TC_UNSUPPORTED(...)      \
  do {                           \
    set_error(__VA_ARGS__);      \
    return PB_EUNSUPPORTED;      \
  } while (0)
  if (a->a_dtype != PB_BF16) TC_UNSUPPORTED("pb_tc_gemm: operands must be bf16");
  const bool geglu = (a->epilogue & PB_EPI_GEGLU) != 0;
  if (a->N % 16 != 0 || (geglu && a->N % 32 != 0)) TC_UNSUPPORTED("pb_tc_gemm: N=%lld not a multiple of 16", (long long)a->N);
  if (a->lda % 8 != 0) TC_UNSUPPORTED("pb_tc_gemm: lda must be a multiple of 8 (16-byte rows)");
  if ((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->w)) & 15)
    TC_UNSUPPORTED("pb_tc_gemm: operands must be 16-byte aligned");
  if (geglu && (a->epilogue & (PB_EPI_ROWBIAS | PB_EPI_RESIDUAL)))
    TC_UNSUPPORTED("pb_tc_gemm: GEGLU cannot be combined with rowbias/residual");

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.c = a->c; p.bias = a->bias; p.rowbias = a->rowbias; p.residual = a->residual;
  p.M = a->M; p.ldc = a->ldc; p.ldr = a->ldr; p.ldrb = a->ld_rowbias;
  p.rb_rows = a->rowbias_rows > 0 ? a->rowbias_rows : 1;
  p.N = (int)a->N; p.epi = a->epilogue;
  p.geglu_half = geglu ? (int)(a->N / 2) : 0;

  CUtensorMap tmA, tmB;
  const int cin = a->a_mode == PB_A_LINEAR ? (int)a->K : a->cin;
  if (cin % BLOCK_K != 0) TC_UNSUPPORTED("pb_tc_gemm: K per tap (%d) must be a multiple of 64", cin);
  p.k_chunks = cin / BLOCK_K;
  const uint64_t esz = 2;
  if (a->a_mode == PB_A_LINEAR) {
    p.taps = 1;
    p.m_tiles = (int)ceil_div(a->M, BLOCK_M);
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->lda * esz};
    uint32_t box[2] = {BLOCK_K, BLOCK_M};
    rc = encode_tmap_bf16(&tmA, a->a, 2, dims, str, box, nullptr);
  } else if (a->a_mode == PB_A_CONV3X3) {
    if (a->upsample) TC_UNSUPPORTED("pb_tc_gemm: upsample is folded only on the CUDA-core path (use pb_upsample2x_rows)");
    p.taps = 9;
    // bricks tile the OUTPUT image; for stride 2 the TMA box walks the input with element stride 2, so the
    // strided gather is done by the copy engine (no im2col, no separate kernel)
    const int s2 = a->stride;
    const int W = a->out_w, H = a->out_h, n_img = (int)(a->M / ((int64_t)H * W));
    p.W = W; p.H = H; p.cstride = s2;
    if (!pick_brick(W, H, n_img, &p.bw, &p.bh, &p.bn_img))
      TC_UNSUPPORTED("pb_tc_gemm: %dx%d x %d images has no exact 128-row brick", H, W, n_img);
    if (s2 * p.bw > 256 || s2 * p.bh > 256) TC_UNSUPPORTED("pb_tc_gemm: strided box exceeds the TMA box limit");
    p.tiles_x = W / p.bw; p.tiles_y = H / p.bh;
    p.m_tiles = p.tiles_x * p.tiles_y * (n_img / p.bn_img);
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)a->in_w, (uint64_t)a->in_h, (uint64_t)n_img};
    uint64_t str[3] = {(uint64_t)a->lda * esz, (uint64_t)a->lda * esz * a->in_w,
                       (uint64_t)a->lda * esz * a->in_w * a->in_h};
    uint32_t box[4] = {BLOCK_K, (uint32_t)(s2 * p.bw), (uint32_t)(s2 * p.bh), (uint32_t)p.bn_img};
    uint32_t est[4] = {1, (uint32_t)s2, (uint32_t)s2, 1};
    rc = encode_tmap_bf16(&tmA, a->a, 4, dims, str, box, s2 == 1 ? nullptr : est);
  } else {
    p.taps = 3;
    const int HW = a->frame_rows, T = a->frames, nb = (int)(a->M / ((int64_t)HW * T));
    p.W = HW; p.H = T; p.bn_img = 1;
    int one = 1;
    if (!pick_brick(HW, T, 1, &p.bw, &p.bh, &one)) TC_UNSUPPORTED("pb_tc_gemm: %d pixels x %d frames has no exact 128-row brick", HW, T);
    p.tiles_x = HW / p.bw; p.tiles_y = T / p.bh;
    p.m_tiles = p.tiles_x * p.tiles_y * nb;
    uint64_t dims[4] = {(uint64_t)cin, (uint64_t)HW, (uint64_t)T, (uint64_t)nb};
    uint64_t str[3] = {(uint64_t)a->lda * esz, (uint64_t)a->lda * esz * HW, (uint64_t)a->lda * esz * HW * T};
    uint32_t box[4] = {BLOCK_K, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    rc = encode_tmap_bf16(&tmA, a->a, 4, dims, str, box, nullptr);
  }
  if (rc != PB_OK) return rc;
  const int n_cols = geglu ? (int)(a->N / 2) : (int)a->N;   // output columns
  const int bn_out = pick_block_n(n_cols, geglu ? 128 : 256, p.m_tiles);
  if (bn_out == 0) TC_UNSUPPORTED("pb_tc_gemm: no tile width divides N=%d", n_cols);
  p.block_n = geglu ? 2 * bn_out : bn_out;
  p.n_tiles = n_cols / bn_out;
  // split-K (deterministic: fp32 partials in the caller's workspace + a fold kernel) when the grid is tiny
  const int k_iters = p.taps * p.k_chunks;
  p.splits = 1; p.kps = k_iters; p.split_stride = 0;
  const int want_splits = geglu ? 1 : pick_splits(p.m_tiles * p.n_tiles, k_iters);
  if (want_splits > 1 && a->workspace &&
      a->workspace_bytes >= (int64_t)want_splits * a->M * a->N * (int64_t)sizeof(float)) {
    p.kps = (k_iters + want_splits - 1) / want_splits;
    p.splits = (k_iters + p.kps - 1) / p.kps;
    p.split_stride = a->M * a->N;
    p.c = a->workspace; p.ldc = a->N; p.epi = 0;
    p.bias = nullptr; p.rowbias = nullptr; p.residual = nullptr;
  }
  // CTA pairs whenever there are enough 256-row tiles to occupy the 74 SM pairs and no split-K
  static const bool pair_enabled = []() { const char* e = getenv("PB_TC_PAIR"); return !(e && e[0] == '0'); }();
  p.pair = 0;
  if (pair_enabled && p.splits == 1 && p.m_tiles >= 2) {
    const int pm = (p.m_tiles + 1) / 2;
    const int bn_pair = pick_block_n(n_cols, geglu ? 128 : 256, 2 * pm);   // waves counted over 74 pairs ~ 148 / 2
    if (bn_pair >= 32 && (long)pm * (n_cols / bn_pair) * 2 >= num_sms() / 2) {
      p.pair = 1;
      p.block_n = geglu ? 2 * bn_pair : bn_pair;
      p.n_tiles = n_cols / bn_pair;
    }
  }
  {
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t str[1] = {(uint64_t)a->K * esz};
    uint32_t box[2] = {BLOCK_K, (uint32_t)(p.block_n / 2)};
    if (!p.pair && !geglu) box[1] = (uint32_t)p.block_n;
    rc = encode_tmap_bf16(&tmB, a->w, 2, dims, str, box, nullptr);
    if (rc != PB_OK) return rc;
  }
  const int stage_bytes = A_STAGE_BYTES + (p.pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
  p.stages = SMEM_PIPE_BUDGET / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 /*alignment slack*/ + 2 * EPI_STAGE_BYTES + BAR_BYTES;

  cudaStream_t st = to_stream(stream);
  const int cd = p.splits > 1 ? PB_F32 : a->c_dtype;
  if (a->a_mode == PB_A_LINEAR)
    rc = geglu ? launch_c<PB_A_LINEAR, true>(tmA, tmB, p, smem, st, cd) : launch_c<PB_A_LINEAR, false>(tmA, tmB, p, smem, st, cd);
  else if (a->a_mode == PB_A_CONV3X3) rc = launch_c<PB_A_CONV3X3, false>(tmA, tmB, p, smem, st, cd);
  else rc = launch_c<PB_A_CONVT3, false>(tmA, tmB, p, smem, st, cd);
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
#undef TC_UNSUPPORTED
}

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
    if (a->N % bn) continue;
    const int s = pick_splits(m_tiles * (int)(a->N / bn), k_iters);
    if (s > worst) worst = s;
  }
  return worst > 1 ? (size_t)worst * a->M * a->N * sizeof(float) : 0;
}
