// sp_comm.cu — sequence-parallel communication as kernels over peer memory (NVLink / NVSwitch).
//
// The reference has no counterpart (no SP / CP of any kind); one clip is sharded by FRAMES for the
// per-frame layers (openaimodel3d.py:570, tensors are (b t) c h w) and by PIXELS for the per-pixel layers
// (attention.py:369-372, (b h w) t c), and the re-shard between the two is a Ulysses-style all-to-all.
//
// pb_sp_reshard: the all-to-all as ONE kernel.  Every rank PULLS: the producing GEMM of each rank wrote its rows
// into a send slot that all peers can address (symmetric memory / CUDA IPC over NVLink, or plain device memory for
// single-process virtual ranks); this kernel gathers this rank's share of every peer's slot straight into the row
// order the consumer wants (no pack pass, no unpack pass, no NCCL call).  Rows are [samples][frames][pixels] x cols:
//     direction 0, frames -> pixels : dst[(b, s*t_local + tl, pl)] = src_s[(b, tl, rank*hwl + pl)]
//     direction 1, pixels -> frames : dst[(b, tl, s*hwl + pl)]     = src_s[(b, rank*t_local + tl, pl)]
// (s = source rank, hwl = hw / P).  A segment (s, b, tl) is hwl consecutive rows on both sides.
//
// Cross-rank ordering is a flag barrier inside the kernel: block 0 publishes "my slot is ready" (epoch e) into every
// peer's flag array, every block waits until all peers have published e, then copies.  The epoch lives in device
// memory and is advanced by the last block to finish, so the kernel is CUDA-graph replayable.  Send slots alternate
// between two buffers: a rank can only be overwriting slot X for exchange k+2 after it has passed the barrier of
// exchange k+1, which every peer enters only after its pull of exchange k has completed (stream order).
// flags == NULL: the caller orders the ranks itself (virtual ranks on one device, host-side barriers).
//
// pb_sp_groupnorm_affine: the clip-wide GroupNorm statistics of the temporal blocks (openaimodel3d.py:259,
// attention.py:331,368) span all ranks: every rank leaves its fp64 (sum, sumsq) per (sample, group) in a peer-visible
// slot, and one small kernel per rank (same flag barrier) adds the P partials and builds scale/shift — replacing the
// per-norm NCCL all-reduce (105 per forward).
#include "common.cuh"

namespace pb {

struct SpBarrier {
  uint32_t* flags[PB_SP_MAX_RANKS];   // flags[d] = rank d's flag array [P] (peer-visible); NULL entries => no barrier
  uint32_t* epoch;                    // local {epoch, ticket}
  int P, rank;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Called by every thread of every block.  Returns the epoch of this exchange (0 when there is no barrier).
__device__ __forceinline__ uint32_t sp_barrier_enter(const SpBarrier& b) {
  if (b.epoch == nullptr) return 0;
  __shared__ uint32_t e_sh;
  if (threadIdx.x == 0) e_sh = ld_volatile_u32(b.epoch) + 1u;
  __syncthreads();
  const uint32_t e = e_sh;
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < b.P) {
    __threadfence_system();                       // everything this rank produced before the kernel is visible first
    st_release_sys(b.flags[threadIdx.x] + b.rank, e);
  }
  if ((int)threadIdx.x < b.P) {
    const uint32_t* mine = b.flags[b.rank] + threadIdx.x;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (++spins > (1u << 28)) {                  // a peer never arrived: trap instead of hanging the GPU
        printf("pandora_b200: sp barrier timeout (rank %d waits for rank %d, epoch %u)\n", b.rank, (int)threadIdx.x, e);
        __trap();
      }
    }
  }
  __syncthreads();
  return e;
}

// Called by every thread at the end: the last block to leave advances the epoch.
__device__ __forceinline__ void sp_barrier_leave(const SpBarrier& b, uint32_t e) {
  if (b.epoch == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t n_blocks = gridDim.x * gridDim.y;
    __threadfence();
    const uint32_t ticket = atomicAdd(b.epoch + 1, 1u);
    if (ticket == n_blocks - 1) {
      b.epoch[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(b.epoch) = e;
    }
  }
}

struct ReshardP {
  const uint8_t* src[PB_SP_MAX_RANKS];
  uint8_t* dst;
  int64_t src_row_bytes, dst_row_bytes;   // row pitches
  int vec_per_row;                        // 16-byte vectors per row (cols * elem_bytes / 16)
  int direction;
  int samples, t_local, hwl;              // frames per rank, pixels per rank
  SpBarrier bar;
};

constexpr int RS_THREADS = 256;
constexpr int RS_UNROLL = 4;

__global__ void __launch_bounds__(RS_THREADS)
sp_reshard_kernel(const ReshardP p) {
  pb::pdl_trigger();
  const uint32_t e = sp_barrier_enter(p.bar);
  const int P = p.bar.P, rank = p.bar.rank;
  const int T = P * p.t_local;
  const int64_t hw = (int64_t)P * p.hwl;
  // blockIdx.y = segment (s, b, tl); blockIdx.x strides over the segment's vectors
  const int seg = blockIdx.y;
  const int tl = seg % p.t_local;
  const int b = (seg / p.t_local) % p.samples;
  const int s = seg / (p.t_local * p.samples);
  int64_t src_row0, dst_row0;
  if (p.direction == 0) {
    src_row0 = ((int64_t)b * p.t_local + tl) * hw + (int64_t)rank * p.hwl;
    dst_row0 = ((int64_t)b * T + (int64_t)s * p.t_local + tl) * p.hwl;
  } else {
    src_row0 = ((int64_t)b * T + (int64_t)rank * p.t_local + tl) * p.hwl;
    dst_row0 = ((int64_t)b * p.t_local + tl) * hw + (int64_t)s * p.hwl;
  }
  const uint8_t* src = p.src[s] + src_row0 * p.src_row_bytes;
  uint8_t* dst = p.dst + dst_row0 * p.dst_row_bytes;
  const int64_t n_vec = (int64_t)p.hwl * p.vec_per_row;
  const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
  for (int64_t v0 = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; v0 < n_vec; v0 += stride * RS_UNROLL) {
    int4 val[RS_UNROLL];
    int64_t off_d[RS_UNROLL];
#pragma unroll
    for (int u = 0; u < RS_UNROLL; ++u) {           // all (remote) loads first: RS_UNROLL x 16 B in flight per thread
      const int64_t v = v0 + u * stride;
      if (v < n_vec) {
        const int64_t row = v / p.vec_per_row, col = v - row * p.vec_per_row;
        val[u] = *reinterpret_cast<const int4*>(src + row * p.src_row_bytes + col * 16);
        off_d[u] = row * p.dst_row_bytes + col * 16;
      }
    }
#pragma unroll
    for (int u = 0; u < RS_UNROLL; ++u)
      if (v0 + u * stride < n_vec) *reinterpret_cast<int4*>(dst + off_d[u]) = val[u];
  }
  sp_barrier_leave(p.bar, e);
}

struct GnPeersP {
  const double* sums[PB_SP_MAX_RANKS];   // per rank: [n_inst, G, 2] partial (sum, sumsq)
  SpBarrier bar;
  int n_inst, C, G;
  double count;                          // elements per (instance, group) over ALL ranks
  float eps;
  const float* gamma;
  const float* beta;
  float* scale_shift;                    // [n_inst, 2, C]
};

__global__ void __launch_bounds__(256)
sp_gn_affine_kernel(const GnPeersP p) {
  pb::pdl_trigger();
  __shared__ float s_mean[64], s_rstd[64];
  const uint32_t e = sp_barrier_enter(p.bar);
  const int ng = p.n_inst * p.G;
  for (int base = 0; base < ng; base += 64) {
    const int i = base + (int)threadIdx.x;
    if ((int)threadIdx.x < 64 && i < ng) {
      double s = 0.0, q = 0.0;
      for (int r = 0; r < p.bar.P; ++r) {           // fixed order: every rank computes bit-identical statistics
        s += p.sums[r][2 * i];
        q += p.sums[r][2 * i + 1];
      }
      const double mean = s / p.count;
      double var = q / p.count - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[threadIdx.x] = (float)mean;
      s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    __syncthreads();
    const int n_here = ng - base < 64 ? ng - base : 64;
    const int cpg = p.C / p.G;
    for (int j = threadIdx.x; j < n_here * cpg; j += blockDim.x) {
      const int gi = j / cpg, inst = (base + gi) / p.G, g = (base + gi) % p.G;
      const int c = g * cpg + (j - gi * cpg);
      const float a = s_rstd[gi] * p.gamma[c];
      p.scale_shift[(int64_t)inst * 2 * p.C + c] = a;
      p.scale_shift[(int64_t)inst * 2 * p.C + p.C + c] = p.beta[c] - s_mean[gi] * a;
    }
    __syncthreads();
  }
  sp_barrier_leave(p.bar, e);
}

static int fill_barrier(SpBarrier* b, const pb_sp_barrier* in, int P, int rank, const char* who) {
  b->P = P; b->rank = rank; b->epoch = nullptr;
  for (int i = 0; i < PB_SP_MAX_RANKS; ++i) b->flags[i] = nullptr;
  if (in == nullptr || in->epoch == nullptr) return PB_OK;
  for (int i = 0; i < P; ++i) {
    if (in->flags[i] == nullptr) { set_error("%s: barrier flags of rank %d missing", who, i); return PB_EINVAL; }
    b->flags[i] = in->flags[i];
  }
  b->epoch = in->epoch;
  return PB_OK;
}

}  // namespace pb

extern "C" {

int pb_sp_reshard(const pb_sp_reshard_args* a, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(a != nullptr, "pb_sp_reshard: null argument block");
  PB_REQUIRE(a->P >= 1 && a->P <= PB_SP_MAX_RANKS && a->rank >= 0 && a->rank < a->P, "pb_sp_reshard: bad rank / P");
  PB_REQUIRE(a->direction == 0 || a->direction == 1, "pb_sp_reshard: direction must be 0 or 1");
  PB_REQUIRE(a->samples > 0 && a->t_local > 0 && a->hw_local > 0 && a->cols > 0, "pb_sp_reshard: bad sizes");
  PB_REQUIRE(a->elem_bytes == 2 || a->elem_bytes == 4, "pb_sp_reshard: elem_bytes must be 2 or 4");
  PB_REQUIRE((a->cols * a->elem_bytes) % 16 == 0 && (a->ld_src * a->elem_bytes) % 16 == 0 &&
                 (a->ld_dst * a->elem_bytes) % 16 == 0 && a->ld_src >= a->cols && a->ld_dst >= a->cols,
             "pb_sp_reshard: rows must be multiples of 16 bytes");
  PB_REQUIRE(a->dst != nullptr && (reinterpret_cast<uintptr_t>(a->dst) & 15) == 0, "pb_sp_reshard: dst must be 16-byte aligned");
  ReshardP p;
  for (int i = 0; i < PB_SP_MAX_RANKS; ++i) p.src[i] = nullptr;
  for (int i = 0; i < a->P; ++i) {
    PB_REQUIRE(a->src[i] != nullptr && (reinterpret_cast<uintptr_t>(a->src[i]) & 15) == 0,
               "pb_sp_reshard: src[%d] missing or not 16-byte aligned", i);
    p.src[i] = reinterpret_cast<const uint8_t*>(a->src[i]);
  }
  p.dst = reinterpret_cast<uint8_t*>(a->dst);
  p.src_row_bytes = a->ld_src * a->elem_bytes;
  p.dst_row_bytes = a->ld_dst * a->elem_bytes;
  p.vec_per_row = (int)(a->cols * a->elem_bytes / 16);
  p.direction = a->direction;
  p.samples = (int)a->samples; p.t_local = (int)a->t_local; p.hwl = (int)a->hw_local;
  int rc = fill_barrier(&p.bar, &a->barrier, a->P, a->rank, "pb_sp_reshard");
  if (rc != PB_OK) return rc;
  const int64_t n_seg = (int64_t)a->P * a->samples * a->t_local;
  PB_REQUIRE(n_seg <= 65535, "pb_sp_reshard: too many segments");
  const int64_t n_vec = (int64_t)p.hwl * p.vec_per_row;
  // enough CTAs to keep ~8 KB in flight per SM, never more than the segment needs
  int64_t bx = ceil_div(n_vec, (int64_t)RS_THREADS * RS_UNROLL);
  const int64_t want = ceil_div(148 * 8, n_seg);
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_seg);
  sp_reshard_kernel<<<grid, RS_THREADS, 0, to_stream(stream)>>>(p);
  PB_LAUNCH_CHECK("sp_reshard_kernel");
  return PB_OK;
}

int pb_sp_groupnorm_affine(const double* const* sums, int32_t P, int32_t rank, const pb_sp_barrier* barrier,
                           int64_t n_inst, int64_t C, int32_t groups, double count, float eps, const float* gamma,
                           const float* beta, float* scale_shift, pb_stream_t stream) {
  using namespace pb;
  PB_REQUIRE(sums && gamma && beta && scale_shift, "pb_sp_groupnorm_affine: null pointer");
  PB_REQUIRE(P >= 1 && P <= PB_SP_MAX_RANKS && rank >= 0 && rank < P, "pb_sp_groupnorm_affine: bad rank / P");
  PB_REQUIRE(n_inst > 0 && C > 0 && groups > 0 && C % groups == 0 && count > 0, "pb_sp_groupnorm_affine: bad sizes");
  GnPeersP p;
  for (int i = 0; i < PB_SP_MAX_RANKS; ++i) p.sums[i] = nullptr;
  for (int i = 0; i < P; ++i) {
    PB_REQUIRE(sums[i] != nullptr, "pb_sp_groupnorm_affine: sums[%d] missing", i);
    p.sums[i] = sums[i];
  }
  int rc = fill_barrier(&p.bar, barrier, P, rank, "pb_sp_groupnorm_affine");
  if (rc != PB_OK) return rc;
  p.n_inst = (int)n_inst; p.C = (int)C; p.G = groups; p.count = count; p.eps = eps;
  p.gamma = gamma; p.beta = beta; p.scale_shift = scale_shift;
  sp_gn_affine_kernel<<<1, 256, 0, to_stream(stream)>>>(p);
  PB_LAUNCH_CHECK("sp_gn_affine_kernel");
  return PB_OK;
}

}  // extern "C"
