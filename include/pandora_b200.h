/* pandora_b200.h — C ABI of the B200-native video-denoising hot path (libpandora_b200.so).
 *
 * The reference (Open-Pandora / DynamiCrafter) has no FFI: its hot path is a Python class surface
 * (`UNetModel.forward`, `DDIMSampler.sample`) whose arithmetic dispatches to torch/cuDNN/cuBLAS and
 * optionally xformers.  Every entry point below replaces the library kernel(s) one reference call
 * site dispatches to; the call site is cited as file:line relative to
 * /root/reference/DynamiCrafter/lvdm/.  The Python mirror of the reference interface
 * (`open-pandora_b200/pandora_b200/{unet,sampler}.py`) is the only caller and binds these with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary.
 *   - every pointer is DEVICE memory owned by the caller (PyTorch caching allocator), including
 *     workspaces.  The library allocates nothing on the device.
 *   - every call enqueues work on `stream` and returns immediately: no device sync, no host reads
 *     of device data => CUDA-graph capturable, safe under CUDA_LAUNCH_BLOCKING=1.
 *   - return 0 on success, negative PB_E* on error; pb_last_error() gives a thread-local message.
 *     Nothing throws, nothing calls exit().
 *   - activations are "rows": row-major [rows, channels] with an explicit row stride (`ld*`, in
 *     elements), channels contiguous.  A video tensor (b, c, t, h, w) is stored as rows ordered
 *     (b, t, y, x) — one layout end to end, so none of the reference's rearranges
 *     (modules/networks/openaimodel3d.py:43-45,236-238,570,606; modules/attention.py:300,307,369-372,
 *     391,406) exist here.
 *   - dtype codes: PB_F32 = 0, PB_BF16 = 1.
 */
#ifndef PANDORA_B200_H_
#define PANDORA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pb_stream_t; /* a cudaStream_t */

enum { PB_F32 = 0, PB_BF16 = 1 };
enum { PB_OK = 0, PB_EINVAL = -1, PB_ECUDA = -2, PB_EUNSUPPORTED = -3 };

/* A-operand addressing of pb_gemm (implicit GEMM): how logical A[m, k] is gathered from rows. */
enum {
  PB_A_LINEAR = 0, /* A[m,k] = a[m*lda + k]                                   nn.Linear / 1x1 conv  */
  PB_A_CONV3X3 = 1, /* m=(n,oy,ox), k=(ky,kx,c): zero-padded 3x3 window, stride 1|2, optional     */
                    /* nearest x2 upsample of the input folded into the addressing               */
  PB_A_CONVT3 = 2   /* m=(b,t,p),  k=(kt,c): zero-padded 3-tap window over frames                 */
};

/* epilogue flags of pb_gemm / pb_tc_* */
enum {
  PB_EPI_BIAS = 1,     /* + bias[n]                                                               */
  PB_EPI_ROWBIAS = 2,  /* + rowbias[(m / rowbias_rows) * ld_rowbias + n]   (timestep-embedding add) */
  PB_EPI_RESIDUAL = 4, /* + residual[m*ldr + n] (fp32)                                            */
  PB_EPI_GEGLU = 8     /* out[m,j] = u[m,j] * gelu(u[m, j + N/2]), u = acc + bias; N_out = N/2    */
};

typedef struct pb_gemm_args {
  const void* a;        /* activations, rows of a_dtype                                           */
  const void* w;        /* weights [N, K] row-major, K contiguous (torch Linear layout; conv kernels
                           repacked to k = tap*Cin + c), same dtype as a                          */
  void* c;              /* output rows [M, N_out] of c_dtype                                      */
  const float* bias;    /* [N] fp32 or NULL                                                       */
  const float* rowbias; /* fp32 or NULL                                                           */
  const float* residual;/* fp32 rows or NULL                                                      */
  int64_t M, N, K;
  int64_t lda, ldc, ldr, ld_rowbias;
  int64_t rowbias_rows;
  int32_t a_dtype, c_dtype;
  int32_t epilogue;     /* PB_EPI_* bit set                                                       */
  int32_t a_mode;       /* PB_A_*                                                                 */
  /* PB_A_CONV3X3: input is n_img images of in_h x in_w pixels x cin channels (row stride lda);
     output out_h x out_w; stride 1|2; upsample!=0 => conv runs on the nearest-x2 upsampled input */
  int32_t cin, in_h, in_w, out_h, out_w, stride, upsample;
  /* PB_A_CONVT3: rows ordered (b, t, p): frames = T, pixels per frame = HW                       */
  int32_t frames, frame_rows;
  /* optional scratch for pb_tc_gemm's split-K (small-M launches that would leave most SMs idle): size from
     pb_tc_gemm_workspace_bytes(); NULL or too small => no split (results identical up to fp32 summation order) */
  void* workspace;
  int64_t workspace_bytes;
  /* optional (pb_tc_gemm; of the ROUNDED values for a bf16 output): per-column (sum, sumsq) of the stored values over every 32-row slab of
     the output, colstats[(m / 32) * ld_colstats + 2 * n + {0, 1}] (ld_colstats in floats, >= 2 * N_out).  This is
     the first pass of the GroupNorm that consumes the output (openaimodel3d.py:155,179,259-268; attention.py:265,331)
     fused into the producer; pb_groupnorm_stats_cols folds it.  Written only when pb_tc_gemm_colstats_ok() says so
     (tile rows must map to aligned 32-row slabs, no split-K); NULL = off. */
  float* colstats;
  int64_t ld_colstats;
} pb_gemm_args;

/* ---- library ---------------------------------------------------------------------------------- */
int pb_version(void);
const char* pb_last_error(void);
/* 1 if the running device is sm_100 (tcgen05 paths usable), 0 otherwise, <0 on error */
int pb_device_supports_tc(void);

/* ---- GEMM / implicit-GEMM convolution ---------------------------------------------------------
 * Replaces nn.Linear (modules/attention.py:53-57,86-99,418-438,269,290,336,362), ResBlock /
 * Downsample / Upsample nn.Conv2d 3x3 (modules/networks/openaimodel3d.py:68,96,157,182,188),
 * the 1x1 skip conv (:190), the TemporalConvBlock nn.Conv3d (3,1,1) (:258-269) and the Conv1d k=1
 * projections of init_attn (modules/attention.py:332-334,360), with the bias add, the
 * `h + emb_out` add (:231), the skip / residual adds (:233,282; attention.py:243-245,310,412) and
 * GEGLU (attention.py:421-422) fused as epilogues.
 * pb_gemm: CUDA-core fp32-accumulate path, any shape, fp32 or bf16 operands (parity / fp32 mode).   */
int pb_gemm(const pb_gemm_args* args, pb_stream_t stream);

/* tcgen05 + TMEM + TMA path (sm_100a only): bf16 operands, fp32 accumulation in tensor memory.
 * Requires K % 64 == 0 (per tap for the conv modes), N % 16 == 0, 16-byte aligned rows.
 * Same argument block and semantics as pb_gemm.  Returns PB_EUNSUPPORTED for shapes it cannot take
 * (the caller then decides; there is no silent fallback inside the library).                      */
int pb_tc_gemm(const pb_gemm_args* args, pb_stream_t stream);
size_t pb_tc_gemm_workspace_bytes(const pb_gemm_args* args);
/* 1 if pb_tc_gemm(args) will fill args->colstats, 0 if it will not (shape, split-K, dtype) */
int pb_tc_gemm_colstats_ok(const pb_gemm_args* args);

/* ---- GroupNorm(32) (+SiLU) ---------------------------------------------------------------------
 * Replaces nn.GroupNorm at basics.py:76-88, openaimodel3d.py:155,179,259-268,547 (eps 1e-5) and
 * attention.py:265,331 (eps 1e-6) plus the following nn.SiLU.  x is rows [n_inst*rows_per_inst, C];
 * statistics are taken per (instance, group) over rows_per_inst x C/groups elements: an instance is
 * one frame for the 2-D blocks and all T frames of a sample for the temporal blocks.
 * pb_groupnorm_stats reduces the statistics and folds them with the affine parameters into
 * scale_shift [n_inst, 2, C] fp32 (y = x*scale + shift); pb_groupnorm_apply streams x once more,
 * applies it (+SiLU) and writes the operand dtype of the consuming GEMM.
 * Workspace for the partial sums: pb_groupnorm_ws_bytes.                                          */
size_t pb_groupnorm_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int64_t C);
int pb_groupnorm_stats(const void* x, int32_t x_dtype, int64_t ldx, int64_t n_inst,
                       int64_t rows_per_inst, int64_t C, int32_t groups, float eps, const float* gamma,
                       const float* beta, float* scale_shift /* [n_inst, 2, C] */, void* ws,
                       size_t ws_bytes, pb_stream_t stream);
/* Same result as pb_groupnorm_stats, but from the column statistics a producing pb_tc_gemm left behind
 * (pb_gemm_args.colstats) instead of a pass over x: colstats rows are 32-row slabs, rows_per_inst % 32 == 0. */
size_t pb_groupnorm_cols_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int32_t groups);
int pb_groupnorm_stats_cols(const float* colstats, int64_t ld_colstats, int64_t n_inst, int64_t rows_per_inst,
                            int64_t C, int32_t groups, float eps, const float* gamma, const float* beta,
                            float* scale_shift /* [n_inst, 2, C] */, void* ws, size_t ws_bytes, pb_stream_t stream);
/* The hot-path form: GroupNorm(+SiLU) of x from the producer's column statistics in two launches (slab-range partial
 * sums in fp64, then the apply pass, which folds the partials in its prologue: no finalize launch, no scale/shift
 * buffer).  ws: pb_groupnorm_apply_cols_ws_bytes. */
size_t pb_groupnorm_apply_cols_ws_bytes(int64_t n_inst, int64_t rows_per_inst, int32_t groups);
int pb_groupnorm_apply_cols(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                            const float* colstats, int64_t ld_colstats, int64_t n_inst, int64_t rows_per_inst,
                            int64_t C, int32_t groups, float eps, const float* gamma, const float* beta, int32_t silu,
                            void* ws, size_t ws_bytes, void* x_copy, int64_t ld_copy, pb_stream_t stream);
/* x_copy (optional, NULL = off): the apply pass also writes x itself, converted to y's dtype, with row pitch ld_copy —
 * the bf16 operand of a ResBlock's 1x1 skip convolution (openaimodel3d.py:190,233), which reads the same trunk tensor
 * the norm reads, without a second pass over it.                                                        */
int pb_groupnorm_apply(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype,
                       int64_t ldy, int64_t n_inst, int64_t rows_per_inst, int64_t C,
                       const float* scale_shift, int32_t silu, void* x_copy, int64_t ld_copy, pb_stream_t stream);

/* GroupNorm(+SiLU) of SMALL instances in ONE launch (a CTA per (instance, group) sums its slab and streams it again out
 * of L1 / L2 to apply): for the deep levels whose tensors take ~2 us to move, where three launches dominate.
 * PB_EUNSUPPORTED when C/groups is not a multiple of 4 or a slab exceeds 4 Mi elements.                  */
int pb_groupnorm_small(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                       int64_t n_inst, int64_t rows_per_inst, int64_t C, int32_t groups, float eps, const float* gamma,
                       const float* beta, int32_t silu, pb_stream_t stream);

/* Sequence-parallel split of pb_groupnorm_stats (temporal blocks normalise over ALL frames and pixels of a
 * clip, openaimodel3d.py:259, attention.py:331,368, while a rank holds a slice): pb_groupnorm_sums writes this
 * rank's fp64 (sum, sumsq) per (instance, group); the caller all-reduces them; pb_groupnorm_affine turns the
 * global sums (count = global elements per group) into scale_shift for pb_groupnorm_apply.          */
int pb_groupnorm_sums(const void* x, int32_t x_dtype, int64_t ldx, int64_t n_inst, int64_t rows_per_inst,
                      int64_t C, int32_t groups, double* sums /* [n_inst, groups, 2] */, void* ws,
                      size_t ws_bytes, pb_stream_t stream);
int pb_groupnorm_affine(const double* sums, int64_t n_inst, int64_t C, int32_t groups, double count,
                        float eps, const float* gamma, const float* beta, float* scale_shift,
                        pb_stream_t stream);

/* ---- LayerNorm ---------------------------------------------------------------------------------
 * Replaces nn.LayerNorm(dim) eps 1e-5 at attention.py:225-227,243-245.                            */
int pb_layernorm(const void* x, int32_t x_dtype, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy,
                 int64_t rows, int64_t C, const float* gamma, const float* beta, float eps,
                 pb_stream_t stream);

/* ---- row softmax -------------------------------------------------------------------------------
 * y[r, :] = softmax(x[r, :] * scale): the score matrix of the VAE decoder's single-head, d = C AttnBlock
 * (modules/networks/ae_modules.py:53-67: bmm, * C^-0.5, softmax over keys), between the two GEMMs that compute
 * q k^T and P v.  x fp32 rows; y fp32 or bf16 rows.                                                */
int pb_softmax_rows(const float* x, int64_t ldx, void* y, int32_t y_dtype, int64_t ldy, int64_t rows, int64_t n,
                    float scale, pb_stream_t stream);

/* ---- attention ---------------------------------------------------------------------------------
 * out = softmax(q k^T * scale) v per (batch, head), head_dim 64.  Replaces the einsum/softmax/einsum
 * core of CrossAttention.forward (attention.py:101-125,129-135) and xformers
 * memory_efficient_attention (:175,187).  Batch n of an operand lives at
 *   base + (n / bdiv) * bs_outer + (n % bdiv) * bs_inner,   row i at + i * rs,  head h at + h*64
 * (all in elements), which covers: spatial self-attention (batch = frame), text cross-attention
 * (K/V batch = frame / T, i.e. bdiv = T, bs_inner = 0), per-frame image cross-attention, and
 * temporal self-attention straight out of the (b, t, p) row order (batch = (b, p), rs = HW*ld).
 * accumulate != 0 adds into `out` (attention.py:142: out = out_text + 1.0 * out_image).           */
typedef struct pb_attn_operand {
  const void* ptr;
  int64_t bdiv, bs_outer, bs_inner, rs;
} pb_attn_operand;

typedef struct pb_attn_args {
  pb_attn_operand q, k, v;
  void* out;
  int64_t o_bdiv, o_bs_outer, o_bs_inner, o_rs;
  int64_t batch, heads, n_q, n_kv;
  float scale;
  int32_t dtype;      /* q/k/v/out element type */
  int32_t accumulate;
  /* > 0 (pb_tc_attention, n_kv <= 128): keys [0, kv_split) and [kv_split, n_kv) are two independent softmaxes whose
     outputs are summed: out = softmax(q k_text^T) v_text + softmax(q k_img^T) v_img with the text and image K/V
     rows concatenated (attention.py:128-142, image_cross_attention_scale = 1).  0 = one softmax. */
  int32_t kv_split;
} pb_attn_args;

int pb_attention(const pb_attn_args* args, pb_stream_t stream);    /* CUDA-core path, any n_q/n_kv */
int pb_tc_attention(const pb_attn_args* args, pb_stream_t stream); /* tcgen05 flash attention, bf16 */

/* ---- small / streaming kernels ---------------------------------------------------------------- */
/* GEGLU on materialised projections (attention.py:421-422): y[m,j] = u[m,j]*gelu(u[m,j+n_out])   */
int pb_geglu(const void* u, int32_t u_dtype, int64_t ldu, void* y, int32_t y_dtype, int64_t ldy,
             int64_t rows, int64_t n_out, pb_stream_t stream);
/* y = silu(x), flat (nn.SiLU of time_embed / emb_layers, openaimodel3d.py:172,376)               */
int pb_silu(const float* x, float* y, int64_t n, pb_stream_t stream);
/* y[i] = a[i] + b[i]  (emb + fs_embed, openaimodel3d.py:581)                                     */
int pb_add(const float* a, const float* b, float* y, int64_t n, pb_stream_t stream);
/* strided row copy with dtype conversion: channel concat of skip connections
 * (openaimodel3d.py:600) and fp32-trunk -> operand casts                                          */
int pb_copy_rows(const void* src, int32_t s_dtype, int64_t lds, void* dst, int32_t d_dtype,
                 int64_t ldd, int64_t rows, int64_t cols, pb_stream_t stream);
/* nearest x2 upsample of NHWC rows [n_img*H*W, cols] -> [n_img*2H*2W, cols] with dtype conversion
 * (Upsample, openaimodel3d.py:98-109; feeds the tcgen05 conv, pb_gemm folds it instead)           */
int pb_upsample2x_rows(const void* src, int32_t s_dtype, int64_t lds, void* dst, int32_t d_dtype,
                       int64_t ldd, int64_t n_img, int64_t H, int64_t W, int64_t cols,
                       pb_stream_t stream);
/* sinusoidal embedding with a caller-supplied frequency table (utils_diffusion.py:8-28; the table
 * is the bf16-evaluated one, uploaded as fp32): out[b, :half]=cos(t*f), out[b, half:]=sin(t*f)   */
int pb_timestep_embedding(const int64_t* t, const float* freqs, float* out, int64_t batch,
                          int64_t half, pb_stream_t stream);
/* (b, c, t, h, w) [+ second tensor concatenated on c] -> rows (b,t,y,x) x (c1+c2)
 * (DiffusionWrapper 'hybrid' cat, models/ddpm3d.py:1079, + rearrange openaimodel3d.py:570)       */
int pb_video_to_rows(const void* x1, int64_t c1, const void* x2, int64_t c2, int32_t in_dtype,
                     float* rows, int64_t ld, int64_t b, int64_t t, int64_t hw, pb_stream_t stream);
/* rows (b,t,y,x) x c -> (b, c, t, h, w)   (openaimodel3d.py:606)                                  */
int pb_rows_to_video(const float* rows, int64_t ld, void* out, int32_t out_dtype, int64_t b,
                     int64_t c, int64_t t, int64_t hw, pb_stream_t stream);

/* ---- CFG combine + DDIM update -----------------------------------------------------------------
 * One fused pass for DDIMSampler.p_sample_ddim (models/samplers/ddim.py:238-288) with
 * v-parameterisation (models/ddpm3d.py:235-247), dynamic rescale (ddim.py:273-277) and
 * rescale_noise_cfg (models/utils_diffusion.py:147-158, unbiased std) done as a 2-pass reduction.
 * Per-step scalars are read ON THE DEVICE from `tables` [n_steps, PB_DDIM_NCOEF] at row
 * *step_index (a device int32 the kernel post-increments when `advance`), so the 50-step loop has
 * no host<->device scalar traffic (the reference does 4-6 .item()-style reads per step,
 * ddim.py:263-266,274-275).  v_uncond may be NULL (no CFG).  Tensors are flat [batch, n] fp32.
 * x_prev MAY alias x (in-place update: every element is read once and then written by the same thread);
 * no other pair of arguments may overlap.                                                          */
enum { PB_DDIM_SQRT_AC = 0,      /* sqrt_alphas_cumprod[t]            (bf16 buffer value)           */
       PB_DDIM_SQRT_OMAC = 1,    /* sqrt_one_minus_alphas_cumprod[t]  (bf16 buffer value)           */
       PB_DDIM_SQRT_A_PREV = 2,  /* a_prev.sqrt()                                                   */
       PB_DDIM_SIGMA = 3,        /* sigma_t                                                         */
       PB_DDIM_X0_RESCALE = 4,   /* prev_scale_t / scale_t (1 when dynamic rescale is off)          */
       PB_DDIM_DIR_COEF = 5,     /* (1 - a_prev - sigma_t**2).sqrt()                                */
       PB_DDIM_NCOEF = 8 };
size_t pb_ddim_ws_bytes(int64_t batch);
int pb_ddim_cfg_step(const float* x, const float* v_cond, const float* v_uncond, const float* noise,
                     float* x_prev, float* pred_x0, int64_t batch, int64_t n, float cfg_scale,
                     float guidance_rescale, float temperature, const float* tables,
                     int32_t* step_index, int32_t advance, void* ws, size_t ws_bytes,
                     pb_stream_t stream);
/* The same step for DDIMSampler_multicond.p_sample_ddim (models/samplers/ddim_multiplecond.py:211-283): a third
 * U-Net output v_uncond_img (image conditioning kept, text dropped) and a second scale,
 *   v = v_uncond + cfg_img*(v_uncond_img - v_uncond) + cfg_scale*(v_cond - v_uncond_img)        (:234)
 * followed by the identical rescale / v->(eps,x0) / update.  v_uncond_img == NULL is pb_ddim_cfg_step.      */
int pb_ddim_cfg3_step(const float* x, const float* v_cond, const float* v_uncond, const float* v_uncond_img,
                      const float* noise, float* x_prev, float* pred_x0, int64_t batch, int64_t n,
                      float cfg_scale, float cfg_img, float guidance_rescale, float temperature,
                      const float* tables, int32_t* step_index, int32_t advance, void* ws, size_t ws_bytes,
                      pb_stream_t stream);


/* ---- sequence-parallel communication (one clip across P GPUs) ---------------------------------
 * No reference counterpart: the reference never shards a clip.  Frames are sharded for the
 * per-frame layers (modules/networks/openaimodel3d.py:570) and pixels for the per-pixel layers
 * (modules/attention.py:369-372); these two entry points are the data-path exchanges between the shardings,
 * written as kernels over peer-addressable memory (NVLink) instead of NCCL calls.
 *
 * pb_sp_barrier: flags[d] = rank d's flag array (P x uint32, peer-addressable), epoch = this rank's private
 * {epoch, ticket} pair (2 x uint32, zero-initialised).  epoch == NULL => no device-side barrier: the caller orders
 * the ranks itself (single-process virtual ranks with host barriers; tests).                        */
#define PB_SP_MAX_RANKS 16
typedef struct pb_sp_barrier {
  uint32_t* flags[PB_SP_MAX_RANKS];
  uint32_t* epoch;
} pb_sp_barrier;

/* Frame <-> pixel re-shard as ONE pull kernel.  Rows are [samples][frames][pixels] x cols.
 *   direction 0 (frames -> pixels): src[s] = rank s's frame-sharded rows (samples, t_local, hw_local*P);
 *                                   dst    = this rank's pixel-sharded rows (samples, t_local*P, hw_local)
 *   direction 1 (pixels -> frames): src[s] = rank s's pixel-sharded rows; dst = this rank's frame-sharded rows
 * src[rank] is this rank's own send buffer.  ld_* in elements; cols * elem_bytes must be a multiple of 16.   */
typedef struct pb_sp_reshard_args {
  const void* src[PB_SP_MAX_RANKS];
  void* dst;
  int64_t ld_src, ld_dst, cols;
  int64_t samples, t_local, hw_local;
  int32_t P, rank, direction, elem_bytes;
  pb_sp_barrier barrier;
} pb_sp_reshard_args;
int pb_sp_reshard(const pb_sp_reshard_args* args, pb_stream_t stream);

/* Clip-wide GroupNorm statistics across ranks: sums[r] = rank r's fp64 (sum, sumsq) per (instance, group)
 * (pb_groupnorm_sums output, in a peer-addressable slot); adds the P partials in rank order and writes
 * scale_shift [n_inst, 2, C] for pb_groupnorm_apply.  count = elements per (instance, group) over ALL ranks.
 * Replaces one NCCL all-reduce per temporal GroupNorm (openaimodel3d.py:259-268, attention.py:331).          */
int pb_sp_groupnorm_affine(const double* const* sums, int32_t P, int32_t rank, const pb_sp_barrier* barrier,
                           int64_t n_inst, int64_t C, int32_t groups, double count, float eps, const float* gamma,
                           const float* beta, float* scale_shift, pb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PANDORA_B200_H_ */
