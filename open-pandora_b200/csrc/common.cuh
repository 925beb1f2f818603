// common.cuh — shared helpers for the pandora_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pandora_b200.h"

namespace pb {

// Programmatic dependent launch (-DPB_PDL).  Every kernel lets its dependents start launching as soon as all of its own
// CTAs are resident (`pdl_trigger` first thing); the tcgen05 GEMM kernels are launched with the programmatic-serialisation
// attribute and call `pdl_wait` AFTER their prologue (barrier init, tensor-memory allocation, tensor-map prefetch) and
// before they touch global memory, so that prologue runs under the predecessor's tail.  `griddepcontrol.wait` returns
// when the prerequisite grids have completed and their memory is visible: results cannot depend on the overlap.
__device__ __forceinline__ void pdl_trigger() {
#ifdef PB_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifdef PB_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}


// thread-local last-error text (pb_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int validate_gemm_args(const pb_gemm_args* a, const char* who);
int validate_attn_args(const pb_attn_args* a, const char* who);

#define PB_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      pb::set_error(__VA_ARGS__);        \
      return PB_EINVAL;                  \
    }                                    \
  } while (0)

#define PB_LAUNCH_CHECK(what)                                   \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return pb::cuda_fail(e__, what);    \
  } while (0)

static inline cudaStream_t to_stream(pb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store with conversion to/from float (pointer must be 4-element aligned)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// exact (erf) GELU, matching torch.nn.functional.gelu default
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// dispatch a functor template on a PB_F32/PB_BF16 code
#define PB_DISPATCH_DTYPE(code, T, ...)                      \
  if ((code) == PB_F32) { using T = float; __VA_ARGS__ }     \
  else if ((code) == PB_BF16) { using T = pb::bf16; __VA_ARGS__ } \
  else { pb::set_error("bad dtype code %d", (int)(code)); return PB_EINVAL; }

}  // namespace pb
