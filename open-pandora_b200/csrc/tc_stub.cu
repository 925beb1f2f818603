// tc_stub.cu — placeholders until the tcgen05 kernels land (gemm_tc.cu / attention_tc.cu).
#include "common.cuh"

extern "C" int pb_tc_attention(const pb_attn_args*, pb_stream_t) {
  pb::set_error("pb_tc_attention: not built");
  return PB_EUNSUPPORTED;
}
