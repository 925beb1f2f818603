// abi.cu — library-level entry points: version, error reporting, device capability.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return PB_ECUDA;
}

}  // namespace pb

extern "C" {

int pb_version(void) { return 100; }

const char* pb_last_error(void) { return pb::g_err; }

int pb_device_supports_tc(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return pb::cuda_fail(e, "cudaGetDevice");
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return (major == 10) ? 1 : 0;
}

}  // extern "C"
