// Library-level entry points: version, error string, device probe.
#include <stdarg.h>

#include "common.cuh"

namespace trs {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
}  // namespace trs

extern "C" {

const char* trs_version(void) { return "torecsys_b200 0.1.0 (sm_100a)"; }

const char* trs_last_error(void) { return trs::g_error; }

int trs_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  TRS_CUDA(cudaGetDevice(&dev));
  TRS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TRS_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

}  // extern "C"
