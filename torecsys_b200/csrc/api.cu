// Library-level entry points: version, error string, device probe.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace trs {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// Stream-ordered scratch (activation ping-pong buffers of the MLP chain, pre-split tensor-core weights, xDeepFM's
// embedding tile).  The device's default memory pool releases everything it holds at every synchronisation point unless
// told otherwise, which turns the allocation in each call into a fresh physical mapping (milliseconds) whenever the
// caller synchronises between calls -- e.g. the 'sync' index-check mode of the Python modules.  Keep up to 4 GB cached.
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s) {
  static thread_local int prepared_device = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  static const bool leave_pool_alone = getenv("TRS_POOL_RELEASE_DEFAULT") != nullptr;   // for A/B measurements
  if (dev != prepared_device && !leave_pool_alone) {
    cudaMemPool_t pool;
    e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e != cudaSuccess) return e;
    uint64_t keep = 0;
    e = cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (e != cudaSuccess) return e;
    const uint64_t want = 4ull << 30;
    if (keep < want) {
      keep = want;
      e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      if (e != cudaSuccess) return e;
    }
    prepared_device = dev;
  }
  return cudaMallocAsync(ptr, bytes, s);
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
