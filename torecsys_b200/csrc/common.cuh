// Shared helpers for the torecsys_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/torecsys_b200.h"

#ifndef __CUDA_ARCH__
#else
#if __CUDA_ARCH__ < 1000
#error "torecsys_b200 kernels are written for sm_100a (B200) only"
#endif
#endif

namespace trs {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- host-side error plumbing -----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
// cudaMallocAsync from the device's default pool, with the pool told to keep its memory across synchronisation points
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s);

#define TRS_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::trs::set_error(__VA_ARGS__);           \
      return TRS_ERR_INVALID_ARGUMENT;         \
    }                                          \
  } while (0)

#define TRS_UNSUPPORTED(cond, ...)             \
  do {                                         \
    if (cond) {                                \
      ::trs::set_error(__VA_ARGS__);           \
      return TRS_ERR_UNSUPPORTED;              \
    }                                          \
  } while (0)

#define TRS_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::trs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return TRS_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// opt a kernel in to the full 227 KB of dynamic shared memory, once per process AND device (the attribute is
// per device: a second GPU driven from the same process needs its own call)
constexpr int kMaxDynSmem = 227 * 1024;
#define TRS_SMEM_OPT_IN(kernel)                                                                          \
  do {                                                                                                   \
    static unsigned long long _done = 0;                                                                 \
    int _dev = 0;                                                                                        \
    TRS_CUDA(cudaGetDevice(&_dev));                                                                      \
    if (!((_done >> (_dev & 63)) & 1ull)) {                                                              \
      TRS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ::trs::kMaxDynSmem)); \
      _done |= 1ull << (_dev & 63);                                                                      \
    }                                                                                                    \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return TRS_ERR_CUDA;
  }
  return TRS_OK;
}

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// grid for a grid-stride kernel: enough CTAs for the work, capped at `waves` x resident CTAs on 148 SMs
inline int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t need = (work_items + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

// ---- device helpers ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// 128-bit read-only load that does not allocate in L1 (rows of a huge table are touched once)
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float2 ldg_stream_f2(const float2* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// streaming (evict-first) 128-bit store for outputs that are not re-read by this kernel
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <int IdxBits>
__device__ __forceinline__ int64_t load_index(const void* idx, int64_t pos) {
  if (IdxBits == 64) return __ldg(reinterpret_cast<const long long*>(idx) + pos);
  return static_cast<int64_t>(__ldg(reinterpret_cast<const int*>(idx) + pos));
}

// records an out-of-range lookup (see the `status` convention in torecsys_b200.h)
__device__ __forceinline__ void report_oob(int32_t* status, int64_t flat_pos) {
  if (status != nullptr) {
    atomicAdd(&status[0], 1);
    status[1] = static_cast<int32_t>(flat_pos & 0x7fffffff);
  }
}

// a zero the compiler cannot see through: keeps loop counters derived from it out of the uniform datapath
__device__ __forceinline__ int opaque_zero() {
  int z;
  asm volatile("mov.u32 %0, 0;" : "=r"(z));
  return z;
}

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

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case TRS_ACT_RELU: return fmaxf(v, 0.0f);
    case TRS_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case TRS_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// lexicographic pair index p -> (i, j), i < j < n.  Row i starts at p_i = i*(2n-i-1)/2.
__device__ __forceinline__ void pair_from_index(int p, int n, int& i, int& j) {
  float fn = static_cast<float>(2 * n - 1);
  int ii = static_cast<int>((fn - sqrtf(fn * fn - 8.0f * static_cast<float>(p))) * 0.5f);
  if (ii < 0) ii = 0;
  if (ii > n - 2) ii = n - 2;
  while (ii > 0 && ii * (2 * n - ii - 1) / 2 > p) --ii;
  while ((ii + 1) * (2 * n - ii - 2) / 2 <= p) ++ii;
  i = ii;
  j = p - ii * (2 * n - ii - 1) / 2 + ii + 1;
}

#endif  // __CUDACC__

}  // namespace trs
