// Test infrastructure only: a minimal CPU emulation of the CUDA execution model (one OS thread per CUDA thread, CTAs run
// one after another, __syncthreads = std::barrier, CTA-uniform warp shuffles through a per-CTA exchange buffer) so that
// the *logic* of a kernel -- index arithmetic, ownership of accumulators, prefetch rings, reductions -- can be checked in
// the CPU test suite of a container without a GPU.  Nothing in the product imports or links this; the GPU parity tests
// (tests/test_gpu_*.py) remain the parity proof.  Shuffles here require every thread of the CTA to execute them
// (true for the kernels emulated so far).
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct emu_uint3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emu_uint3 threadIdx, blockIdx;
static emu_uint3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
static std::barrier<>* block_barrier = nullptr;
static float shfl_buf[1024];
static std::mutex atomic_mu;
alignas(16) static float dyn_smem[58 * 1024];   // 227 KB
template <class F>
void launch(unsigned gx, unsigned gy, unsigned threads, F body) {
  gridDim.x = gx, gridDim.y = gy, gridDim.z = 1;
  blockDim.x = threads, blockDim.y = blockDim.z = 1;
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx) {
      std::barrier<> bar(threads);
      block_barrier = &bar;
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([=] {
          threadIdx.x = t, blockIdx.x = bx, blockIdx.y = by;
          body();
        });
      for (auto& th : pool) th.join();
    }
}
}  // namespace emu

static inline void __syncthreads() { emu::block_barrier->arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned t = threadIdx.x;
  emu::shfl_buf[t] = v;
  emu::block_barrier->arrive_and_wait();
  const float r = emu::shfl_buf[(t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u)];
  emu::block_barrier->arrive_and_wait();
  return r;
}
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline float atomicAdd(float* p, float v) {
  std::lock_guard<std::mutex> lock(emu::atomic_mu);
  const float old = *p;
  *p = old + v;
  return old;
}

namespace trs {
static inline float4 ldg_stream_f4(const float4* p) { return *p; }
static inline float2 ldg_stream_f2(const float2* p) { return *p; }
static inline float ldg_stream_f1(const float* p) { return *p; }
static inline int opaque_zero() { return 0; }
static inline void stg_stream_f4(float4* p, const float4& v) { *p = v; }
// lexicographic pair index -> (i, j): the plain definition (csrc/common.cuh has the closed form)
static inline void pair_from_index(int p, int n, int& i, int& j) {
  int q = 0;
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b, ++q)
      if (q == p) { i = a, j = b; return; }
  i = j = -1;
}
}  // namespace trs
