// Test infrastructure only: a minimal CPU emulation of the CUDA execution model (one OS thread per CUDA thread, CTAs run
// one after another, __syncthreads = std::barrier, CTA-uniform warp shuffles through a per-CTA exchange buffer) so that
// the *logic* of a kernel -- index arithmetic, ownership of accumulators, prefetch rings, reductions -- can be checked in
// the CPU test suite of a container without a GPU.  Nothing in the product imports or links this; the GPU parity tests
// (tests/test_gpu_*.py) remain the parity proof.  Warp-collective intrinsics (full-mask shuffles, mma.sync) synchronise
// the 32 threads of the calling warp through a per-warp barrier.
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

#include "../../include/torecsys_b200.h"

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
static std::vector<std::barrier<>*> warp_barrier;   // one per warp of the running CTA (warp-collective intrinsics)
static float shfl_buf[1024];
static uint32_t mma_buf[32][32][6];                 // [warp][lane][a0..a3, b0, b1]
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
      std::vector<std::barrier<>*> wb;
      for (unsigned w = 0; w < (threads + 31) / 32; ++w)
        wb.push_back(new std::barrier<>(threads - 32 * w < 32 ? threads - 32 * w : 32));
      warp_barrier = wb;
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([=] {
          threadIdx.x = t, blockIdx.x = bx, blockIdx.y = by;
          body();
        });
      for (auto& th : pool) th.join();
      for (auto* b : wb) delete b;
    }
}
}  // namespace emu

static inline void __syncthreads() { emu::block_barrier->arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {   // full-mask shuffles only
  const unsigned t = threadIdx.x;
  emu::shfl_buf[t] = v;
  emu::warp_barrier[t >> 5]->arrive_and_wait();
  const float r = emu::shfl_buf[(t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u)];
  emu::warp_barrier[t >> 5]->arrive_and_wait();
  return r;
}
static inline uint32_t __float_as_uint(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float v; std::memcpy(&v, &u, 4); return v; }

// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 with the PTX fragment layout (g = lane / 4, t = lane % 4):
//   A (16x8, row): a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]
//   B (8x8, col):  b0 = B[k = t][n = g], b1 = B[k = t+4][n = g]
//   C/D (16x8):    d0 = D[g][2t], d1 = D[g][2t+1], d2 = D[g+8][2t], d3 = D[g+8][2t+1]
// Operands are taken as the tf32 bit patterns the kernel prepared (low 13 mantissa bits already zero); products and
// sums in double, rounded once to float per call -- at least as accurate as the hardware's fp32 accumulation.
static inline void emu_mma_m16n8k8_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
  const unsigned t = threadIdx.x, w = t >> 5, lane = t & 31;
  uint32_t* mine = emu::mma_buf[w][lane];
  mine[0] = a0, mine[1] = a1, mine[2] = a2, mine[3] = a3, mine[4] = b0, mine[5] = b1;
  emu::warp_barrier[w]->arrive_and_wait();
  const int g = lane >> 2, tt = lane & 3;
  auto A = [&](int r, int k) {   // element (r, k) of the 16x8 A tile
    const uint32_t* src = emu::mma_buf[w][(r & 7) * 4 + (k & 3)];
    return (double)__uint_as_float(src[(r >> 3) + 2 * (k >> 2)]);
  };
  auto B = [&](int k, int n) { return (double)__uint_as_float(emu::mma_buf[w][n * 4 + (k & 3)][4 + (k >> 2)]); };
  double acc[4] = {d[0], d[1], d[2], d[3]};
  for (int k = 0; k < 8; ++k) {
    acc[0] += A(g, k) * B(k, 2 * tt);
    acc[1] += A(g, k) * B(k, 2 * tt + 1);
    acc[2] += A(g + 8, k) * B(k, 2 * tt);
    acc[3] += A(g + 8, k) * B(k, 2 * tt + 1);
  }
  emu::warp_barrier[w]->arrive_and_wait();
  for (int q = 0; q < 4; ++q) d[q] = (float)acc[q];
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
static inline float apply_act(float v, int act) {
  switch (act) {
    case TRS_ACT_RELU: return v > 0.f ? v : 0.f;
    case TRS_ACT_SIGMOID: return 1.0f / (1.0f + std::exp(-v));
    case TRS_ACT_TANH: return std::tanh(v);
    default: return v;
  }
}
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
