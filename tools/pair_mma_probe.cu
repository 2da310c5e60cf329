// Probe: tcgen05.mma.cta_group::2 (a CTA pair, M = 256 across two SMs) with kind::tf32, SWIZZLE_NONE K-major operands.
// Verifies, numerically, the operand placement this repo would rely on before the CIN kernel is moved to CTA pairs:
//   * A: each CTA supplies ITS 128 rows at the same shared-memory offset;   * B: N columns split in halves, CTA r holding
//     columns [r * N/2, (r + 1) * N/2) at the same offset;   * D: each CTA's tensor memory gets its 128 rows x N columns;
//   * commit with multicast arrives on the same barrier offset in both CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pair_mma_probe tools/pair_mma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../torecsys_b200/csrc/tc5.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

using namespace trs::tc5;

constexpr int kN = 64;   // total N; each CTA holds kN / 2 columns of B
constexpr int kK = 16;   // two k-steps

__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_mma(const float* __restrict__ a_in /* [256][kK] */, const float* __restrict__ b_in /* [kN][kK] */,
         float* __restrict__ d_out /* [256][kN] */, int* __restrict__ flag) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* a_s = reinterpret_cast<float*>(sm);                       // [kK/4][128 rows][4]
  float* b_s = reinterpret_cast<float*>(sm + 128 * kK * 4);        // [kK/4][kN/2 rows][4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 128 * kK * 4 + (kN / 2) * kK * 4);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = cta_rank();
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * kK; i += blockDim.x) {
    const int r = i / kK, k = i - r * kK;
    a_s[((k >> 2) * 128 + r) * 4 + (k & 3)] = a_in[(rank * 128 + r) * kK + k];
  }
  for (int i = threadIdx.x; i < (kN / 2) * kK; i += blockDim.x) {
    const int n = i / kK, k = i - n * kK;
    b_s[((k >> 2) * (kN / 2) + n) * 4 + (k & 3)] = b_in[(rank * (kN / 2) + n) * kK + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kN >> 3) << 17) |
                             (static_cast<uint32_t>(256 >> 4) << 24);
      const uint64_t ad = umma_desc(smem_u32(a_s), 128 * 16, 128);
      const uint64_t bd = umma_desc(smem_u32(b_s), (kN / 2) * 16, 128);
      for (int ks = 0; ks < kK / 8; ++ks) {
        const uint64_t a_k = ad + ((2 * 128 * 16 * ks) >> 4), b_k = bd + ((2 * (kN / 2) * 16 * ks) >> 4);
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem), "l"(a_k), "l"(b_k), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  {
    const int r = threadIdx.x;   // TMEM lane = row of this CTA
    const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * warp) << 16);
    for (int c = 0; c < kN; c += 16) {
      uint32_t raw[16];
      tmem_ld16(taddr + c, raw);
      for (int j = 0; j < 16; ++j) d_out[(rank * 128 + r) * kN + c + j] = __uint_as_float(raw[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
  if (threadIdx.x == 0 && rank == 0) *flag = 1;
}

int main() {
  float *a, *b, *d;
  int* flag;
  CK(cudaMallocManaged(&a, 256 * kK * 4));
  CK(cudaMallocManaged(&b, kN * kK * 4));
  CK(cudaMallocManaged(&d, 256 * kN * 4));
  CK(cudaMallocManaged(&flag, 4));
  for (int i = 0; i < 256 * kK; ++i) a[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < kN * kK; ++i) b[i] = (float)((i * 5 + 1) % 13 - 6);
  for (int i = 0; i < 256 * kN; ++i) d[i] = -12345.f;
  *flag = 0;
  const size_t smem = 128 * kK * 4 + (kN / 2) * kK * 4 + 64;
  pair_mma<<<2, 128, smem>>>(a, b, d, flag);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch + sync: %s, flag %d\n", cudaGetErrorString(e), *flag);
  if (e != cudaSuccess) return 1;
  int bad = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < kN; ++n) {
      float want = 0.f;
      for (int k = 0; k < kK; ++k) want += a[r * kK + k] * b[n * kK + k];
      if (fabsf(want - d[r * kN + n]) > 1e-3f) {
        if (bad < 8) printf("mismatch row %d col %d: got %g want %g\n", r, n, d[r * kN + n], want);
        ++bad;
      }
    }
  printf("pair MMA (M = 256 over two CTAs, N = %d split %d + %d, K = %d): %d mismatches of %d\n", kN, kN / 2, kN / 2, kK, bad,
         256 * kN);
  return bad != 0;
}
