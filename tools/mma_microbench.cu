// How fast is the legacy mma.sync TF32 path on B200?  (decides MMA-vs-FFMA for the small fused MLP layers)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_microbench tools/mma_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MODE, int CHAINS>
__global__ void k(float* out, int iters, uint32_t seed) {
  float acc[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  float f0 = __uint_as_float((a0 & 0x007fffff) | 0x3f000000), f1 = 1.0001f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (MODE == 0) mma_tf32(acc[c], a0, a1, a2, a3, b0, b1);
      else if (MODE == 1) mma_bf16(acc[c], a0, a1, a2, a3, b0, b1);
      else { acc[c][0] = fmaf(acc[c][0], f1, f0); acc[c][1] = fmaf(acc[c][1], f1, f0); acc[c][2] = fmaf(acc[c][2], f1, f0); acc[c][3] = fmaf(acc[c][3], f1, f0); }
    }
  }
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += acc[c][i];
  if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE, int CHAINS>
void run(const char* name, int ctas_per_sm, int threads, double flops_per_op) {
  float* out; cudaMalloc(&out, 4096);
  const int iters = 4096;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, CHAINS><<<148 * ctas_per_sm, threads>>>(out, iters, 1);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k<MODE, CHAINS><<<148 * ctas_per_sm, threads>>>(out, iters, 2);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double warps = 148.0 * ctas_per_sm * threads / 32;
  double ops = warps * iters * CHAINS;   // warp-level ops
  printf("%-28s ctas/sm=%d thr=%d chains=%d : %8.3f ms  %8.2f Gwarp-ops/s  %8.1f TFLOP/s  (%.2f cycles/op/SM @1.9GHz)\n", name, ctas_per_sm,
         threads, CHAINS, ms, ops / ms / 1e6, ops * flops_per_op / ms / 1e9, 1.9e9 * ms * 1e-3 / (ops / 148));
  cudaFree(out);
}

int main() {
  run<0, 8>("mma.sync m16n8k8 tf32", 4, 256, 2.0 * 16 * 8 * 8);
  run<0, 2>("mma.sync m16n8k8 tf32", 4, 256, 2.0 * 16 * 8 * 8);
  run<0, 8>("mma.sync m16n8k8 tf32", 1, 256, 2.0 * 16 * 8 * 8);
  run<1, 8>("mma.sync m16n8k16 bf16", 4, 256, 2.0 * 16 * 8 * 16);
  run<2, 8>("ffma x4 (per warp-op 128 fma)", 4, 256, 2.0 * 128);
  return 0;
}
