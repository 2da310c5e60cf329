// Does the random-row request rate of B200 depend on the FOOTPRINT of the table (TLB reach / page walks) or only on
// the number of requests?  Random 64 B and 128 B rows out of tables of 0.25 .. 64 GiB, 2.56 M and 20.5 M lookups.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_footprint tools/gather_footprint.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_na(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ float4 ldg_na_128(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// MASK = which of the 8 16-byte chunks of a 128 B row are read; HINT = 1 adds the L2::128B prefetch-size hint
template <int MASK, int HINT, int U>
__global__ void gather_mask(const float4* __restrict__ table, const uint32_t* __restrict__ rows, int64_t n, float* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t group = tid / 8;
  const int sub = tid % 8;
  const int64_t groups = (int64_t)gridDim.x * blockDim.x / 8;
  float acc = 0.f;
  for (int64_t base = group; base < n; base += groups * U) {
    uint32_t r[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + u * groups;
      ok[u] = i < n;
      r[u] = ok[u] ? __ldg(rows + i) : 0;
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0, 0, 0, 0);
      if (ok[u] && ((MASK >> sub) & 1)) {
        const float4* p = table + (int64_t)r[u] * 8 + sub;
        v[u] = HINT ? ldg_na_128(p) : ldg_na(p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[tid] = acc;
}

// LPR lanes read ACTIVE*16 bytes of a row of PITCH16*16 bytes; U rows in flight per thread
template <int LPR, int ACTIVE, int PITCH16, int U>
__global__ void gather_rows(const float4* __restrict__ table, const uint32_t* __restrict__ rows, int64_t n, float* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t group = tid / LPR;
  const int sub = tid % LPR;
  const int64_t groups = (int64_t)gridDim.x * blockDim.x / LPR;
  float acc = 0.f;
  for (int64_t base = group; base < n; base += groups * U) {
    uint32_t r[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + u * groups;
      ok[u] = i < n;
      r[u] = ok[u] ? __ldg(rows + i) : 0;
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0, 0, 0, 0);
      if (ok[u] && sub < ACTIVE) v[u] = ldg_na(table + (int64_t)r[u] * PITCH16 + sub);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[tid] = acc;
}

__global__ void fill_rows(uint32_t* rows, int64_t n, uint32_t modulo, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    rows[i] = (uint32_t)(z % modulo);
  }
}

template <class F>
float time_it(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  const size_t max_bytes = (size_t)64 << 30;
  float4* table;
  CK(cudaMalloc(&table, max_bytes));
  CK(cudaMemset(table, 0, max_bytes));
  const int64_t n_small = 65536 * 39, n_big = 8 * n_small;
  uint32_t* rows;
  float* out;
  CK(cudaMalloc(&rows, 2 * n_big * sizeof(uint32_t)));
  CK(cudaMalloc(&out, 1 << 20));
  const double gib[] = {12.5, 25};
  printf("footprint_GiB  pattern  lookups  us  Mrows/s  DRAM_line_GB/s\n");
  for (double g : gib) {
    const size_t bytes = (size_t)(g * (1u << 30));
    int which = 0;
#define RUN(NAME, LPR, ACTIVE, PITCH16, U, N)                                                                  \
    {                                                                                                          \
      fill_rows<<<1024, 256>>>(rows, 2 * n_big, (uint32_t)(bytes / (PITCH16 * 16)), 77 + which);               \
      float ms = time_it([&] { gather_rows<LPR, ACTIVE, PITCH16, U><<<148 * 8, 256>>>(table, rows + (which++ & 1) * N, N, out); }, 10); \
      printf("%6.2f  %-22s %9lld  %8.1f  %8.1f  %8.1f\n", g, NAME, (long long)N, ms * 1e3, N / ms / 1e3, N * 128.0 / ms / 1e6);  \
    }
    RUN("64B row, 64B pitch", 4, 4, 4, 8, n_small)
    RUN("64B row, 64B pitch", 4, 4, 4, 8, n_big)
    RUN("80B of 128B pitch", 8, 5, 8, 8, n_small)
    RUN("80B of 128B pitch", 8, 5, 8, 8, n_big)
    RUN("128B row", 8, 8, 8, 8, n_big)
#define RUNM(NAME, MASK, HINT, N)                                                                              \
    {                                                                                                          \
      fill_rows<<<1024, 256>>>(rows, 2 * n_big, (uint32_t)(bytes / 128), 77 + which);                          \
      float ms = time_it([&] { gather_mask<MASK, HINT, 8><<<148 * 8, 256>>>(table, rows + (which++ & 1) * N, N, out); }, 10); \
      printf("%6.2f  %-22s %9lld  %8.1f  %8.1f  %8.1f\n", g, NAME, (long long)N, ms * 1e3, N / ms / 1e3, N * 128.0 / ms / 1e6);  \
    }
    if (g >= 12) {
      RUNM("chunks 0-4 (80B)", 0x1f, 0, n_big)
      RUNM("chunks 0-4 +L2::128B", 0x1f, 1, n_big)
      RUNM("chunks 0-4,7 (4 sect)", 0x9f, 0, n_big)
      RUNM("chunks 0-5 (96B)", 0x3f, 0, n_big)
      RUNM("chunks 0-3 (64B)", 0x0f, 0, n_big)
      RUNM("chunks 0-3 +L2::128B", 0x0f, 1, n_big)
      RUNM("chunks 0,2,4,6", 0x55, 0, n_big)
      RUNM("chunks 0-7 (128B)", 0xff, 0, n_big)
      RUNM("chunk 0 only (16B)", 0x01, 0, n_big)
      RUNM("chunks 0-4 (80B) small", 0x1f, 0, n_small)
      RUNM("chunks 0-4,7 small", 0x9f, 0, n_small)
      RUNM("chunks 0-7 small", 0xff, 0, n_small)
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
