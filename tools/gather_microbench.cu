// Random-gather microbenchmark for B200: what does a random R-byte row read cost in DRAM traffic and time?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_microbench tools/gather_microbench.cu
//   ./gather_microbench [granularity 32|64|128]   (sets cudaLimitMaxL2FetchGranularity first)
// Prints useful GB/s per configuration; run under `ncu --metrics dram__bytes_read.sum` for the traffic.
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
__device__ __forceinline__ float4 ldg_plain(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldg_l2_64(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// each group of LPR lanes reads one row of LPR*16 bytes; U rows in flight per thread
template <int LPR, int U, int MODE>
__global__ void gather_rows(const float4* __restrict__ table, const int* __restrict__ rows, int64_t n_rows_to_read,
                            float* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t group = tid / LPR;
  const int sub = tid % LPR;
  const int64_t groups = (int64_t)gridDim.x * blockDim.x / LPR;
  float acc = 0.f;
  for (int64_t base = group; base < n_rows_to_read; base += groups * U) {
    int r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + u * groups;
      r[u] = i < n_rows_to_read ? __ldg(rows + i) : -1;
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0, 0, 0, 0);
      if (r[u] >= 0) {
        const float4* p = table + (int64_t)r[u] * LPR + sub;
        v[u] = MODE == 0 ? ldg_na(p) : (MODE == 1 ? ldg_plain(p) : ldg_l2_64(p));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 12345.678f) out[tid] = acc;
}

// packed-row pattern: rows of 128 B pitch; 4 lanes read 64 B, then lane 0 of the group reads 4 B at +64 (same line)
template <int U, int WITH_W>
__global__ void gather_packed(const float4* __restrict__ table, const int* __restrict__ rows, int64_t n, float* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t group = tid / 4;
  const int sub = tid % 4;
  const int64_t groups = (int64_t)gridDim.x * blockDim.x / 4;
  float acc = 0.f;
  for (int64_t base = group; base < n; base += groups * U) {
    int r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { int64_t i = base + u * groups; r[u] = i < n ? __ldg(rows + i) : -1; }
    float4 v[U]; float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_float4(0, 0, 0, 0); w[u] = 0.f;
      if (r[u] >= 0) {
        const float4* p = table + (int64_t)r[u] * 8;
        v[u] = ldg_na(p + sub);
        if (WITH_W && sub == (u & 3)) w[u] = __ldg(reinterpret_cast<const float*>(p + 4));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w + w[u];
  }
  if (acc == 12345.678f) out[tid] = acc;
}

// 4-byte random reads
template <int U>
__global__ void gather_scalars(const float* __restrict__ table, const int* __restrict__ rows, int64_t n, float* out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t threads = (int64_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int64_t base = tid; base < n; base += threads * U) {
    int r[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t i = base + u * threads;
      r[u] = i < n ? __ldg(rows + i) : -1;
    }
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = r[u] >= 0 ? __ldg(table + r[u]) : 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
  }
  if (acc == 12345.678f) out[tid] = acc;
}

__global__ void fill_rows(int* rows, int64_t n, uint32_t modulo, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    rows[i] = (int)(z % modulo);
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

int main(int argc, char** argv) {
  if (argc > 1) {
    size_t g = atoi(argv[1]);
    CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g));
  }
  size_t gran = 0;
  cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
  printf("cudaLimitMaxL2FetchGranularity = %zu\n", gran);
  const size_t table_bytes = (size_t)12800 << 20;  // 12.8 GiB-ish
  float4* table;
  CK(cudaMalloc(&table, table_bytes));
  CK(cudaMemset(table, 0, table_bytes));
  const int64_t n = 65536 * 39;  // lookups per "batch"
  const int nbatch = 8;
  int* rows;
  float* out;
  CK(cudaMalloc(&rows, n * nbatch * sizeof(int)));
  CK(cudaMalloc(&out, 1 << 20));
  int which = 0;
#define RUN_ROWS(LPR, U, MODE, CTAS)                                                                     \
  {                                                                                                      \
    fill_rows<<<1024, 256>>>(rows, n * nbatch, (uint32_t)(table_bytes / (LPR * 16)), 77);                \
    float ms = time_it([&] { gather_rows<LPR, U, MODE><<<148 * CTAS, 256>>>(table, rows + (which++ % nbatch) * n, n, out); }, 16); \
    printf("rows %4d B  U=%d mode=%d ctas/sm=%d : %8.1f us  useful %7.1f GB/s  %6.1f Mrows/s\n", LPR * 16, U, MODE, \
           CTAS, ms * 1e3, n * LPR * 16.0 / ms / 1e6, n / ms / 1e3);                                      \
  }
  RUN_ROWS(1, 8, 0, 8)
  RUN_ROWS(2, 8, 0, 8)
  RUN_ROWS(4, 4, 0, 8)
  RUN_ROWS(4, 8, 0, 8)
  RUN_ROWS(4, 8, 1, 8)
  RUN_ROWS(4, 8, 2, 8)
  RUN_ROWS(4, 16, 0, 8)
  RUN_ROWS(4, 8, 0, 4)
  RUN_ROWS(4, 8, 0, 2)
  RUN_ROWS(8, 8, 0, 8)
  RUN_ROWS(16, 8, 0, 8)
#define RUN_PACKED(U, W, CTAS, THREADS)                                                                   \
  {                                                                                                      \
    fill_rows<<<1024, 256>>>(rows, n * nbatch, (uint32_t)(table_bytes / 128), 55);                       \
    float ms = time_it([&] { gather_packed<U, W><<<148 * CTAS, THREADS>>>(table, rows + (which++ % nbatch) * n, n, out); }, 16); \
    printf("packed 128B pitch: 64B%s U=%d ctas/sm=%d thr=%d : %8.1f us %6.1f Mrows/s\n", W ? "+4B(w)" : "       ", U, CTAS, THREADS, ms * 1e3, n / ms / 1e3); \
  }
  RUN_PACKED(8, 0, 8, 256)
  RUN_PACKED(8, 1, 8, 256)
  RUN_PACKED(8, 0, 1, 256)
  RUN_PACKED(8, 1, 1, 256)
  RUN_PACKED(16, 1, 1, 256)
  RUN_PACKED(8, 1, 2, 256)
  RUN_PACKED(8, 1, 1, 512)
  RUN_PACKED(8, 1, 1, 1024)
  {
    fill_rows<<<1024, 256>>>(rows, n * nbatch, (uint32_t)((size_t)800 << 20 >> 2), 99);
    float ms = time_it([&] { gather_scalars<8><<<148 * 8, 256>>>((const float*)table, rows + (which++ % nbatch) * n, n, out); }, 16);
    printf("scalars 4 B from 800 MB, U=8 : %8.1f us  useful %7.1f GB/s %6.1f Mrows/s\n", ms * 1e3, n * 4.0 / ms / 1e6, n / ms / 1e3);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
