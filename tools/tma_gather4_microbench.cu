// TMA tile::gather4 microbenchmark (sm_100a): random 128-byte-pitch rows fetched four at a time by
// cp.async.bulk.tensor.2d ... tile::gather4 into a per-warp shared-memory ring, one issuing lane per warp, completion
// on mbarriers -- against the LDGSTS (cp.async, 8 lanes per row) scheme of deepfm_packed.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_gather4 tools/tma_gather4_microbench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kWarps = 8;
constexpr int kRowsPerStage = 80;        // rows of one warp-tile (16 samples x 5 fields)
constexpr int kOps = kRowsPerStage / 4;  // gather4 ops per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: returns false after ~2^22 polls (a wrong descriptor must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
template <int BOX_FLOATS>
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// rows: (n) int32 row ids; every warp takes stages of 80 rows round-robin
template <int BOX_FLOATS, int kStages>
__global__ void __launch_bounds__(kWarps * 32, 1) tma_gather_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ rows,
                                                                    int64_t n_stages, float* __restrict__ out, int* __restrict__ err) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kOpBytes = 4 * BOX_FLOATS * 4;                         // bytes one gather4 delivers
  constexpr int kOpPitch = (kOpBytes + 127) / 128 * 128;              // destinations are 128-byte aligned
  constexpr int kStageBytes = kOps * kOpPitch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* my = smem + (size_t)warp * kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kWarps * kStages * kStageBytes) + warp * kStages;
  if (lane == 0)
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(bars + s), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int64_t w_global = (int64_t)blockIdx.x * kWarps + warp, w_total = (int64_t)gridDim.x * kWarps;
  auto issue = [&](int64_t st, int slot) {
    if (st >= n_stages) return;
    if (lane == 0) {
      const uint32_t bar = smem_u32(bars + slot);
      mbar_expect_tx(bar, kOps * kOpBytes);
      const int4* ids = reinterpret_cast<const int4*>(rows + st * kRowsPerStage);
#pragma unroll 4
      for (int o = 0; o < kOps; ++o) {
        const int4 r = __ldg(ids + o);
        gather4<BOX_FLOATS>(smem_u32(my + (size_t)slot * kStageBytes + o * kOpPitch), &map, 0, r.x, r.y, r.z, r.w, bar);
      }
    }
  };
  int64_t st = w_global;
  for (int s = 0; s < kStages - 1; ++s) issue(st + s * w_total, s);
  float acc = 0.f;
  uint32_t phase[kStages];
  for (int s = 0; s < kStages; ++s) phase[s] = 0;
  int slot = 0, fill = kStages - 1;
  for (; st < n_stages; st += w_total) {
    issue(st + (kStages - 1) * w_total, fill);
    if (!mbar_wait(smem_u32(bars + slot), phase[slot])) {
      if (lane == 0) atomicAdd(err, 1);
      return;
    }
    phase[slot] ^= 1;
    // consume: every lane reads a few words of the stage (first float of rows lane, lane+32, lane+64)
    const unsigned char* base = my + (size_t)slot * kStageBytes;
    for (int r = lane; r < kRowsPerStage; r += 32)
      acc += *reinterpret_cast<const float*>(base + (r >> 2) * kOpPitch + (r & 3) * BOX_FLOATS * 4);
    __syncwarp();
    slot = slot + 1 == kStages ? 0 : slot + 1;
    fill = fill + 1 == kStages ? 0 : fill + 1;
  }
  if (out) out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void fill_table(float* t, int64_t rows) {   // row r: every float = r (exact up to 2^24) ... low bits via mod
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * 32; i += (int64_t)gridDim.x * blockDim.x)
    t[i] = static_cast<float>((i >> 5) & 0xffffff);
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
__global__ void expected_sum(const int* rows, int64_t n, double* out) {   // sum over all rows of (row & 0xffffff)
  double s = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += rows[i] & 0xffffff;
  atomicAdd(out, s);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BOX_FLOATS, int kStages>
void run(EncodeTiled encode, float* table, int64_t table_rows, const int* rows, int64_t n_rows, float* out, int* err, double want) {
  CUtensorMap map;
  cuuint64_t dims[2] = {32, (cuuint64_t)table_rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {BOX_FLOATS, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("box %d floats: cuTensorMapEncodeTiled failed (%d)\n", BOX_FLOATS, (int)r); return; }
  constexpr int kOpPitch = (4 * BOX_FLOATS * 4 + 127) / 128 * 128;
  const size_t smem = (size_t)kWarps * kStages * kOps * kOpPitch + kWarps * kStages * 8 + 128;
  CK(cudaFuncSetAttribute(tma_gather_kernel<BOX_FLOATS, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_stages = n_rows / kRowsPerStage;
  CK(cudaMemset(err, 0, 4));
  auto launch = [&] { tma_gather_kernel<BOX_FLOATS, kStages><<<148, kWarps * 32, smem>>>(map, rows, n_stages, out, err); };
  launch();
  CK(cudaDeviceSynchronize());
  int herr = 0;
  CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  if (herr) { printf("box %d floats: %d warps timed out waiting for the TMA (bad descriptor / semantics)\n", BOX_FLOATS, herr); return; }
  // correctness: the consumer summed the first float of every gathered row = its row id (mod 2^24)
  float* hout = (float*)malloc(148 * kWarps * 32 * sizeof(float));
  CK(cudaMemcpy(hout, out, 148 * kWarps * 32 * sizeof(float), cudaMemcpyDeviceToHost));
  double got = 0;
  for (int i = 0; i < 148 * kWarps * 32; ++i) got += hout[i];
  free(hout);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) launch();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
  printf("gather4 box %3d B/row, %d stages, %lld rows, smem %zu KB: %8.1f us  %6.2f G rows/s   checksum %s (got %.6g want %.6g)\n", BOX_FLOATS * 4, kStages,
         (long long)(n_stages * kRowsPerStage), smem >> 10, ms * 1e3, n_stages * kRowsPerStage / ms / 1e6,
         fabs(got - want) <= 1e-3 * want ? "ok" : "MISMATCH", got, want);
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int64_t table_rows = 200000000;   // 25.6 GB of 128-byte rows, like the packed DeepFM table
  float* table; CK(cudaMalloc(&table, table_rows * 128));
  fill_table<<<148 * 8, 256>>>(table, table_rows);
  const int64_t n = 65536 * 40;           // lookups per batch (multiple of 80)
  int* rows; CK(cudaMalloc(&rows, n * sizeof(int)));
  fill_rows<<<1024, 256>>>(rows, n, (uint32_t)table_rows, 99);
  float* out; CK(cudaMalloc(&out, 148 * kWarps * 32 * sizeof(float)));
  int* err; CK(cudaMalloc(&err, 4));
  double* dsum; CK(cudaMalloc(&dsum, 8)); CK(cudaMemset(dsum, 0, 8));
  expected_sum<<<256, 256>>>(rows, n / kRowsPerStage * kRowsPerStage, dsum);
  double want; CK(cudaMemcpy(&want, dsum, 8, cudaMemcpyDeviceToHost));
  CK(cudaDeviceSynchronize());
  run<20, 3>(encode, table, table_rows, rows, n, out, err, want);   // [v16 | w | 3 pad] = 80 B per row (what DeepFM needs)
  run<20, 2>(encode, table, table_rows, rows, n, out, err, want);
  run<32, 2>(encode, table, table_rows, rows, n, out, err, want);   // the whole 128-byte line
  run<16, 3>(encode, table, table_rows, rows, n, out, err, want);   // 64 B
  run<16, 5>(encode, table, table_rows, rows, n, out, err, want);
  run<8, 5>(encode, table, table_rows, rows, n, out, err, want);    // 32 B
  run<8, 10>(encode, table, table_rows, rows, n, out, err, want);
  return 0;
}
