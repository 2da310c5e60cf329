// Round-2 probe: how fast can a kernel on GPU 0 pull random contiguous chunks out of GPU 1's HBM over NVLink, as a
// function of the chunk size and of how the chunk is requested (16-byte cp.async per lane vs one bulk copy per chunk)?
// Decides the exchange granularity of the row-sharded FFM (configs[4]) -- VERDICT r1 item 3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/peer_probe tools/peer_probe.cu
//   ./build/peer_probe            (needs 2 GPUs with peer access; with 1 GPU it measures the local numbers only)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../torecsys_b200/csrc/tc5.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

using namespace trs::tc5;

__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_n(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z *= 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Every warp owns a ring of `stages` slots of `group` chunks each.  Unit u of the warp = `group` random chunks of
// `chunk` bytes (chunk-aligned, `pitch` bytes apart in the table).  bulk = 0: the 32 lanes cover the unit's 16-byte
// pieces with cp.async; bulk = 1: lane g issues one cp.async.bulk for chunk g, completion on the slot's mbarrier.
__global__ void __launch_bounds__(256, 1) peer_gather(const unsigned char* __restrict__ table, uint32_t rows, int pitch,
                                                       int chunk, int group, int stages, int bulk, int units_per_warp,
                                                       float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int slot_bytes = group * chunk;
  unsigned char* my = smem + (size_t)warp * ((size_t)stages * slot_bytes + 128);
  const uint32_t my_s = smem_u32(my);
  const uint32_t bar0 = smem_u32(my + (size_t)stages * slot_bytes);
  if (bulk && lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const uint64_t wid = (uint64_t)blockIdx.x * warps + warp;
  const int pieces = slot_bytes / 16, per_chunk = chunk / 16;
  auto issue = [&](int u, int stage) {
    if (u >= units_per_warp) return;
    const uint32_t base = my_s + stage * slot_bytes;
    if (!bulk) {
      for (int p = lane; p < pieces; p += 32) {
        const int g = p / per_chunk, q = p - g * per_chunk;
        const uint32_t r = (uint32_t)(mix((wid * units_per_warp + u) * 64 + g) % rows);
        cp16(base + p * 16, table + (size_t)r * pitch + q * 16);
      }
    } else {
      if (lane == 0) mbar_expect_tx(bar0 + 8 * stage, slot_bytes);
      __syncwarp();
      for (int g = lane; g < group; g += 32) {
        const uint32_t r = (uint32_t)(mix((wid * units_per_warp + u) * 64 + g) % rows);
        bulk_g2s(base + g * chunk, table + (size_t)r * pitch, chunk, bar0 + 8 * stage);
      }
    }
  };
  float acc = 0.f;
  for (int s = 0; s < stages - 1; ++s) {
    issue(s, s);
    if (!bulk) cp_commit();
  }
  int stage = 0, fill = stages - 1;
  uint32_t phase = 0;
  for (int u = 0; u < units_per_warp; ++u) {
    issue(u + stages - 1, fill);
    if (!bulk) {
      cp_commit();
      cp_wait_n(stages - 1);
    } else {
      mbar_wait(bar0 + 8 * stage, phase);
    }
    __syncwarp();
    acc += *reinterpret_cast<const float*>(my + (size_t)stage * slot_bytes + (lane * 16) % slot_bytes);
    __syncwarp();
    if (++stage == stages) { stage = 0; phase ^= 1; }
    if (++fill == stages) fill = 0;
  }
  if (!bulk) cp_wait_n(0);
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

// plain register loads: lane-contiguous float4 over the chunk, `unroll` independent loads in flight per lane
__global__ void __launch_bounds__(256, 2) peer_ldg(const unsigned char* __restrict__ table, uint32_t rows, int pitch,
                                                    int chunk, int units_per_warp, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const uint64_t wid = (uint64_t)blockIdx.x * warps + warp;
  const int per_chunk = chunk / 16;
  float acc = 0.f;
  for (int u = 0; u < units_per_warp; u += 1) {
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int p = k * 32 + lane;
      const int g = p / per_chunk, q = p - g * per_chunk;
      const uint32_t r = (uint32_t)(mix((wid * units_per_warp + u) * 64 + g) % rows);
      v[k] = __ldcg(reinterpret_cast<const float4*>(table + (size_t)r * pitch + q * 16));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].w;
  }
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

static float run(const unsigned char* table, uint32_t rows, int pitch, int chunk, int group, int stages, int bulk,
                 int warps, int64_t total_bytes, float* out) {
  const int ctas = 148;
  const int units = (int)(total_bytes / ((int64_t)ctas * warps * group * chunk));
  const size_t smem = (size_t)warps * ((size_t)stages * group * chunk + 128);
  if (smem > 227 * 1024) return -1.f;
  CK(cudaFuncSetAttribute(peer_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  peer_gather<<<ctas, warps * 32, smem>>>(table, rows, pitch, chunk, group, stages, bulk, units, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < 3; ++i) peer_gather<<<ctas, warps * 32, smem>>>(table, rows, pitch, chunk, group, stages, bulk, units, out);
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)ctas * warps * units * group * chunk;
  return (float)(bytes * 3 / (ms * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  const size_t table_bytes = (size_t)8 << 30;   // 8 GB: the per-GPU share of configs[4]
  unsigned char* tables[2] = {nullptr, nullptr};
  CK(cudaSetDevice(0));
  CK(cudaMalloc(&tables[0], table_bytes));
  CK(cudaMemset(tables[0], 1, table_bytes));
  if (ndev >= 2) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    printf("devices %d, peer access 0->1: %d\n", ndev, can);
    if (can) {
      CK(cudaSetDevice(1));
      CK(cudaMalloc(&tables[1], table_bytes));
      CK(cudaMemset(tables[1], 1, table_bytes));
      CK(cudaDeviceSynchronize());
      CK(cudaSetDevice(0));
      CK(cudaDeviceEnablePeerAccess(1, 0));
    }
  }
  float* out;
  CK(cudaMalloc(&out, 4096));
  const int64_t total = (int64_t)2 << 30;   // bytes pulled per launch
  if (tables[1]) {   // reference: the copy engine
    unsigned char* dst;
    CK(cudaMalloc(&dst, (size_t)1 << 30));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    CK(cudaMemcpyPeer(dst, 0, tables[1], 1, (size_t)1 << 30));
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < 4; ++i) CK(cudaMemcpyPeerAsync(dst, 0, tables[1], 1, (size_t)1 << 30, 0));
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemcpyPeer 1 GiB x4: %.1f GB/s\n", 4.0 * (1 << 30) / (ms * 1e-3) / 1e9);
    CK(cudaFree(dst));
  }
  for (int where = 0; where < 2; ++where) {
    if (!tables[where]) continue;
    printf("== table on GPU %d (%s), kernel on GPU 0; GB/s of useful bytes\n", where, where ? "NVLink peer" : "local HBM");
    printf("%6s %6s %6s %6s %5s %6s | %8s\n", "chunk", "pitch", "group", "stages", "bulk", "warps", "GB/s");
    struct Cfg { int chunk, pitch, group, stages, bulk, warps; };
    const Cfg cfgs[] = {
        {64, 64, 16, 4, 0, 8},     {64, 64, 32, 4, 0, 8},     {64, 64, 32, 6, 0, 8},   {64, 64, 16, 4, 1, 8},
        {128, 128, 16, 4, 0, 8},   {128, 128, 16, 4, 1, 8},   {256, 256, 8, 4, 0, 8},  {256, 256, 8, 4, 1, 8},
        {320, 320, 5, 4, 0, 8},    {320, 320, 5, 8, 0, 8},    {320, 320, 10, 4, 0, 8}, {320, 320, 5, 4, 1, 8},
        {320, 320, 5, 8, 1, 8},    {320, 320, 10, 4, 1, 8},   {320, 320, 10, 6, 1, 8}, {320, 384, 5, 8, 1, 8},
        {384, 384, 5, 8, 1, 8},    {384, 384, 5, 8, 0, 8},    {640, 640, 5, 4, 1, 8},  {640, 640, 5, 4, 0, 8},
        {1280, 1280, 4, 4, 1, 8},  {1280, 1280, 4, 4, 0, 8},  {2496, 2496, 2, 4, 1, 8}, {2496, 2496, 2, 4, 0, 8},
        {320, 320, 5, 4, 1, 4},    {320, 320, 5, 8, 1, 4},    {320, 320, 5, 4, 1, 16}, {320, 320, 10, 3, 1, 16},
        {320, 320, 5, 2, 1, 8},    {320, 320, 5, 3, 1, 8},
    };
    for (const Cfg& c : cfgs) {
      const uint32_t rows = (uint32_t)(table_bytes / c.pitch);
      const float g = run(tables[where], rows, c.pitch, c.chunk, c.group, c.stages, c.bulk, c.warps, total, out);
      printf("%6d %6d %6d %6d %5d %6d | %8.1f\n", c.chunk, c.pitch, c.group, c.stages, c.bulk, c.warps, g);
      fflush(stdout);
    }
    // register loads
    for (int chunk : {64, 128, 320, 1280}) {
      const int pitch = chunk;
      const uint32_t rows = (uint32_t)(table_bytes / pitch);
      const int units = (int)(total / ((int64_t)296 * 8 * 4096));
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      peer_ldg<<<296, 256>>>(tables[where], rows, pitch, chunk, units, out);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      peer_ldg<<<296, 256>>>(tables[where], rows, pitch, chunk, units, out);
      cudaEventRecord(b);
      CK(cudaEventSynchronize(b));
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      printf("ld.global.cg.v4 x8 per lane, chunk %4d: %8.1f GB/s\n", chunk, 296.0 * 8 * units * 4096 / (ms * 1e-3) / 1e9);
    }
  }
  // ---- both directions at once: GPU 0 pulls from GPU 1 while GPU 1 pulls from GPU 0 (what an all-to-all exchange sees:
  // read requests of one flow share the link direction with the response data of the other)
  if (tables[1]) {
    CK(cudaSetDevice(1));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    float* out1;
    CK(cudaMalloc(&out1, 4096));
    CK(cudaFuncSetAttribute(peer_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaSetDevice(0));
    printf("== both directions at once (bulk copies, 8 warps): GB/s of useful bytes INTO each GPU\n");
    struct Cfg { int chunk, pitch, group, stages; };
    const Cfg cfgs[] = {{64, 64, 16, 4}, {128, 128, 16, 4}, {256, 256, 8, 4}, {320, 320, 10, 4}, {384, 384, 5, 8},
                        {640, 640, 5, 4}, {1280, 1280, 4, 4}};
    for (const Cfg& c : cfgs) {
      const int warps = 8, ctas = 148;
      const uint32_t rows = (uint32_t)(table_bytes / c.pitch);
      const int units = (int)(total / ((int64_t)ctas * warps * c.group * c.chunk));
      const size_t smem = (size_t)warps * ((size_t)c.stages * c.group * c.chunk + 128);
      cudaEvent_t a0, b0, a1, b1;
      CK(cudaSetDevice(0)); cudaEventCreate(&a0); cudaEventCreate(&b0);
      CK(cudaSetDevice(1)); cudaEventCreate(&a1); cudaEventCreate(&b1);
      for (int rep = 0; rep < 2; ++rep) {   // first round warms up
        CK(cudaSetDevice(0));
        cudaEventRecord(a0);
        for (int i = 0; i < 3; ++i) peer_gather<<<ctas, warps * 32, smem>>>(tables[1], rows, c.pitch, c.chunk, c.group, c.stages, 1, units, out);
        cudaEventRecord(b0);
        CK(cudaSetDevice(1));
        cudaEventRecord(a1);
        for (int i = 0; i < 3; ++i) peer_gather<<<ctas, warps * 32, smem>>>(tables[0], rows, c.pitch, c.chunk, c.group, c.stages, 1, units, out1);
        cudaEventRecord(b1);
        CK(cudaSetDevice(0)); CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
      }
      float ms0, ms1;
      cudaEventElapsedTime(&ms0, a0, b0);
      cudaEventElapsedTime(&ms1, a1, b1);
      const double bytes = (double)ctas * warps * units * c.group * c.chunk * 3;
      printf("chunk %5d pitch %5d | into GPU 0 %7.1f GB/s | into GPU 1 %7.1f GB/s\n", c.chunk, c.pitch,
             bytes / (ms0 * 1e-3) / 1e9, bytes / (ms1 * 1e-3) / 1e9);
      fflush(stdout);
    }
    CK(cudaSetDevice(0));
  }
  return 0;
}
