// Round-2 hardware probes for the DeepFM headline kernel redesign (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/r2_probe tools/r2_probe.cu
//   ./build/r2_probe [gather|mma|trunc|all]
// 1. gather : a persistent CTA per SM, producer warps stage random 128-byte-pitch rows global -> shared with cp.async
//             (LDGSTS) or one bulk copy (UBLKCP) per row, no compute: the rate the memory system gives as a function of
//             (a) which 16-byte chunks of the line are requested and (b) how many rows are in flight per SM.
// 2. trunc  : does tcgen05.mma kind::tf32 truncate or round raw fp32 operands (A from smem, A from TMEM)?
// 3. mma    : cycles per tcgen05.mma (M = 128, K = 8, kind::tf32) as a function of N, SS and TS forms, one or two
//             accumulators; chip-wide dense TF32 rate with N = 256.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../torecsys_b200/csrc/tc5.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

using namespace trs::tc5;

// ------------------------------------------------------------------------------------------------------------ gather
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void fill_rows(int* rows, int64_t n, uint32_t modulo, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    rows[i] = (int)(z % modulo);
  }
}

// MODE 0: chunks 0-4 (80 B, what round 1 reads)   1: chunks 0-7 (the whole line)   2: chunks 0-4 and 7 (all 4 sectors)
// MODE 3: chunks 0-3 (64 B)                        4: one 128-byte bulk copy per row   5: one 80-byte bulk copy per row
// RL = rows per lane and stage (a warp stages 32*RL rows per stage), STAGES-1 stages are in flight.
template <int MODE, int RL, int STAGES>
__global__ void __launch_bounds__(1024, 1) gather_pipe(const float4* __restrict__ table, const int* __restrict__ rows,
                                                        int64_t n, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kRowsPerStage = 32 * RL;
  constexpr int kSlot = (MODE == 4) ? 128 : 80;                        // bytes kept per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  unsigned char* my = smem + (size_t)warp * (STAGES * kRowsPerStage * kSlot + 512 + 64);
  unsigned char* dump = my + STAGES * kRowsPerStage * kSlot;           // 32 x 16 B: chunks nobody reads
  const uint32_t my_s = smem_u32(my), dump_s = smem_u32(dump) + lane * 16;
  const uint32_t bar0 = smem_u32(dump + 512);                          // STAGES <= 8 mbarriers (bulk modes)
  if (MODE >= 4 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  // the warp's blocks of kRowsPerStage consecutive entries of rows[]: block j = (cta * per_cta + j * warps + warp)
  const int64_t blocks = n / kRowsPerStage;
  const int64_t per_cta = (blocks + gridDim.x - 1) / gridDim.x;
  const int64_t first = blockIdx.x * per_cta;
  const int64_t last = first + per_cta < blocks ? first + per_cta : blocks;
  const int my_blocks = first + warp < last ? (int)((last - first - warp + warps - 1) / warps) : 0;
  const int sub = lane & 7, rsel = lane >> 3;
  bool active = true;
  if (MODE == 0) active = sub < 5;
  if (MODE == 2) active = sub < 5 || sub == 7;
  if (MODE == 3) active = sub < 4;
  int nxt[RL];
  auto load_ids = [&](int j) {
#pragma unroll
    for (int k = 0; k < RL; ++k)
      nxt[k] = j < my_blocks ? __ldg(rows + (first + (int64_t)j * warps + warp) * kRowsPerStage + 32 * k + lane) : 0;
  };
  auto issue = [&](int j, int stage) {
    if (j >= my_blocks) return;
    const uint32_t base = my_s + stage * kRowsPerStage * kSlot;
    if (MODE < 4) {
#pragma unroll
      for (int i = 0; i < RL * 8; ++i) {
        const int r = __shfl_sync(0xffffffffu, nxt[i >> 3], 4 * (i & 7) + rsel);
        const float4* src = table + (int64_t)r * 8 + sub;
        const uint32_t dst = sub < 5 ? base + (4 * i + rsel) * kSlot + sub * 16 : dump_s;
        if (active) cp16(dst, src);
      }
    } else {
      constexpr uint32_t kBytes = MODE == 4 ? 128 : 80;
      if (lane == 0) mbar_expect_tx(bar0 + 8 * stage, kRowsPerStage * kBytes);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < RL; ++k)
        bulk_g2s(base + (32 * k + lane) * kSlot, table + (int64_t)nxt[k] * 8, kBytes, bar0 + 8 * stage);
    }
  };
  float acc = 0.f;
  load_ids(0);
  for (int s = 0; s < STAGES - 1; ++s) {
    issue(s, s);
    load_ids(s + 1);
    if (MODE < 4) cp_commit();
  }
  int stage = 0, fill = STAGES - 1;
  uint32_t phase = 0;
  for (int j = 0; j < my_blocks; ++j) {
    issue(j + STAGES - 1, fill);
    load_ids(j + STAGES);
    if (MODE < 4) {
      cp_commit();
      cp_wait<STAGES - 1>();
    } else {
      mbar_wait(bar0 + 8 * stage, phase);
    }
    __syncwarp();
    acc += *reinterpret_cast<const float*>(my + (size_t)stage * kRowsPerStage * kSlot + lane * kSlot);   // touch
    __syncwarp();
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
    if (++fill == STAGES) fill = 0;
  }
  if (MODE < 4) cp_wait<0>();
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

template <class F>
static float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  CK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return ms / reps;
}

template <int MODE, int RL, int STAGES>
static void run_gather(const float4* table, const int* rows, int64_t n_small, int64_t n_big, int warps, float* out) {
  constexpr int kSlot = (MODE == 4) ? 128 : 80;
  const size_t smem = (size_t)warps * (STAGES * 32 * RL * kSlot + 512 + 64);
  if (smem > 227 * 1024) return;
  CK(cudaFuncSetAttribute(gather_pipe<MODE, RL, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int which = 0;
  auto launch_small = [&] {
    gather_pipe<MODE, RL, STAGES><<<148, warps * 32, smem>>>(table, rows + (which++ % 8) * n_small, n_small, out);
  };
  auto launch_big = [&] { gather_pipe<MODE, RL, STAGES><<<148, warps * 32, smem>>>(table, rows, n_big, out); };
  const float ms_small = time_ms(launch_small, 20);
  const float ms_big = time_ms(launch_big, 3);
  CK(cudaGetLastError());
  printf("gather mode=%d warps=%2d RL=%d stages=%d inflight_rows_per_sm=%5d smem_KB=%3zu | batch %7.1f us %6.2f Grows/s | 8xbatch %7.1f us %6.2f Grows/s\n",
         MODE, warps, RL, STAGES, warps * 32 * RL * (STAGES - 1), smem >> 10, ms_small * 1e3, n_small / ms_small * 1e-6,
         ms_big * 1e3, n_big / ms_big * 1e-6);
  fflush(stdout);
}

static void gather_suite(int quick) {
  const int64_t full_rows = 199999488;   // configs[1]: 39 x 5 128 192 rows, 128-byte pitch = 25.6 GB
  float4* table;
  CK(cudaMalloc(&table, (size_t)full_rows * 128));
  CK(cudaMemset(table, 0, (size_t)full_rows * 128));
  const int64_t n = 65536 * 39;
  int* rows;
  float* out;
  CK(cudaMalloc(&rows, n * 8 * sizeof(int)));
  CK(cudaMalloc(&out, 1 << 20));
  const int64_t footprints[] = {full_rows, 1 << 25, 1 << 23, 1 << 21, 1 << 19};   // 25.6 GB, 4 GB, 1 GB, 256 MB, 64 MB
  for (int fi = 0; fi < (quick ? 1 : 5); ++fi) {
    const int64_t table_rows = footprints[fi];
    printf("-- table footprint %.2f GB\n", table_rows * 128 / 1e9);
    fill_rows<<<1024, 256>>>(rows, n * 8, (uint32_t)table_rows, 12345);
    CK(cudaDeviceSynchronize());
#define G(MODE, RL, ST, W) run_gather<MODE, RL, ST>(table, rows, n, n * 8, W, out);
    G(0, 3, 2, 8) G(0, 3, 3, 8)
    if (fi == 0 && !quick) {
      G(0, 1, 2, 8) G(0, 1, 3, 8) G(0, 1, 5, 8) G(0, 1, 3, 4) G(0, 3, 2, 4) G(0, 3, 4, 4) G(0, 3, 6, 4)
      G(1, 3, 2, 8) G(1, 3, 3, 8) G(2, 3, 3, 8) G(3, 3, 3, 8)
      G(2, 1, 4, 16) G(2, 1, 6, 16)
      G(4, 3, 2, 8) G(5, 3, 3, 8)
    }
#undef G
  }
  cudaFree(table);
  cudaFree(rows);
  cudaFree(out);
}

// -------------------------------------------------------------------------------------------------- tcgen05: rounding
// D[m][n] = sum_k A[m][k] * B[n][k] with B[n][k] = (k == n % 8): D[m][n] is A[m][n % 8] as the tensor core sees it.
__global__ void __launch_bounds__(128, 1) trunc_probe(const float* __restrict__ a_in, float* __restrict__ d_ss,
                                                      float* __restrict__ d_ts) {
  __shared__ __align__(128) unsigned char a_s[2 * 128 * 16];
  __shared__ __align__(128) unsigned char b_s[2 * 16 * 16];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  // A: K-major no swizzle: chunk plane kc (4 floats of k) at kc * 128 * 16, row r at r * 16
  for (int kc = 0; kc < 2; ++kc)
    *reinterpret_cast<float4*>(a_s + kc * 2048 + r * 16) = *reinterpret_cast<const float4*>(a_in + r * 8 + 4 * kc);
  if (r < 16)
    for (int kc = 0; kc < 2; ++kc) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = r % 8;
      if (k / 4 == kc) (&v.x)[k % 4] = 1.f;
      *reinterpret_cast<float4*>(b_s + kc * 256 + r * 16) = v;
    }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(32 * warp) << 16);
  // A also into tensor memory (columns 32..39) for the TS form
  {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = j < 8 ? __float_as_uint(a_in[r * 8 + j]) : 0u;
    tmem_st16(lane_base + 32, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc(smem_u32(a_s), 2048, 128), bd = umma_desc(smem_u32(b_s), 256, 128);
    const uint32_t idesc = umma_idesc_tf32(16);
    umma_tf32(tmem + 0, ad, bd, idesc, 0);
    umma_tf32_ts(tmem + 16, tmem + 32, bd, idesc, 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(lane_base + 0, v);
  for (int j = 0; j < 16; ++j) d_ss[r * 16 + j] = __uint_as_float(v[j]);
  tmem_ld16(lane_base + 16, v);
  for (int j = 0; j < 16; ++j) d_ts[r * 16 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static void trunc_suite() {
  float h_a[128 * 8];
  uint32_t s = 12345;
  for (int i = 0; i < 128 * 8; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t bits = 0x3f800000u | (s >> 9);   // [1, 2) with a random 23-bit mantissa
    float f;
    memcpy(&f, &bits, 4);
    h_a[i] = (i & 1) ? -f : f;
  }
  float *a, *dss, *dts;
  CK(cudaMalloc(&a, sizeof(h_a)));
  CK(cudaMalloc(&dss, 128 * 16 * 4));
  CK(cudaMalloc(&dts, 128 * 16 * 4));
  CK(cudaMemcpy(a, h_a, sizeof(h_a), cudaMemcpyHostToDevice));
  trunc_probe<<<1, 128>>>(a, dss, dts);
  CK(cudaDeviceSynchronize());
  float h_ss[128 * 16], h_ts[128 * 16];
  CK(cudaMemcpy(h_ss, dss, sizeof(h_ss), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_ts, dts, sizeof(h_ts), cudaMemcpyDeviceToHost));
  for (int form = 0; form < 2; ++form) {
    const float* d = form == 0 ? h_ss : h_ts;
    int n_trunc = 0, n_rna = 0, n_exact = 0, n_other = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 16; ++n) {
        uint32_t bits;
        memcpy(&bits, &h_a[m * 8 + n % 8], 4);
        uint32_t tb = bits & 0xffffe000u, rb = (bits + 0x1000u) & 0xffffe000u, db;
        memcpy(&db, &d[m * 16 + n], 4);
        if (db == bits) ++n_exact;
        else if (db == tb && db == rb) { ++n_trunc; ++n_rna; }
        else if (db == tb) ++n_trunc;
        else if (db == rb) ++n_rna;
        else ++n_other;
      }
    printf("trunc %s: of 2048 outputs  == fp32 exact %d | == truncation %d | == round-to-nearest %d | neither %d  "
           "(values where both agree are counted in both)\n", form == 0 ? "A from smem (SS)" : "A from TMEM (TS)",
           n_exact, n_trunc, n_rna, n_other);
  }
  cudaFree(a); cudaFree(dss); cudaFree(dts);
}

// -------------------------------------------------------------------------------------------- tcgen05: cost per MMA
// One CTA per SM; one thread issues `iters` MMAs (M = 128, K = 8) of width N round-robin over NACC accumulators,
// eight per loop iteration with compile-time accumulator addresses (the issuing thread must not be the limiter).
// FORM 0: A and B from shared memory; 1: A from tensor memory; 2: alternate (SS N = n, SS N = n2) pairs;
// FORM 3: alternate (SS N = n, TS N = n2) pairs.  swz = 1: SWIZZLE_128B descriptors (layout_type 2) for timing only.
template <int FORM, int NACC>
__global__ void __launch_bounds__(128, 1) mma_stream(int n, int n2, int swz, int iters, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* a_s = sm;                      // 2 planes x 128 rows x 16 B (or 128 rows x 128 B swizzled: 16 KB)
  unsigned char* b_s = sm + 16384;              // up to 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // (elect.sync, not `threadIdx.x == 0`: under a lane-id branch ptxas wraps every MMA in an ELECT loop, ~45 cycles each)
  if (warp == 0 && elect_one()) {
    uint64_t ad, bd, bd2;
    if (swz) {   // K-major SWIZZLE_128B: rows of 128 B, 8-row atoms 1024 B apart (SBO), LBO unused (1)
      const uint64_t sw = static_cast<uint64_t>(2) << 61;
      ad = (umma_desc(smem_u32(a_s), 16, 1024)) | sw;
      bd = (umma_desc(smem_u32(b_s), 16, 1024)) | sw;
      bd2 = bd;
    } else {
      ad = umma_desc(smem_u32(a_s), 2048, 128);
      bd = umma_desc(smem_u32(b_s), n * 16, 128);
      bd2 = umma_desc(smem_u32(b_s), n2 * 16, 128);
    }
    const uint32_t idesc = umma_idesc_tf32(n), idesc2 = umma_idesc_tf32(n2 > 0 ? n2 : 8);
    const uint32_t stride = FORM >= 2 ? n + n2 : n;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tmem + (u % NACC) * stride;
        if (FORM == 0) umma_tf32(d, ad, bd, idesc, 1);
        else if (FORM == 1) umma_tf32_ts(d, tmem + 480, bd, idesc, 1);
        else if (FORM == 2) {
          umma_tf32(d, ad, bd, idesc, 1);
          umma_tf32(d + n, ad, bd2, idesc2, 1);
        } else {
          umma_tf32(d, ad, bd, idesc, 1);
          umma_tf32_ts(d + n, tmem + 480, bd2, idesc2, 1);
        }
      }
    }
    umma_commit(smem_u32(bar));
    mbar_wait(smem_u32(bar), 0);
    const long long t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int FORM, int NACC>
static void run_mma(int n, int n2, int swz, int grid, long long* cyc) {
  const int iters = 4096;
  const size_t smem = 16384 + 32768 + 64;
  CK(cudaFuncSetAttribute(mma_stream<FORM, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  float ms = time_ms([&] { mma_stream<FORM, NACC><<<grid, 128, smem>>>(n, n2, swz, iters, cyc); }, 3);
  CK(cudaGetLastError());
  long long h[148];
  CK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += h[i];
  mean /= grid;
  const int per_iter = FORM >= 2 ? 2 : 1;
  const double flop = 2.0 * 128 * 8 * (FORM >= 2 ? n + n2 : n) * (double)iters * grid;
  printf("mma form=%d swz=%d N=%3d N2=%3d acc=%d grid=%3d | %7.1f cycles/MMA | kernel %8.1f us | %7.1f TFLOP/s (tf32, chip)\n",
         FORM, swz, n, n2, NACC, grid, mean / iters / per_iter, ms * 1e3, flop / (ms * 1e-3) * 1e-12);
  fflush(stdout);
}

static void mma_suite() {
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  for (int n : {16, 32, 64, 128, 256}) {
    run_mma<0, 1>(n, 0, 0, 148, cyc);
    run_mma<0, 2>(n, 0, 0, 148, cyc);
    run_mma<0, 2>(n, 0, 1, 148, cyc);
    if (n <= 128) run_mma<1, 2>(n, 0, 0, 148, cyc);
  }
  run_mma<0, 4>(16, 0, 0, 148, cyc);
  run_mma<0, 4>(32, 0, 0, 148, cyc);
  run_mma<2, 1>(32, 16, 0, 148, cyc);
  run_mma<2, 2>(32, 16, 0, 148, cyc);
  run_mma<3, 1>(32, 16, 0, 148, cyc);
  run_mma<3, 2>(32, 16, 0, 148, cyc);
  run_mma<0, 2>(256, 0, 0, 1, cyc);
  cudaFree(cyc);
}

// ------------------------------------------------------------------------------------- tensor memory: ld / st rate
// One CTA per SM, `warps` warps (4 per TMEM lane quarter group); every warp issues `iters` rounds of U tcgen05.st.x16
// (MODE 0), tcgen05.ld.x16 (MODE 1), or a ld + two st per round like the epilogue of dcn_tc5.cu (MODE 2), one wait per
// round.  Reports bytes per cycle and SM.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) tmem_rate(int iters, long long* __restrict__ cycles, float* __restrict__ sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + (warp >> 2) * 64;
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = threadIdx.x + j;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      tmem_st16(base, v);
      tmem_st16(base + 16, v);
      tmem_st16(base + 32, v);
      tmem_st16(base + 48, v);
      tmem_st_wait();
    } else if (MODE == 1) {
      uint32_t r[16];
      tmem_ld16(base, r);
      acc += __uint_as_float(r[3]);
      tmem_ld16(base + 16, r);
      acc += __uint_as_float(r[5]);
      tmem_ld16(base + 32, r);
      acc += __uint_as_float(r[7]);
      tmem_ld16(base + 48, r);
      acc += __uint_as_float(r[9]);
    } else {
      uint32_t r[16];
      tmem_ld16(base, r);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = r[j] + 1;
      tmem_st16(base + 16, v);
      tmem_st16(base + 32, v);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.f) sink[threadIdx.x] = acc + v[0];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

static void tmem_suite() {
  long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 148 * sizeof(long long)));
  CK(cudaMalloc(&sink, 4096));
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      if (warps > 8 && false) continue;
      if (mode == 0) tmem_rate<0><<<148, warps * 32>>>(iters, cyc, sink);
      else if (mode == 1) tmem_rate<1><<<148, warps * 32>>>(iters, cyc, sink);
      else tmem_rate<2><<<148, warps * 32>>>(iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      long long h[148];
      CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double mean = 0;
      for (int i = 0; i < 148; ++i) mean += h[i];
      mean /= 148;
      const double bytes = (mode == 2 ? 3.0 : 4.0) * 2048 * warps * iters;   // 32 lanes x 16 columns x 4 B per x16 op
      printf("tmem %s  warps %2d | %8.1f cycles per round | %7.1f B/cycle/SM\n",
             mode == 0 ? "st x4 + wait   " : mode == 1 ? "ld x4 (+waits) " : "ld, st x2, wait", warps, mean / iters,
             bytes / mean);
    }
  fflush(stdout);
}

int main(int argc, char** argv) {
  const char* what = argc > 1 ? argv[1] : "all";
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("device: %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  if (!strcmp(what, "trunc") || !strcmp(what, "all")) trunc_suite();
  if (!strcmp(what, "mma") || !strcmp(what, "all")) mma_suite();
  if (!strcmp(what, "tmem") || !strcmp(what, "all")) tmem_suite();
  if (!strcmp(what, "gather") || !strcmp(what, "all")) gather_suite(0);
  if (!strcmp(what, "gatherq")) gather_suite(1);
  return 0;
}
