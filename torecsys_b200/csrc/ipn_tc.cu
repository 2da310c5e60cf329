// a9 on the tensor pipe: InnerProductNetworkLayer = upper triangle of the per-sample Gram matrix X X^T.
//
// out[b, p(i,j)] = <x_i, x_j>, i < j.  One warp per sample: X (N, E) is staged in shared memory already split into TF32
// hi / lo planes; for every 16 x 8 output tile that touches the strict upper triangle (9 of 15 at N = 39) the warp
// issues E/8 k-steps of 3 mma.sync.m16n8k8 (3xTF32, fp32-accurate) with A and B fragments both read from those planes,
// then scatters the accumulator elements with i < j < N to their pair slots (row base table in shared memory) of a
// per-warp staging row in shared memory, which leaves for HBM as fully coalesced 128-byte stores (the accumulator
// layout would otherwise turn the (B, P) output into scattered 4-byte writes).
// The generic kernel in pairwise.cu does 2 LDS per FMA; this one is bound by the (B, P) output write.
#include <stdlib.h>

#include "common.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS>
__global__ void __launch_bounds__(kWarps * 32) ipn_tc_kernel(const float* __restrict__ x, int64_t batch, int fields,
                                                             float* __restrict__ out) {
  constexpr int E = 8 * KS;
  constexpr int kPitch = E + 4;   // rows 4 banks apart: the 8 rows x 4 columns of a fragment load hit 32 banks
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rows_pad = (fields + 15) & ~15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  int* rowbase = reinterpret_cast<int*>(smem_raw);                       // [rows_pad]: p(i, j) = rowbase[i] + j
  uint32_t* planes = reinterpret_cast<uint32_t*>(rowbase + rows_pad) + (size_t)warp * 2 * rows_pad * kPitch;
  uint32_t* xh = planes;                                                 // [rows_pad][kPitch] TF32 hi
  uint32_t* xl = planes + rows_pad * kPitch;                             // lo
  const int pairs = fields * (fields - 1) / 2;
  float* obuf = reinterpret_cast<float*>(reinterpret_cast<uint32_t*>(rowbase + rows_pad) +
                                         (size_t)kWarps * 2 * rows_pad * kPitch) + (size_t)warp * pairs;   // [pairs]
  for (int i = threadIdx.x; i < rows_pad; i += blockDim.x) rowbase[i] = i * (2 * fields - i - 1) / 2 - i - 1;
  // zero the padding rows once (they are never rewritten)
  for (int i = lane + fields * kPitch; i < rows_pad * kPitch; i += 32) {
    xh[i] = 0u;
    xl[i] = 0u;
  }
  __syncthreads();
  const int mt_count = rows_pad / 16, nt_count = (fields + 7) / 8;

  // The next sample's tile is prefetched into registers (up to kPre values per lane) while the current one is in its
  // MMA phase, so the global-load latency is off the warp's critical path; larger tiles load in place.
  constexpr int kPre = 24;
  const int tile = fields * E;
  const bool prefetch = tile <= 32 * kPre;
  const int64_t stride = (int64_t)gridDim.x * kWarps;
  float nxt[kPre];
  auto load_tile = [&](int64_t bb) {
    const float* src = x + bb * tile;
#pragma unroll
    for (int k = 0; k < kPre; ++k) {
      const int i = lane + 32 * k;
      nxt[k] = (bb < batch && i < tile) ? ldg_stream_f1(src + i) : 0.f;
    }
  };
  if (prefetch) load_tile((int64_t)blockIdx.x * kWarps + warp);
  for (int64_t b = (int64_t)blockIdx.x * kWarps + warp; b < batch; b += stride) {
    if (prefetch) {
#pragma unroll
      for (int k = 0; k < kPre; ++k) {
        const int i = lane + 32 * k;
        if (i < tile) {
          const int n = i / E, e = i - n * E;
          const float v = nxt[k];
          const uint32_t hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
          xh[n * kPitch + e] = hi;
          xl[n * kPitch + e] = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
        }
      }
      load_tile(b + stride);
    } else {
      const float* src = x + b * tile;
      for (int i = lane; i < tile; i += 32) {
        const int n = i / E, e = i - n * E;
        const float v = ldg_stream_f1(src + i);
        const uint32_t hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
        xh[n * kPitch + e] = hi;
        xl[n * kPitch + e] = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
      }
    }
    __syncwarp();
    float* dst = out + b * pairs;
    for (int mt = 0; mt < mt_count; ++mt) {
      // A fragments of this row block, all k-steps: rows 16mt+g / +8, columns 8ks+t / +4
      uint32_t ah[KS][4], al[KS][4];
      const int r0 = (16 * mt + g) * kPitch, r1 = r0 + 8 * kPitch;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int c = 8 * ks + t;
        ah[ks][0] = xh[r0 + c]; ah[ks][1] = xh[r1 + c]; ah[ks][2] = xh[r0 + c + 4]; ah[ks][3] = xh[r1 + c + 4];
        al[ks][0] = xl[r0 + c]; al[ks][1] = xl[r1 + c]; al[ks][2] = xl[r0 + c + 4]; al[ks][3] = xl[r1 + c + 4];
      }
      for (int nt = (16 * mt) / 8; nt < nt_count; ++nt) {   // tiles with some column j > some row i
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int rb = (8 * nt + g) * kPitch;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int c = 8 * ks + t;
          const uint32_t bh0 = xh[rb + c], bh1 = xh[rb + c + 4], bl0 = xl[rb + c], bl1 = xl[rb + c + 4];
          mma_tf32(acc, al[ks], bh0, bh1);
          mma_tf32(acc, ah[ks], bl0, bl1);
          mma_tf32(acc, ah[ks], bh0, bh1);
        }
        const int i0 = 16 * mt + g, i1 = i0 + 8, j0 = 8 * nt + 2 * t, j1 = j0 + 1;
        if (i0 < j0 && j0 < fields) obuf[rowbase[i0] + j0] = acc[0];
        if (i0 < j1 && j1 < fields) obuf[rowbase[i0] + j1] = acc[1];
        if (i1 < j0 && j0 < fields) obuf[rowbase[i1] + j0] = acc[2];
        if (i1 < j1 && j1 < fields) obuf[rowbase[i1] + j1] = acc[3];
      }
    }
    __syncwarp();
    for (int p = lane; p < pairs; p += 32) dst[p] = obuf[p];   // coalesced: 128 contiguous bytes per instruction
    __syncwarp();
  }
}

template <int KS>
int ipn_tc_dispatch(const float* x, int64_t batch, int fields, float* out, cudaStream_t s) {
  constexpr int E = 8 * KS;
  const int rows_pad = (fields + 15) & ~15;
  const size_t smem = (size_t)rows_pad * sizeof(int) + (size_t)kWarps * 2 * rows_pad * (E + 4) * sizeof(uint32_t) +
                      (size_t)kWarps * (fields * (fields - 1) / 2) * sizeof(float);
  if (smem > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    TRS_CUDA(cudaFuncSetAttribute(ipn_tc_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = grid_for(batch * 32, kWarps * 32, smem > 100 * 1024 ? 1 : (smem > 50 * 1024 ? 2 : 4));
  ipn_tc_kernel<KS><<<grid, kWarps * 32, smem, s>>>(x, batch, fields, out);
  return check_launch("ipn_tc_kernel");
}

}  // namespace

// returns TRS_ERR_UNSUPPORTED when the shape is not covered (caller falls back to the generic kernel)
int ipn_tc_launch(const float* x, int64_t batch, int fields, int embed, float* out, cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || fields < 2 || fields > 128) return TRS_ERR_UNSUPPORTED;
  switch (embed) {
    case 8: return ipn_tc_dispatch<1>(x, batch, fields, out, s);
    case 16: return ipn_tc_dispatch<2>(x, batch, fields, out, s);
    case 32: return ipn_tc_dispatch<4>(x, batch, fields, out, s);
    case 64: return ipn_tc_dispatch<8>(x, batch, fields, out, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

}  // namespace trs
