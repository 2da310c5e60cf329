// a10 on the tensor pipe: BilinearInteractionLayer, both weight types, bound by the (B, P, E) output write.
//
// out[b, p(i,j), :] = (x_i W_p) * x_j + bias_p, with W_p = W for bilinear_type 'all' and one (E, E) matrix per pair
// for 'each' (reference bilinear_interaction.py:11-76 / :82-149: matmul(input1, weight) * input2 + bias).
//
// One CTA owns a tile of 32 samples: x (32, N, E) is staged once in shared memory (row pitch = 4 mod 32 banks, so an
// mma A-fragment load touches 32 different banks).  A warp owns a CONTIGUOUS range of pairs: consecutive pairs share
// the left field i, so the 3xTF32 hi/lo A fragments of x_i (two m16 tiles) stay in registers until i changes; per
// pair the warp loads the W_p B fragments from L2 (759 KB of pair weights, re-read once per 32 samples), splits them,
// issues 2 x KS x NT x 3 mma.sync.m16n8k8 (fp32-accurate) and writes acc * x_j + bias as float2 -- the 64 bytes of a
// (sample, pair) row back to back with the neighbouring pair's, so L2 merges full lines before they go to HBM.
// The generic kernels in pairwise.cu (one FFMA per LDS pair, 4-byte stores; a CTA per pair with scattered 64-byte
// writes) stay as the fallback for other embed sizes.
#include <stdlib.h>

#include "common.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;
constexpr int kSamples = 32;   // two m16 tiles

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

// smallest pitch >= width + 4 with pitch = 4 (mod 32): the 8 rows x 4 columns of a fragment load hit 32 banks
__host__ __device__ inline int row_pitch(int width) { return width + 4 + ((32 - width % 32) % 32); }

// OPN = true: the same tiles compute OuterProductNetworkLayer's 'mat' kernel (outer_product_network.py:80-131),
//     out[b, p] = sum_h x_j[h] * (sum_e kernel[h, p, e] * x_i[e]),    kernel (E, P, E),
// i.e. the bilinear form with W_p[k = e][o = h] = kernel[h, p, e] read in place (8 consecutive floats per fragment
// pair) and the product with x_j summed over o: four lanes add their columns with two shuffles, one 4-byte store per
// (sample, pair) -- the (B, P) output is 16x smaller than the bilinear layer's, so this form is bound by the tensor pipe.
template <int E, bool OPN>
__global__ void __launch_bounds__(kWarps * 32, E <= 16 ? 2 : 1) bilinear_tc_kernel(const float* __restrict__ x,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias, int each_type,
                                                                     int64_t batch, int fields,
                                                                     int64_t out_stride, float* __restrict__ out) {
  constexpr int KS = E / 8, NT = E / 8;   // out_stride: floats between the output rows of consecutive samples
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int width = fields * E;
  const int pairs = fields * (fields - 1) / 2;
  const int xpitch = row_pitch(width);                            // = 4 (mod 32 banks)
  float* xs = reinterpret_cast<float*>(smem_raw);                 // [kSamples][xpitch]
  int* ptab = reinterpret_cast<int*>(xs + (size_t)kSamples * xpitch);   // [pairs] (i << 16) | j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  // this warp's pair range
  const int per_warp = (pairs + kWarps - 1) / kWarps;
  const int p_begin = warp * per_warp;
  const int p_end = p_begin + per_warp < pairs ? p_begin + per_warp : pairs;
  const int64_t w_stride = each_type ? (int64_t)E * E : 0;
  const int b_stride = each_type ? E : 0;

  const int64_t tiles = (batch + kSamples - 1) / kSamples;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t b0 = tile * kSamples;
    const int valid = static_cast<int>(batch - b0 < kSamples ? batch - b0 : kSamples);
    __syncthreads();   // previous tile fully consumed
    {
      const int chunks = width >> 2;   // E % 8 == 0 -> rows are float4 multiples
      const float4* src = reinterpret_cast<const float4*>(x + b0 * width);
      for (int c = threadIdx.x; c < kSamples * chunks; c += blockDim.x) {
        const int s = c / chunks, k = c - s * chunks;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < valid) v = ldg_stream_f4(src + (int64_t)s * chunks + k);
        *reinterpret_cast<float4*>(xs + s * xpitch + 4 * k) = v;
      }
    }
    __syncthreads();
    int cur_i = -1;
    uint32_t ah[2][KS][4], al[2][KS][4];
    for (int p = p_begin; p < p_end; ++p) {
      const int ij = ptab[p];
      const int i = ij >> 16, j = ij & 0xffff;
      if (i != cur_i) {   // A fragments of x_i for both m-tiles: rows g / g+8 (+16 mt), columns 8ks + t / + 4
        cur_i = i;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* r0 = xs + (16 * mt + g) * xpitch + i * E;
          const float* r1 = r0 + 8 * xpitch;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            split_tf32(r0[8 * ks + t], ah[mt][ks][0], al[mt][ks][0]);
            split_tf32(r1[8 * ks + t], ah[mt][ks][1], al[mt][ks][1]);
            split_tf32(r0[8 * ks + t + 4], ah[mt][ks][2], al[mt][ks][2]);
            split_tf32(r1[8 * ks + t + 4], ah[mt][ks][3], al[mt][ks][3]);
          }
        }
      }
      // B fragments of W_p: B[k][n] = W_p[k, o = 8nt + n]; b0 = B[8ks + t][g], b1 = B[8ks + t + 4][g]
      const float* wp = w + p * w_stride;
      uint32_t bh[KS][NT][2], bl[KS][NT][2];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          if constexpr (OPN) {   // B[k = e][n = h] = kernel[h, p, e]
            const float* kp = w + ((int64_t)(8 * nt + g) * pairs + p) * E + 8 * ks + t;
            split_tf32(__ldg(kp), bh[ks][nt][0], bl[ks][nt][0]);
            split_tf32(__ldg(kp + 4), bh[ks][nt][1], bl[ks][nt][1]);
          } else {
            split_tf32(__ldg(wp + (8 * ks + t) * E + 8 * nt + g), bh[ks][nt][0], bl[ks][nt][0]);
            split_tf32(__ldg(wp + (8 * ks + t + 4) * E + 8 * nt + g), bh[ks][nt][1], bl[ks][nt][1]);
          }
        }
      float2 bv[NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        bv[nt] = bias ? __ldg(reinterpret_cast<const float2*>(bias + (int64_t)p * b_stride + 8 * nt + 2 * t))
                      : make_float2(0.f, 0.f);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            mma_tf32(acc[nt], al[mt][ks], bh[ks][nt][0], bh[ks][nt][1]);
            mma_tf32(acc[nt], ah[mt][ks], bl[ks][nt][0], bl[ks][nt][1]);
            mma_tf32(acc[nt], ah[mt][ks], bh[ks][nt][0], bh[ks][nt][1]);
          }
        }
        // acc[nt] = {(row g, cols 2t, 2t+1), (row g+8, cols 2t, 2t+1)} of this m-tile
        const int s0 = 16 * mt + g, s1 = s0 + 8;
        const float* xj0 = xs + s0 * xpitch + j * E;
        const float* xj1 = xs + s1 * xpitch + j * E;
        if constexpr (OPN) {
          float v0 = 0.f, v1 = 0.f;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int c = 8 * nt + 2 * t;
            const float2 a0 = *reinterpret_cast<const float2*>(xj0 + c);
            const float2 a1 = *reinterpret_cast<const float2*>(xj1 + c);
            v0 = fmaf(acc[nt][0], a0.x, fmaf(acc[nt][1], a0.y, v0));
            v1 = fmaf(acc[nt][2], a1.x, fmaf(acc[nt][3], a1.y, v1));
          }
          v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
          v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
          if (t == 0 && s0 < valid) out[(b0 + s0) * pairs + p] = v0;
          if (t == 0 && s1 < valid) out[(b0 + s1) * pairs + p] = v1;
          continue;
        }
        float* o0 = out + (b0 + s0) * out_stride + (int64_t)p * E;
        float* o1 = out + (b0 + s1) * out_stride + (int64_t)p * E;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + 2 * t;
          const float2 a0 = *reinterpret_cast<const float2*>(xj0 + c);
          const float2 a1 = *reinterpret_cast<const float2*>(xj1 + c);
          if (s0 < valid)
            *reinterpret_cast<float2*>(o0 + c) =
                make_float2(fmaf(acc[nt][0], a0.x, bv[nt].x), fmaf(acc[nt][1], a0.y, bv[nt].y));
          if (s1 < valid)
            *reinterpret_cast<float2*>(o1 + c) =
                make_float2(fmaf(acc[nt][2], a1.x, bv[nt].x), fmaf(acc[nt][3], a1.y, bv[nt].y));
        }
      }
    }
  }
}

template <int E, bool OPN>
int bilinear_tc_dispatch(const float* x, const float* w, const float* bias, int each_type, int64_t batch, int fields,
                         int64_t out_stride, float* out, cudaStream_t s) {
  const int xpitch = row_pitch(fields * E);
  const int pairs = fields * (fields - 1) / 2;
  const size_t smem = (size_t)kSamples * xpitch * sizeof(float) + (size_t)pairs * sizeof(int);
  if (smem > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    TRS_CUDA(cudaFuncSetAttribute(bilinear_tc_kernel<E, OPN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t tiles = (batch + kSamples - 1) / kSamples;
  const int per_sm = smem > 110 * 1024 ? 1 : 2;
  const int grid = static_cast<int>(tiles < (int64_t)kNumSMs * per_sm ? tiles : (int64_t)kNumSMs * per_sm);
  bilinear_tc_kernel<E, OPN><<<grid, kWarps * 32, smem, s>>>(x, w, bias, each_type, batch, fields, out_stride, out);
  return check_launch("bilinear_tc_kernel");
}

}  // namespace

// returns TRS_ERR_UNSUPPORTED when the shape is not covered (caller falls back to the generic kernels)
int bilinear_tc_launch(const float* x, const float* w, const float* bias, int each_type, int64_t batch, int fields,
                       int embed, int64_t out_stride, float* out, cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || fields < 2 || fields > 1024 || !aligned16(x) || !aligned16(out) || (out_stride & 3) ||
      (bias && (reinterpret_cast<uintptr_t>(bias) & 7u)))
    return TRS_ERR_UNSUPPORTED;
  switch (embed) {
    case 8: return bilinear_tc_dispatch<8, false>(x, w, bias, each_type, batch, fields, out_stride, out, s);
    case 16: return bilinear_tc_dispatch<16, false>(x, w, bias, each_type, batch, fields, out_stride, out, s);
    case 32: return bilinear_tc_dispatch<32, false>(x, w, bias, each_type, batch, fields, out_stride, out, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

// OuterProductNetworkLayer, kernel_type 'mat', on the same tiles (TRS_ERR_UNSUPPORTED -> the FFMA kernel of pnn_senet.cu)
int opn_mat_tc_launch(const float* x, const float* kernel, int64_t batch, int fields, int embed, float* out,
                      cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || fields < 2 || fields > 1024 || !aligned16(x)) return TRS_ERR_UNSUPPORTED;
  switch (embed) {
    case 8: return bilinear_tc_dispatch<8, true>(x, kernel, nullptr, 1, batch, fields, 0, out, s);
    case 16: return bilinear_tc_dispatch<16, true>(x, kernel, nullptr, 1, batch, fields, 0, out, s);
    case 32: return bilinear_tc_dispatch<32, true>(x, kernel, nullptr, 1, batch, fields, 0, out, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

}  // namespace trs
