// a11 on the tensor pipe: AttentionalFactorizationMachineLayer forward for embed % 8 == 0, attn <= 64.
//
// Per sample: P = N(N-1)/2 pair products prod_p = x_i * x_j (P x E), hidden = relu(prod W1^T + b1) (P x A),
// score_p = hidden . w2 + b2, softmax over the P pairs, out = sum_p s_p prod_p.  The P x E x A contraction
// (190 K MACs per sample at N = 39, E = A = 16) is 95 % of the work and FFMA-bound in the generic kernel
// (pairwise.cu: 5 TFLOP/s); here it runs as mma.sync.m16n8k8 TF32 with the 3xTF32 split:
//   * one CTA per sample (8 warps), x (N, E) staged once in shared memory;
//   * a warp takes 16 pairs at a time: each lane builds its A-fragment elements prod[p][k] = x_i[k] * x_j[k] straight
//     from shared memory (rows g / g+8 = pairs, columns t / t+4), splits them hi/lo, multiplies with the pre-split
//     W1 fragments held in REGISTERS (A/8 n-tiles x E/8 k-steps), applies bias + ReLU + the w2 dot in the
//     accumulator layout and reduces the score over the 4 lanes of a row;
//   * scores go to shared memory; block softmax; the weighted sum re-forms the products (cheaper than storing them).
#include <stdlib.h>

#include "common.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// KS = embed / 8 k-steps, NT = ceil(attn / 8) n-tiles
template <int KS, int NT>
__global__ void __launch_bounds__(kWarps * 32) afm_tc_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                             const float* __restrict__ b1,
                                                             const float* __restrict__ w2,
                                                             const float* __restrict__ b2, int64_t batch, int fields,
                                                             int attn, float* __restrict__ out,
                                                             float* __restrict__ scores) {
  constexpr int E = KS * 8;
  extern __shared__ __align__(16) float smem[];
  const int pairs = fields * (fields - 1) / 2;
  const int epitch = E + 1;
  float* xs = smem;                          // (N, E+1)
  float* sc = xs + fields * epitch;          // (P) scores
  float* red = sc + ((pairs + 3) & ~3);      // (256) reduction scratch
  int* ptab = reinterpret_cast<int*>(red + kWarps * 32);   // (P) i << 16 | j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  // W1 B-fragments in registers: b0 = W1[a = 8nt + g][k = 8ks + t], b1 = W1[8nt + g][8ks + t + 4]
  uint32_t wh[KS][NT][2], wl[KS][NT][2];
  float bb[NT][2], ww[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int arow = 8 * nt + g;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const float v = arow < attn ? __ldg(w1 + arow * E + 8 * ks + t + 4 * q) : 0.f;
        wh[ks][nt][q] = tf32_rna(v);
        wl[ks][nt][q] = tf32_rna(v - __uint_as_float(wh[ks][nt][q]));
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // accumulator columns of this lane: a = 8nt + 2t + q
      const int a = 8 * nt + 2 * t + q;
      bb[nt][q] = a < attn ? __ldg(b1 + a) : 0.f;
      ww[nt][q] = a < attn ? __ldg(w2 + a) : 0.f;
    }
  }
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  const float bias2 = __ldg(b2);
  const int tiles = (pairs + 15) / 16;

  for (int64_t b = blockIdx.x; b < batch; b += gridDim.x) {
    __syncthreads();
    const float* src = x + b * fields * E;
    for (int i = threadIdx.x; i < fields * E; i += blockDim.x) {
      const int n = i / E, e = i - n * E;
      xs[n * epitch + e] = ldg_stream_f1(src + i);
    }
    __syncthreads();
    // ---- scores: 16 pairs per warp step on the tensor pipe ---------------------------------------------------------
    float lmax = -INFINITY;
    for (int tile = warp; tile < tiles; tile += kWarps) {
      const int pa = tile * 16 + g, pb = pa + 8;
      const int ija = ptab[pa < pairs ? pa : pairs - 1], ijb = ptab[pb < pairs ? pb : pairs - 1];
      const float* xia = xs + (ija >> 16) * epitch;
      const float* xja = xs + (ija & 0xffff) * epitch;
      const float* xib = xs + (ijb >> 16) * epitch;
      const float* xjb = xs + (ijb & 0xffff) * epitch;
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k0 = 8 * ks + t, k1 = k0 + 4;
        uint32_t ah[4], al[4];
        split_fast(xia[k0] * xja[k0], ah[0], al[0]);   // row g,   col t
        split_fast(xib[k0] * xjb[k0], ah[1], al[1]);   // row g+8, col t
        split_fast(xia[k1] * xja[k1], ah[2], al[2]);   // row g,   col t+4
        split_fast(xib[k1] * xjb[k1], ah[3], al[3]);   // row g+8, col t+4
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          mma_tf32(acc[nt], al, wh[ks][nt][0], wh[ks][nt][1]);
          mma_tf32(acc[nt], ah, wl[ks][nt][0], wl[ks][nt][1]);
          mma_tf32(acc[nt], ah, wh[ks][nt][0], wh[ks][nt][1]);
        }
      }
      float sa = 0.f, sb = 0.f;   // rows g and g+8
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sa = fmaf(fmaxf(acc[nt][0] + bb[nt][0], 0.f), ww[nt][0], sa);
        sa = fmaf(fmaxf(acc[nt][1] + bb[nt][1], 0.f), ww[nt][1], sa);
        sb = fmaf(fmaxf(acc[nt][2] + bb[nt][0], 0.f), ww[nt][0], sb);
        sb = fmaf(fmaxf(acc[nt][3] + bb[nt][1], 0.f), ww[nt][1], sb);
      }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      sa += bias2;
      sb += bias2;
      if (t == 0) {
        if (pa < pairs) { sc[pa] = sa; lmax = fmaxf(lmax, sa); }
        if (pb < pairs) { sc[pb] = sb; lmax = fmaxf(lmax, sb); }
      }
    }
    // ---- block softmax over the P scores -----------------------------------------------------------------------------
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float gmax = red[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) gmax = fmaxf(gmax, red[w]);
    __syncthreads();
    float lsum = 0.f;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      const float ev = expf(sc[p] - gmax);
      sc[p] = ev;
      lsum += ev;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float gsum = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) gsum += red[w];
    const float inv = 1.0f / gsum;
    __syncthreads();
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      const float sv = sc[p] * inv;
      sc[p] = sv;
      scores[b * pairs + p] = sv;
    }
    __syncthreads();
    // ---- out[e] = sum_p s_p x_i[e] x_j[e]: thread = (pair slice, e), slices reduced through shared memory ---------------
    constexpr int kSlices = kWarps * 32 / E;
    const int e = threadIdx.x % E, slice = threadIdx.x / E;
    float o = 0.f;
    if (slice < kSlices) {
      for (int p = slice; p < pairs; p += kSlices) {
        const int ij = ptab[p];
        o = fmaf(sc[p], xs[(ij >> 16) * epitch + e] * xs[(ij & 0xffff) * epitch + e], o);
      }
    }
    red[threadIdx.x] = slice < kSlices ? o : 0.f;
    __syncthreads();
    if (threadIdx.x < E) {
      float tot = 0.f;
#pragma unroll 4
      for (int s = 0; s < kSlices; ++s) tot += red[s * E + threadIdx.x];
      out[b * E + threadIdx.x] = tot;
    }
  }
}

template <int KS, int NT>
int afm_tc_dispatch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch,
                    int fields, int attn, float* out, float* scores, cudaStream_t s) {
  constexpr int E = KS * 8;
  const int pairs = fields * (fields - 1) / 2;
  const size_t smem = ((size_t)fields * (E + 1) + ((pairs + 3) & ~3) + kWarps * 32 + pairs) * sizeof(float);
  if (smem > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    TRS_CUDA(cudaFuncSetAttribute(afm_tc_kernel<KS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = static_cast<int>(batch < (int64_t)kNumSMs * 8 ? batch : (int64_t)kNumSMs * 8);
  afm_tc_kernel<KS, NT><<<grid, kWarps * 32, smem, s>>>(x, w1, b1, w2, b2, batch, fields, attn, out, scores);
  return check_launch("afm_tc_kernel");
}

}  // namespace

// returns TRS_ERR_UNSUPPORTED when the shape is not covered (caller falls back to the FFMA kernel)
int afm_tc_launch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch,
                  int fields, int embed, int attn, float* out, float* scores, cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || fields < 2 || attn < 1 || attn > 32) return TRS_ERR_UNSUPPORTED;
  const int nt = (attn + 7) / 8;
#define AFM_CASE(KS, NT) \
  if (embed == 8 * KS && nt == NT) return afm_tc_dispatch<KS, NT>(x, w1, b1, w2, b2, batch, fields, attn, out, scores, s);
  AFM_CASE(1, 1) AFM_CASE(1, 2) AFM_CASE(1, 4)
  AFM_CASE(2, 1) AFM_CASE(2, 2) AFM_CASE(2, 4)
  AFM_CASE(4, 1) AFM_CASE(4, 2) AFM_CASE(4, 4)
  AFM_CASE(8, 1) AFM_CASE(8, 2)
#undef AFM_CASE
  return TRS_ERR_UNSUPPORTED;
}

}  // namespace trs
