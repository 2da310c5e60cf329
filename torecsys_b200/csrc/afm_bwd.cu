// Backward of AttentionalFactorizationMachineLayer (SURVEY.md 8f-2;
// torecsys/layers/ctr/attentional_factorization_machine.py:86-120, eval mode: both dropouts are identities).
// Forward, pairs p = (i<j) lexicographic:
//     prod_p = x_i * x_j          h_p = relu(W1 prod_p + b1)         s_p = <w2, h_p> + b2
//     a = softmax_p(s)            out = sum_p a_p prod_p             (the layer returns out and a)
// With go = grad_out (B,E), gs = grad_scores (B,P) (optional) and the attention a saved by the forward:
//     da_p = <go, prod_p> + gs_p          c = sum_q a_q da_q           ds_p = a_p (da_p - c)
//     dh_p = w2 * ds_p * [h_p > 0]        dprod_p = a_p go + W1^T dh_p
//     grad_x[i] += dprod_p * x_j          grad_x[j] += dprod_p * x_i
//     grad_W1 += dh_p (x) prod_p          grad_b1 += dh_p              grad_w2 += ds_p h_p         grad_b2 += ds_p
//
// One kernel, FP32 FFMA.  CTA = 16 samples x 8 lanes; a sample's fields and its grad_x accumulators live in shared
// memory.  Lane `og` of a sample owns the attention units [og*A/8, (og+1)*A/8) -- their rows of W1 stay in registers for
// the whole kernel, and so do the lane's accumulators of grad_W1 / grad_b1 / grad_w2 over ALL its pairs and samples -- and
// the embedding columns [og*E/8, (og+1)*E/8) of grad_x.  Per pair the lane builds prod_p, its units' h and dh, its
// partial W1^T dh over all E columns, and an 8-lane butterfly reduce-scatter (E*7/8 shuffles) hands every lane the full
// dprod_p of its own columns: every shared-memory accumulator has exactly one owner thread, so the pair loop has no
// atomics and no CTA barrier.  Two passes over the pairs of a tile (c first), parameter gradients reduced over the CTA
// in shared memory at the end and added to global memory once per CTA.
#include "common.cuh"

namespace trs {
namespace {

constexpr int kAbSamples = 16;   // samples per CTA tile (8 lanes each -> 128 threads)

template <int E>
__device__ __forceinline__ void load_field(const float* row, float (&v)[E]) {
#pragma unroll
  for (int k = 0; k < E; k += 4) {   // 16-byte broadcasts: the 8 lanes of a sample read the same address
    const float4 q = *reinterpret_cast<const float4*>(row + k);
    v[k] = q.x, v[k + 1] = q.y, v[k + 2] = q.z, v[k + 3] = q.w;
  }
}

// part[e]: this lane's partial sums for all E columns.  Returns in full[0 .. E/8) the sums over the sample's 8 lanes for
// the lane's own columns og*E/8 + idx (butterfly reduce-scatter over lane bits 4, 2, 1).
template <int E>
__device__ __forceinline__ void lane8_reduce_scatter(const float (&part)[E], int og, float (&full)[E / 8]) {
  constexpr int H1 = E / 2, H2 = E / 4, H3 = E / 8;
  float h1[H1], h2[H2];
  const bool up4 = (og & 4) != 0, up2 = (og & 2) != 0, up1 = (og & 1) != 0;
#pragma unroll
  for (int q = 0; q < H1; ++q) {
    const float send = up4 ? part[q] : part[q + H1];
    const float keep = up4 ? part[q + H1] : part[q];
    h1[q] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int q = 0; q < H2; ++q) {
    const float send = up2 ? h1[q] : h1[q + H2];
    const float keep = up2 ? h1[q + H2] : h1[q];
    h2[q] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
#pragma unroll
  for (int q = 0; q < H3; ++q) {
    const float send = up1 ? h2[q] : h2[q + H3];
    const float keep = up1 ? h2[q + H3] : h2[q];
    full[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
}

template <int E, int A>
__global__ void __launch_bounds__(kAbSamples * 8) afm_backward_kernel(
    const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ w2, const float* __restrict__ scores, const float* __restrict__ grad_out,
    const float* __restrict__ grad_scores, int64_t batch, int fields, float* __restrict__ grad_x,
    float* __restrict__ grad_w1, float* __restrict__ grad_b1, float* __restrict__ grad_w2,
    float* __restrict__ grad_b2) {
  constexpr int V = E / 8, VA = A / 8, Q = VA * E + 2 * VA + 1;   // Q: parameter-gradient words per thread
  extern __shared__ __align__(16) float ab_smem[];
  const int tile = fields * E;
  float* xs = ab_smem;                       // [16][tile]
  float* ds = ab_smem + kAbSamples * tile;   // [16][tile]
  const int tid = threadIdx.x, s = tid >> 3, og = tid & 7;
  const int pairs = fields * (fields - 1) / 2;
  const int64_t tiles = (batch + kAbSamples - 1) / kAbSamples;
  float* xrow = xs + s * tile;
  float* drow = ds + s * tile;

  float w1r[VA][E], b1r[VA], w2r[VA];
  float gw1[VA][E], gb1[VA], gw2[VA], gb2 = 0.f;
#pragma unroll
  for (int r = 0; r < VA; ++r) {
    b1r[r] = __ldg(b1 + og * VA + r);
    w2r[r] = __ldg(w2 + og * VA + r);
    gb1[r] = gw2[r] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      w1r[r][e] = __ldg(w1 + (og * VA + r) * E + e);
      gw1[r][e] = 0.f;
    }
  }

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t b0 = t * kAbSamples;
    const int64_t live = (batch - b0 < kAbSamples ? batch - b0 : kAbSamples) * tile;
    for (int c = tid * 4; c < kAbSamples * tile; c += kAbSamples * 8 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < live) v = ldg_stream_f4(reinterpret_cast<const float4*>(x + b0 * tile + c));
      *reinterpret_cast<float4*>(xs + c) = v;
      *reinterpret_cast<float4*>(ds + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const bool valid = b0 + s < batch;
    const float* arow = scores + (b0 + s) * pairs;                                  // dereferenced only when valid
    const float* gsrow = grad_scores != nullptr ? grad_scores + (b0 + s) * pairs : nullptr;
    float go[E], go_own[V];
#pragma unroll
    for (int e = 0; e < E; ++e) go[e] = valid ? __ldg(grad_out + (b0 + s) * E + e) : 0.f;
#pragma unroll
    for (int c = 0; c < V; ++c) go_own[c] = valid ? __ldg(grad_out + (b0 + s) * E + og * V + c) : 0.f;

    // pass 1: c = sum_p a_p da_p  (every lane of the sample computes the same value)
    float csum = 0.f;
    {
      int i = opaque_zero(), j = i + 1;   // per-thread counters (see bilinear_bwd.cu)
      for (int p = 0; p < pairs; ++p) {
        float xi[E], xj[E];
        load_field<E>(xrow + i * E, xi);
        load_field<E>(xrow + j * E, xj);
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) dot = fmaf(go[e], xi[e] * xj[e], dot);
        const float a = valid ? __ldg(arow + p) : 0.f;
        const float da = dot + ((valid && gsrow != nullptr) ? __ldg(gsrow + p) : 0.f);
        csum = fmaf(a, da, csum);
        if (++j == fields) {
          ++i;
          j = i + 1;
        }
      }
    }

    // pass 2
    {
      int i = opaque_zero(), j = i + 1;
      float acci[V];
#pragma unroll
      for (int c = 0; c < V; ++c) acci[c] = 0.f;
      for (int p = 0; p < pairs; ++p) {
        float prod[E], part[E];
        {
          float xj[E];
          load_field<E>(xrow + i * E, prod);
          load_field<E>(xrow + j * E, xj);
#pragma unroll
          for (int e = 0; e < E; ++e) prod[e] *= xj[e];
        }
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) dot = fmaf(go[e], prod[e], dot);
        const float a = valid ? __ldg(arow + p) : 0.f;
        const float da = dot + ((valid && gsrow != nullptr) ? __ldg(gsrow + p) : 0.f);
        const float dsv = a * (da - csum);
        if (og == 0) gb2 += dsv;
#pragma unroll
        for (int e = 0; e < E; ++e) part[e] = 0.f;
#pragma unroll
        for (int r = 0; r < VA; ++r) {
          float h = b1r[r];
#pragma unroll
          for (int e = 0; e < E; ++e) h = fmaf(w1r[r][e], prod[e], h);
          const bool on = h > 0.f;
          const float dh = on ? w2r[r] * dsv : 0.f;
          gw2[r] = fmaf(dsv, on ? h : 0.f, gw2[r]);
          gb1[r] += dh;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            gw1[r][e] = fmaf(dh, prod[e], gw1[r][e]);
            part[e] = fmaf(w1r[r][e], dh, part[e]);
          }
        }
        float full[V];
        lane8_reduce_scatter<E>(part, og, full);
#pragma unroll
        for (int c = 0; c < V; ++c) {
          const float dprod = fmaf(a, go_own[c], full[c]);
          const float xi_own = xrow[i * E + og * V + c], xj_own = xrow[j * E + og * V + c];
          drow[j * E + og * V + c] += dprod * xi_own;   // this lane owns these columns of the sample
          acci[c] = fmaf(dprod, xj_own, acci[c]);
        }
        if (++j == fields) {   // the run of pairs (i, *) is over
#pragma unroll
          for (int c = 0; c < V; ++c) {
            drow[i * E + og * V + c] += acci[c];
            acci[c] = 0.f;
          }
          ++i;
          j = i + 1;
        }
      }
    }
    __syncthreads();
    for (int c = tid * 4; c < live; c += kAbSamples * 8 * 4) {
      stg_stream_f4(reinterpret_cast<float4*>(grad_x + b0 * tile + c), *reinterpret_cast<const float4*>(ds + c));
    }
    __syncthreads();   // the next tile (or the reduction below) overwrites xs / ds
  }

  // parameter gradients: red[tid][Q] -> sum over the 16 samples of every lane index -> one atomic per word and CTA
  float* red = ab_smem;
  {
    float* mine = red + tid * Q;
#pragma unroll
    for (int r = 0; r < VA; ++r) {
#pragma unroll
      for (int e = 0; e < E; ++e) mine[r * E + e] = gw1[r][e];
      mine[VA * E + r] = gb1[r];
      mine[VA * E + VA + r] = gw2[r];
    }
    mine[VA * E + 2 * VA] = gb2;
  }
  __syncthreads();
  for (int idx = tid; idx < 8 * Q; idx += kAbSamples * 8) {
    const int lane = idx / Q, q = idx - lane * Q;
    float v = 0.f;
    for (int smp = 0; smp < kAbSamples; ++smp) v += red[(smp * 8 + lane) * Q + q];
    if (q < VA * E) {
      atomicAdd(grad_w1 + (lane * VA + q / E) * E + q % E, v);
    } else if (q < VA * E + VA) {
      atomicAdd(grad_b1 + lane * VA + (q - VA * E), v);
    } else if (q < VA * E + 2 * VA) {
      atomicAdd(grad_w2 + lane * VA + (q - VA * E - VA), v);
    } else if (lane == 0) {
      atomicAdd(grad_b2, v);
    }
  }
}

template <int E, int A>
int afm_backward_run(const float* x, const float* w1, const float* b1, const float* w2, const float* scores,
                     const float* grad_out, const float* grad_scores, int64_t batch, int fields, float* grad_x,
                     float* grad_w1, float* grad_b1, float* grad_w2, float* grad_b2, cudaStream_t s) {
  constexpr int Q = (A / 8) * E + 2 * (A / 8) + 1;
  size_t smem = (size_t)2 * kAbSamples * fields * E * sizeof(float);
  const size_t red = (size_t)kAbSamples * 8 * Q * sizeof(float);
  if (smem < red) smem = red;
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_afm_backward: %d fields x %d do not fit shared memory", fields, E);
  TRS_SMEM_OPT_IN((afm_backward_kernel<E, A>));
  int resident = static_cast<int>((size_t)(220 * 1024) / (smem + 1024));
  if (resident < 1) resident = 1;
  if (resident > 4) resident = 4;
  const int64_t tiles = (batch + kAbSamples - 1) / kAbSamples;
  const int64_t cap = (int64_t)kNumSMs * resident;
  afm_backward_kernel<E, A><<<static_cast<int>(tiles < cap ? tiles : cap), kAbSamples * 8, smem, s>>>(
      x, w1, b1, w2, scores, grad_out, grad_scores, batch, fields, grad_x, grad_w1, grad_b1, grad_w2, grad_b2);
  return check_launch("afm_backward_kernel");
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_afm_backward_supported(int embed, int attn) {
  return (embed == 8 && (attn == 8 || attn == 16 || attn == 32)) || (embed == 16 && (attn == 8 || attn == 16)) ||
         (embed == 32 && attn == 8);
}

extern "C" int trs_afm_backward(const float* x, const float* w1, const float* b1, const float* w2,
                                const float* scores, const float* grad_out, const float* grad_scores, int64_t batch,
                                int fields, int embed, int attn, float* grad_x, float* grad_w1, float* grad_b1,
                                float* grad_w2, float* grad_b2, void* stream) {
  TRS_REQUIRE(x && w1 && b1 && w2 && scores && grad_out && grad_x && grad_w1 && grad_b1 && grad_w2 && grad_b2,
              "trs_afm_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0 && attn > 0, "trs_afm_backward: bad sizes");
  TRS_UNSUPPORTED(!trs_afm_backward_supported(embed, attn),
                  "trs_afm_backward: (embed, attn) must be one of (8, 8|16|32), (16, 8|16), (32, 8); got (%d, %d)",
                  embed, attn);
  TRS_UNSUPPORTED(!aligned16(x) || !aligned16(grad_x), "trs_afm_backward: x and grad_x must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TRS_CUDA(cudaMemsetAsync(grad_w1, 0, (size_t)attn * embed * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_b1, 0, (size_t)attn * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_w2, 0, (size_t)attn * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_b2, 0, sizeof(float), s));
  if (batch == 0) return TRS_OK;
#define TRS_AFM_BWD(E_, A_)                                                                                         \
  if (embed == E_ && attn == A_)                                                                                    \
    return afm_backward_run<E_, A_>(x, w1, b1, w2, scores, grad_out, grad_scores, batch, fields, grad_x, grad_w1,   \
                                    grad_b1, grad_w2, grad_b2, s);
  TRS_AFM_BWD(8, 8)
  TRS_AFM_BWD(8, 16)
  TRS_AFM_BWD(8, 32)
  TRS_AFM_BWD(16, 8)
  TRS_AFM_BWD(16, 16)
  TRS_AFM_BWD(32, 8)
#undef TRS_AFM_BWD
  return TRS_ERR_UNSUPPORTED;
}
