// Backward of BilinearInteractionLayer (torecsys/layers/ctr/bilinear_interaction.py:230-255 with
// FieldAllTypeBilinear :72-76 or FieldEachTypeBilinear :144-149).  Forward, pairs p = (i<j) lexicographic:
//     y[b,p,o]   = sum_k x[b,i,k] * W_(p)[k,o]            out[b,p,o] = y[b,p,o] * x[b,j,o] + bias_(p)[o]
// so with g = grad_out and t[b,p,o] = g[b,p,o] * x[b,j,o]:
//     grad_x[b,i,k] += sum_o t[b,p,o] * W_(p)[k,o]         grad_x[b,j,o] += g[b,p,o] * y[b,p,o]
//     grad_W_(p)[k,o] += sum_b x[b,i,k] * t[b,p,o]         grad_bias_(p)[o] += sum_b g[b,p,o]
// ("all" type: one W / bias shared by every pair, so its gradients also sum over p.)
//
// Two kernels, both FP32 FFMA:
//   bilinear_backward_x_kernel  sample-major.  CTA = 16 samples x 8 lanes; the lane group of a sample keeps the
//       sample's fields and its grad_x accumulators in shared memory, lane `og` owns the output columns
//       [og*E/8, (og+1)*E/8) of every pair.  The pair loop is flat (p = 0..P-1) with a 4-deep register ring of
//       grad_out prefetches; the partial grad_x[i] stays in registers for the whole run of pairs (i, *) and is
//       reduced over the 8 lanes once per field.  No CTA barrier and no atomics inside the loop: every shared-memory
//       accumulator has exactly one owner thread.  W_(p) is read through L1 (1 KB per pair at E = 16, shared by the CTA).
//   bilinear_backward_w_kernel  pair-major.  CTA = one pair (x one slice of the batch); chunks of 64 samples of x_i,
//       x_j and grad_out[.,p,.] are staged in shared memory, thread (block, group) accumulates a 4x4 block of the E x E
//       outer-product sum over its samples, groups are reduced through shared memory and the CTA adds its result to
//       global memory once (float atomics: order-dependent in the last bits when several CTAs share an output).
#include "common.cuh"

namespace trs {
namespace {

constexpr int kBxSamples = 16;   // samples per CTA tile of the x kernel (8 lanes each -> 128 threads)
constexpr int kBxDepth = 4;      // grad_out prefetch ring, in pairs
constexpr int kBwChunk = 64;     // samples staged per step of the w kernel

template <int V>
struct VecLoad;
template <>
struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void ld_stream(const float* p, float* v) { v[0] = ldg_stream_f1(p); }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float2 r = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = r.x, v[1] = r.y;
  }
  static __device__ __forceinline__ void ld_stream(const float* p, float* v) {
    const float2 r = ldg_stream_f2(reinterpret_cast<const float2*>(p));
    v[0] = r.x, v[1] = r.y;
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float* v) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
  }
  static __device__ __forceinline__ void ld_stream(const float* p, float* v) {
    const float4 r = ldg_stream_f4(reinterpret_cast<const float4*>(p));
    v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
  }
};

template <int E>
__global__ void __launch_bounds__(kBxSamples * 8) bilinear_backward_x_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ grad_out, int each_type,
    int64_t batch, int fields, float* __restrict__ grad_x) {
  constexpr int V = E / 8;   // columns owned by a lane
  extern __shared__ __align__(16) float bx_smem[];
  const int tile = fields * E;                       // floats per sample
  float* xs = bx_smem;                               // [16][tile]  the samples' fields
  float* ds = bx_smem + kBxSamples * tile;           // [16][tile]  grad_x accumulators
  const int tid = threadIdx.x, s = tid >> 3, og = tid & 7;
  const int pairs = fields * (fields - 1) / 2;
  const int64_t tiles = (batch + kBxSamples - 1) / kBxSamples;
  float* xrow = xs + s * tile;
  float* drow = ds + s * tile;

  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int64_t b0 = t * kBxSamples;
    const int64_t live = (batch - b0 < kBxSamples ? batch - b0 : kBxSamples) * tile;   // floats of real samples
    for (int c = tid * 4; c < kBxSamples * tile; c += kBxSamples * 8 * 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < live) v = ldg_stream_f4(reinterpret_cast<const float4*>(x + b0 * tile + c));
      *reinterpret_cast<float4*>(xs + c) = v;
      *reinterpret_cast<float4*>(ds + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const bool valid = b0 + s < batch;
    const float* grow = grad_out + ((b0 + s) * pairs) * E + og * V;   // + p*E per pair (only dereferenced when valid)
    float gq[kBxDepth][V];
#pragma unroll
    for (int d = 0; d < kBxDepth; ++d) {
#pragma unroll
      for (int c = 0; c < V; ++c) gq[d][c] = 0.f;
      if (valid && d < pairs) VecLoad<V>::ld_stream(grow + (int64_t)d * E, gq[d]);
    }
    float acc[E];
#pragma unroll
    for (int k = 0; k < E; ++k) acc[k] = 0.f;
    // (i, j) live in per-thread registers: with CTA-uniform counters ptxas 12.9 kept them in uniform registers and
    // re-read x_i from field i + 2 after the first run of pairs (seen in the SASS and in the results)
    int i = opaque_zero(), j = i + 1;
    for (int p0 = 0; p0 < pairs; p0 += kBxDepth) {
#pragma unroll
      for (int d = 0; d < kBxDepth; ++d) {
        const int p = p0 + d;
        if (p < pairs) {   // uniform over the CTA
          float g[V], xj[V], tt[V], y[V], xi[E];
#pragma unroll
          for (int k = 0; k < E; k += 4) {   // x_i: a 16-byte broadcast per four components
            const float4 v = *reinterpret_cast<const float4*>(xrow + i * E + k);
            xi[k] = v.x, xi[k + 1] = v.y, xi[k + 2] = v.z, xi[k + 3] = v.w;
          }
#pragma unroll
          for (int c = 0; c < V; ++c) {
            g[c] = gq[d][c];
            gq[d][c] = 0.f;
          }
          if (valid && p + kBxDepth < pairs) VecLoad<V>::ld_stream(grow + (int64_t)(p + kBxDepth) * E, gq[d]);
#pragma unroll
          for (int c = 0; c < V; ++c) {
            xj[c] = xrow[j * E + og * V + c];
            tt[c] = g[c] * xj[c];
            y[c] = 0.f;
          }
          const float* wp = w + (each_type ? (int64_t)p * E * E : 0) + og * V;
#pragma unroll
          for (int k = 0; k < E; ++k) {
            float wv[V];
            VecLoad<V>::ld(wp + k * E, wv);
#pragma unroll
            for (int c = 0; c < V; ++c) {
              y[c] = fmaf(xi[k], wv[c], y[c]);
              acc[k] = fmaf(tt[c], wv[c], acc[k]);
            }
          }
#pragma unroll
          for (int c = 0; c < V; ++c) drow[j * E + og * V + c] += g[c] * y[c];   // this lane owns these columns
          if (++j == fields) {   // the run of pairs (i, *) is over: reduce grad_x[i] over the sample's 8 lanes
#pragma unroll
            for (int k = 0; k < E; ++k) {
              float v = acc[k];
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              v += __shfl_xor_sync(0xffffffffu, v, 2);
              v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (k / V == og) drow[i * E + k] += v;
              acc[k] = 0.f;
            }
            ++i;
            j = i + 1;
          }
        }
      }
    }
    __syncthreads();
    for (int c = tid * 4; c < live; c += kBxSamples * 8 * 4) {
      stg_stream_f4(reinterpret_cast<float4*>(grad_x + b0 * tile + c), *reinterpret_cast<const float4*>(ds + c));
    }
    __syncthreads();   // the next tile overwrites xs / ds
  }
}

// grid (pairs, slices).  Thread = (blk = tid % BLK, grp = tid / BLK): block (kb, ob) = 4 rows k x 4 columns o of the
// E x E sum, over the samples grp, grp + GRP, ... of every staged chunk.
template <int E>
__global__ void __launch_bounds__(256) bilinear_backward_w_kernel(const float* __restrict__ x,
                                                                  const float* __restrict__ grad_out, int each_type,
                                                                  int64_t batch, int fields, int64_t per_slice,
                                                                  float* __restrict__ grad_w,
                                                                  float* __restrict__ grad_bias) {
  constexpr int Q = E / 4, BLK = Q * Q, GRP = 256 / BLK;
  static_assert(kBwChunk % GRP == 0 || GRP > kBwChunk, "chunk must split over the groups");
  __shared__ __align__(16) float xis[kBwChunk * E];
  __shared__ __align__(16) float xjs[kBwChunk * E];
  __shared__ __align__(16) float gs[kBwChunk * E];
  __shared__ __align__(16) float red[256 * 16];   // [grp][E*E] partial sums, then [grp][E] for the bias
  const int tid = threadIdx.x, blk = tid % BLK, grp = tid / BLK, kb = blk / Q, ob = blk % Q;
  const int p = blockIdx.x;
  const int pairs = fields * (fields - 1) / 2;
  int i, j;
  pair_from_index(p, fields, i, j);
  const int64_t b_begin = (int64_t)blockIdx.y * per_slice;
  const int64_t b_end = b_begin + per_slice < batch ? b_begin + per_slice : batch;
  float acc[4][4], gsum[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    gsum[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  }
  for (int64_t c0 = b_begin; c0 < b_end; c0 += kBwChunk) {
    for (int c = tid; c < kBwChunk * Q; c += 256) {   // one float4 of one sample's row per item
      const int sm = c / Q, q = c - sm * Q;
      const int64_t b = c0 + sm;
      float4 vi = make_float4(0.f, 0.f, 0.f, 0.f), vj = vi, vg = vi;
      if (b < b_end) {
        vi = __ldg(reinterpret_cast<const float4*>(x + (b * fields + i) * E) + q);
        vj = __ldg(reinterpret_cast<const float4*>(x + (b * fields + j) * E) + q);
        vg = ldg_stream_f4(reinterpret_cast<const float4*>(grad_out + (b * pairs + p) * E) + q);
      }
      reinterpret_cast<float4*>(xis)[c] = vi;
      reinterpret_cast<float4*>(xjs)[c] = vj;
      reinterpret_cast<float4*>(gs)[c] = vg;
    }
    __syncthreads();
    for (int sm = grp; sm < kBwChunk; sm += GRP) {
      const float4 a4 = *reinterpret_cast<const float4*>(xis + sm * E + kb * 4);
      const float4 j4 = *reinterpret_cast<const float4*>(xjs + sm * E + ob * 4);
      const float4 g4 = *reinterpret_cast<const float4*>(gs + sm * E + ob * 4);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      const float tt[4] = {g4.x * j4.x, g4.y * j4.y, g4.z * j4.z, g4.w * j4.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        gsum[r] += g[r];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], tt[c], acc[r][c]);
      }
    }
    __syncthreads();
  }
  // groups -> one E x E matrix: red[grp][k*E + o]
#pragma unroll
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int c = 0; c < 4; ++c) red[grp * (E * E) + (kb * 4 + r) * E + ob * 4 + c] = acc[r][c];
  }
  __syncthreads();
  float* gw = grad_w + (each_type ? (int64_t)p * E * E : 0);
  for (int el = tid; el < E * E; el += 256) {
    float v = 0.f;
    for (int gi = 0; gi < GRP; ++gi) v += red[gi * (E * E) + el];
    atomicAdd(gw + el, v);
  }
  if (grad_bias != nullptr) {
    __syncthreads();
    if (kb == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) red[grp * E + ob * 4 + r] = gsum[r];
    }
    __syncthreads();
    if (tid < E) {
      float v = 0.f;
      for (int gi = 0; gi < GRP; ++gi) v += red[gi * E + tid];
      atomicAdd(grad_bias + (each_type ? (int64_t)p * E : 0) + tid, v);
    }
  }
}

template <int E>
int bilinear_backward_run(const float* x, const float* w, const float* grad_out, int each_type, int64_t batch,
                          int fields, float* grad_x, float* grad_w, float* grad_bias, cudaStream_t s) {
  const size_t smem = (size_t)2 * kBxSamples * fields * E * sizeof(float);
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_bilinear_backward: %d fields x %d do not fit shared memory", fields, E);
  TRS_SMEM_OPT_IN(bilinear_backward_x_kernel<E>);
  int resident = static_cast<int>((size_t)(220 * 1024) / (smem + 1024));
  if (resident < 1) resident = 1;
  if (resident > 8) resident = 8;
  const int64_t tiles = (batch + kBxSamples - 1) / kBxSamples;
  const int64_t cap = (int64_t)kNumSMs * resident;
  bilinear_backward_x_kernel<E><<<static_cast<int>(tiles < cap ? tiles : cap), kBxSamples * 8, smem, s>>>(
      x, w, grad_out, each_type, batch, fields, grad_x);
  int rc = check_launch("bilinear_backward_x_kernel");
  if (rc != TRS_OK) return rc;
  // slices of the batch per pair: enough CTAs for four waves of 148 SMs, at least one chunk each
  const int pairs = fields * (fields - 1) / 2;
  int64_t slices = (4 * kNumSMs + pairs - 1) / pairs;
  const int64_t chunks = (batch + kBwChunk - 1) / kBwChunk;
  if (slices > chunks) slices = chunks;
  if (slices > 65535) slices = 65535;
  if (slices < 1) slices = 1;
  int64_t per_slice = ((chunks + slices - 1) / slices) * kBwChunk;
  slices = (batch + per_slice - 1) / per_slice;
  bilinear_backward_w_kernel<E><<<dim3(pairs, static_cast<unsigned>(slices)), 256, 0, s>>>(
      x, grad_out, each_type, batch, fields, per_slice, grad_w, grad_bias);
  return check_launch("bilinear_backward_w_kernel");
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_bilinear_backward(const float* x, const float* weight, const float* grad_out, int each_type,
                                     int64_t batch, int fields, int embed, float* grad_x, float* grad_weight,
                                     float* grad_bias, void* stream) {
  TRS_REQUIRE(x && weight && grad_out && grad_x && grad_weight, "trs_bilinear_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_bilinear_backward: bad sizes");
  TRS_UNSUPPORTED(embed != 8 && embed != 16 && embed != 32,
                  "trs_bilinear_backward: embed must be 8, 16 or 32 (got %d)", embed);
  TRS_UNSUPPORTED(!aligned16(x) || !aligned16(weight) || !aligned16(grad_out) || !aligned16(grad_x),
                  "trs_bilinear_backward: x, weight, grad_out and grad_x must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t mats = each_type ? (int64_t)fields * (fields - 1) / 2 : 1;
  TRS_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)mats * embed * embed * sizeof(float), s));
  if (grad_bias != nullptr) TRS_CUDA(cudaMemsetAsync(grad_bias, 0, (size_t)mats * embed * sizeof(float), s));
  if (batch == 0) return TRS_OK;
  switch (embed) {
    case 8: return bilinear_backward_run<8>(x, weight, grad_out, each_type, batch, fields, grad_x, grad_weight, grad_bias, s);
    case 16: return bilinear_backward_run<16>(x, weight, grad_out, each_type, batch, fields, grad_x, grad_weight, grad_bias, s);
    default: return bilinear_backward_run<32>(x, weight, grad_out, each_type, batch, fields, grad_x, grad_weight, grad_bias, s);
  }
}
