// The two interaction layers that sit next to the a1..a12 hot path: the outer-product network of
// PNN and the squeeze-and-excitation (SENET / compose-excitation) layer of FiBiNET and FAT-DeepFFM.
//
//   trs_opn_forward    OuterProductNetworkLayer.forward (torecsys/layers/ctr/outer_product_network.py:80-131)
//   trs_senet_forward  ComposeExcitationNetworkLayer.forward (torecsys/layers/ctr/compose_excitation_network.py:72-109)
#include "common.cuh"

namespace trs {
int opn_mat_tc_launch(const float* x, const float* kernel, int64_t batch, int fields, int embed, float* out,
                      cudaStream_t s);   // bilinear_tc.cu
namespace {

constexpr int kMatTile = 256;   // samples per CTA of the 'mat' kernel (one thread per sample)

// ---- OPN 'mat': out[b,p] = sum_h x_j[h] * (sum_e kernel[h,p,e] * x_i[e]) ------------------------------------------------
// One CTA per (pair, 256-sample tile): K_p (E x E, gathered from the (E, P, E) kernel) is staged once in shared memory
// and read as warp-wide broadcasts; a thread owns one sample and keeps x_i in registers (E known at compile time).
template <int E>
__global__ void __launch_bounds__(kMatTile) opn_mat_kernel(const float* __restrict__ x, const float* __restrict__ kernel,
                                                           int64_t batch, int fields, float* __restrict__ out) {
  __shared__ __align__(16) float ks[E * E];   // ks[h * E + e]
  const int pairs = fields * (fields - 1) / 2;
  const int p = blockIdx.x;
  int i, j;
  pair_from_index(p, fields, i, j);
  for (int t = threadIdx.x; t < E * E; t += blockDim.x) {
    const int h = t / E, e = t - h * E;
    ks[t] = __ldg(kernel + ((int64_t)h * pairs + p) * E + e);
  }
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.y * kMatTile + threadIdx.x;
  if (b >= batch) return;
  const float4* xi4 = reinterpret_cast<const float4*>(x + (b * fields + i) * E);
  const float4* xj4 = reinterpret_cast<const float4*>(x + (b * fields + j) * E);
  float xi[E];
#pragma unroll
  for (int c = 0; c < E / 4; ++c) {
    const float4 v = __ldg(xi4 + c);
    xi[4 * c] = v.x; xi[4 * c + 1] = v.y; xi[4 * c + 2] = v.z; xi[4 * c + 3] = v.w;
  }
  float o = 0.f;
#pragma unroll 2
  for (int h4 = 0; h4 < E / 4; ++h4) {
    const float4 q = __ldg(xj4 + h4);
    const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      const float4* kr = reinterpret_cast<const float4*>(ks + (4 * h4 + hh) * E);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < E / 4; ++c) {
        const float4 k4 = kr[c];
        acc = fmaf(k4.x, xi[4 * c], acc);
        acc = fmaf(k4.y, xi[4 * c + 1], acc);
        acc = fmaf(k4.z, xi[4 * c + 2], acc);
        acc = fmaf(k4.w, xi[4 * c + 3], acc);
      }
      o = fmaf(acc, qv[hh], o);
    }
  }
  out[b * pairs + p] = o;
}

// any embed size: x_i of the thread's sample staged in shared memory (odd pitch), K_p broadcast from shared memory
__global__ void __launch_bounds__(kMatTile) opn_mat_generic_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ kernel, int64_t batch,
                                                                   int fields, int embed, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int pitch = embed | 1;
  float* ks = smem;                        // (E, E)
  float* xs = ks + embed * embed;          // (kMatTile, pitch)
  const int pairs = fields * (fields - 1) / 2;
  const int p = blockIdx.x;
  int i, j;
  pair_from_index(p, fields, i, j);
  for (int t = threadIdx.x; t < embed * embed; t += blockDim.x) {
    const int h = t / embed, e = t - h * embed;
    ks[t] = __ldg(kernel + ((int64_t)h * pairs + p) * embed + e);
  }
  const int64_t b0 = (int64_t)blockIdx.y * kMatTile;
  const int nb = static_cast<int>(batch - b0 < kMatTile ? batch - b0 : kMatTile);
  for (int t = threadIdx.x; t < nb * embed; t += blockDim.x) {   // coalesced over e
    const int s = t / embed, e = t - s * embed;
    xs[s * pitch + e] = __ldg(x + ((b0 + s) * fields + i) * embed + e);
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= nb) return;
  const float* mine = xs + threadIdx.x * pitch;
  const float* xj = x + ((b0 + threadIdx.x) * fields + j) * embed;
  float o = 0.f;
  for (int h = 0; h < embed; ++h) {
    float acc = 0.f;
#pragma unroll 4
    for (int e = 0; e < embed; ++e) acc = fmaf(ks[h * embed + e], mine[e], acc);
    o = fmaf(acc, __ldg(xj + h), o);
  }
  out[(b0 + threadIdx.x) * pairs + p] = o;
}

// ---- OPN 'vec' / 'num': out[b,p] = sum_e x_i[e] x_j[e] k[p,e]   /   k[p] * sum_e x_i[e] x_j[e] ---------------------------
// One warp per sample, the (N, E) tile staged transposed (xt[e][n], odd pitch), lane <-> pair as in ipn_kernel; the
// reference multiplies (p * q) first and the kernel second, and so does this kernel.
template <bool kVec>
__global__ void __launch_bounds__(256) opn_vec_kernel(const float* __restrict__ x, const float* __restrict__ kernel,
                                                      int64_t batch, int fields, int embed, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = fields | 1;
  float* xt = smem + (size_t)warp * embed * pitch;
  int* ptab = reinterpret_cast<int*>(smem + (size_t)warps * embed * pitch);
  const int pairs = fields * (fields - 1) / 2;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  __syncthreads();
  const int tile = fields * embed;
  for (int64_t b = (int64_t)blockIdx.x * warps + warp; b < batch; b += (int64_t)gridDim.x * warps) {
    const float* src = x + b * tile;
    for (int t = lane; t < tile; t += 32) {
      const int n = t / embed, e = t - n * embed;
      xt[e * pitch + n] = ldg_stream_f1(src + t);
    }
    __syncwarp();
    float* dst = out + b * pairs;
    for (int p = lane; p < pairs; p += 32) {
      const int ij = ptab[p];
      const int i = ij >> 16, j = ij & 0xffff;
      float acc = 0.f;
      if (kVec) {
        const float* kp = kernel + (int64_t)p * embed;
#pragma unroll 4
        for (int e = 0; e < embed; ++e) acc = fmaf(xt[e * pitch + i] * xt[e * pitch + j], __ldg(kp + e), acc);
      } else {
        const float k = __ldg(kernel + p);
#pragma unroll 4
        for (int e = 0; e < embed; ++e) acc = fmaf(xt[e * pitch + i] * xt[e * pitch + j], k, acc);
      }
      dst[p] = acc;
    }
    __syncwarp();
  }
}

// Pair-major form for E in {8, 16, 32} and up to 4 pairs per thread: thread <-> pair for the whole kernel, so k[p, :] (or
// k[p]) and (i, j) stay in registers; a tile of samples is staged row-major with pitch E + 4 (the 8 lanes of a quarter
// warp read 8 consecutive x_j rows from 8 different bank groups; x_i is a broadcast) and every thread walks the tile's
// samples: 2 x E/4 LDS.128 per 2E flops instead of three scalar loads per FMA, coalesced 4-byte stores across pairs.
// Same arithmetic order as opn_vec_kernel: fma(x_i[e] * x_j[e], k, acc) for e = 0 .. E-1.
constexpr int kVecThreads = 256;

template <int E, int PT, bool kVec>
__global__ void __launch_bounds__(kVecThreads) opn_vec_pairs_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ kernel, int64_t batch,
                                                                    int fields, int tile_samples,
                                                                    float* __restrict__ out) {
  extern __shared__ __align__(16) float vsm[];
  constexpr int PITCH = E + 4;
  const int pairs = fields * (fields - 1) / 2;
  const int row = fields * PITCH;   // floats per staged sample
  int pi[PT], pj[PT];
  float kr[PT][kVec ? E : 1];
#pragma unroll
  for (int q = 0; q < PT; ++q) {
    const int p = threadIdx.x + q * kVecThreads;
    pi[q] = pj[q] = 0;
#pragma unroll
    for (int e = 0; e < (kVec ? E : 1); ++e) kr[q][e] = 0.f;
    if (p < pairs) {
      pair_from_index(p, fields, pi[q], pj[q]);
      if (kVec) {
#pragma unroll
        for (int e = 0; e < (kVec ? E : 1); ++e) kr[q][e] = __ldg(kernel + (int64_t)p * E + e);
      } else {
        kr[q][0] = __ldg(kernel + p);
      }
    }
  }
  const int64_t tiles = (batch + tile_samples - 1) / tile_samples;
  constexpr int CH = E / 4;   // 16-byte chunks per field row
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t b0 = tile * tile_samples;
    const int valid = static_cast<int>(batch - b0 < tile_samples ? batch - b0 : tile_samples);
    __syncthreads();   // previous tile fully consumed
    for (int c = threadIdx.x; c < valid * fields * CH; c += kVecThreads) {
      const int r = c / CH, k = c - r * CH;   // r = sample * fields + field
      *reinterpret_cast<float4*>(vsm + (size_t)r * PITCH + 4 * k) =
          ldg_stream_f4(reinterpret_cast<const float4*>(x + b0 * fields * E) + c);
    }
    __syncthreads();
    for (int smp = 0; smp < valid; ++smp) {
      const float* xs = vsm + (size_t)smp * row;
#pragma unroll
      for (int q = 0; q < PT; ++q) {
        const int p = threadIdx.x + q * kVecThreads;
        if (p < pairs) {
          const float* xi = xs + pi[q] * PITCH;
          const float* xj = xs + pj[q] * PITCH;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < CH; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(xi + 4 * k);
            const float4 b = *reinterpret_cast<const float4*>(xj + 4 * k);
            acc = fmaf(a.x * b.x, kr[q][kVec ? 4 * k : 0], acc);
            acc = fmaf(a.y * b.y, kr[q][kVec ? 4 * k + 1 : 0], acc);
            acc = fmaf(a.z * b.z, kr[q][kVec ? 4 * k + 2 : 0], acc);
            acc = fmaf(a.w * b.w, kr[q][kVec ? 4 * k + 3 : 0], acc);
          }
          out[(b0 + smp) * pairs + p] = acc;
        }
      }
    }
  }
}

template <int E, int PT, bool kVec>
int launch_opn_vec_pairs(const float* x, const float* kernel, int64_t batch, int fields, float* out, cudaStream_t s) {
  const size_t per_sample = (size_t)fields * (E + 4) * sizeof(float);
  int tile_samples = static_cast<int>((size_t)(100 * 1024) / per_sample);   // two CTAs per SM
  if (tile_samples > 64) tile_samples = 64;
  if (tile_samples < 1) return TRS_ERR_UNSUPPORTED;
  const size_t smem = per_sample * tile_samples;
  TRS_SMEM_OPT_IN((opn_vec_pairs_kernel<E, PT, kVec>));
  const int64_t tiles = (batch + tile_samples - 1) / tile_samples;
  const int64_t cap = (int64_t)kNumSMs * 2;
  opn_vec_pairs_kernel<E, PT, kVec><<<static_cast<int>(tiles < cap ? tiles : cap), kVecThreads, smem, s>>>(
      x, kernel, batch, fields, tile_samples, out);
  return check_launch("opn_vec_pairs_kernel");
}

template <bool kVec>
int opn_vec_pairs_dispatch(const float* x, const float* kernel, int64_t batch, int fields, int embed, float* out,
                           cudaStream_t s) {
  const int pairs = fields * (fields - 1) / 2;
  const int pt = (pairs + kVecThreads - 1) / kVecThreads;
  if (!aligned16(x) || pt > 4 || (kVec && pt * embed > 64)) return TRS_ERR_UNSUPPORTED;
#define TRS_OPN_VEC_CASE(E_, PT_)                \
  if (embed == E_ && pt == PT_) return launch_opn_vec_pairs<E_, PT_, kVec>(x, kernel, batch, fields, out, s);
  TRS_OPN_VEC_CASE(8, 1) TRS_OPN_VEC_CASE(8, 2) TRS_OPN_VEC_CASE(8, 3) TRS_OPN_VEC_CASE(8, 4)
  TRS_OPN_VEC_CASE(16, 1) TRS_OPN_VEC_CASE(16, 2) TRS_OPN_VEC_CASE(16, 3) TRS_OPN_VEC_CASE(16, 4)
  TRS_OPN_VEC_CASE(32, 1) TRS_OPN_VEC_CASE(32, 2)
#undef TRS_OPN_VEC_CASE
  if (!kVec && embed == 32 && (pt == 3 || pt == 4)) {
    if (pt == 3) return launch_opn_vec_pairs<32, 3, false>(x, kernel, batch, fields, out, s);
    return launch_opn_vec_pairs<32, 4, false>(x, kernel, batch, fields, out, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

// ---- SENET -----------------------------------------------------------------------------------------------------------
// pooled[b,m] = mean_e x[b,m,e]  (nn.AdaptiveAvgPool1d(1)); one thread per (b, m) row, 16-byte loads when E % 4 == 0
__global__ void __launch_bounds__(256) senet_pool_kernel(const float* __restrict__ x, int64_t rows, int embed, int vec,
                                                         float* __restrict__ pooled) {
  const float inv = 1.0f / static_cast<float>(embed);
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* src = x + r * embed;
    float acc = 0.f;
    if (vec) {
      for (int c = 0; c < embed; c += 4) {
        const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(src + c));
        acc += v.x; acc += v.y; acc += v.z; acc += v.w;
      }
    } else {
      for (int c = 0; c < embed; ++c) acc += ldg_stream_f1(src + c);
    }
    pooled[r] = acc * inv;
  }
}

// out[b,m,:] = x[b,m,:] * act(a[b,m])   (a = the pre-activation output of AdditionLinear)
__global__ void __launch_bounds__(256) senet_scale_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                          int64_t rows, int embed, int vec, int act,
                                                          float* __restrict__ out) {
  const int64_t total = rows * embed;
  if (vec) {
    const int e4 = embed >> 2;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (total >> 2);
         t += (int64_t)gridDim.x * blockDim.x) {
      const float w = apply_act(__ldg(a + t / e4), act);
      float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(x) + t);
      v.x *= w; v.y *= w; v.z *= w; v.w *= w;
      stg_stream_f4(reinterpret_cast<float4*>(out) + t, v);
    }
  } else {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
      out[t] = ldg_stream_f1(x + t) * apply_act(__ldg(a + t / embed), act);
  }
}

template <int E>
int launch_opn_mat(const float* x, const float* kernel, int64_t batch, int fields, float* out, cudaStream_t s) {
  const int pairs = fields * (fields - 1) / 2;
  const int64_t tiles = (batch + kMatTile - 1) / kMatTile;
  TRS_UNSUPPORTED(tiles > 65535, "trs_opn_forward: batch too large for one launch of the 'mat' kernel");
  opn_mat_kernel<E><<<dim3(pairs, (unsigned)tiles), kMatTile, 0, s>>>(x, kernel, batch, fields, out);
  return check_launch("opn_mat_kernel");
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_opn_forward(const float* x, const float* kernel, int kernel_type, int64_t batch, int fields,
                               int embed, float* out, void* stream) {
  TRS_REQUIRE(x && kernel && out, "trs_opn_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_opn_forward: bad sizes");
  TRS_REQUIRE(kernel_type == TRS_OPN_MAT || kernel_type == TRS_OPN_VEC || kernel_type == TRS_OPN_NUM,
              "trs_opn_forward: kernel_type must be TRS_OPN_MAT, TRS_OPN_VEC or TRS_OPN_NUM");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int pairs = fields * (fields - 1) / 2;
  if (kernel_type == TRS_OPN_MAT) {
    {   // embed 8 / 16 / 32: 3xTF32 mma.sync tiles shared with the bilinear layer (bilinear_tc.cu)
      const int rc = opn_mat_tc_launch(x, kernel, batch, fields, embed, out, s);
      if (rc != TRS_ERR_UNSUPPORTED) return rc;
    }
    if (aligned16(x)) {
      switch (embed) {
        case 4: return launch_opn_mat<4>(x, kernel, batch, fields, out, s);
        case 8: return launch_opn_mat<8>(x, kernel, batch, fields, out, s);
        case 16: return launch_opn_mat<16>(x, kernel, batch, fields, out, s);
        case 32: return launch_opn_mat<32>(x, kernel, batch, fields, out, s);
        default: break;
      }
    }
    const size_t smem = ((size_t)embed * embed + (size_t)kMatTile * (embed | 1)) * sizeof(float);
    TRS_UNSUPPORTED(smem > 200 * 1024, "trs_opn_forward: embed too large for the 'mat' kernel");
    TRS_SMEM_OPT_IN(opn_mat_generic_kernel);
    const int64_t tiles = (batch + kMatTile - 1) / kMatTile;
    TRS_UNSUPPORTED(tiles > 65535, "trs_opn_forward: batch too large for one launch of the 'mat' kernel");
    opn_mat_generic_kernel<<<dim3(pairs, (unsigned)tiles), kMatTile, smem, s>>>(x, kernel, batch, fields, embed, out);
    return check_launch("opn_mat_generic_kernel");
  }
  {   // pair-major register kernel for embed 8 / 16 / 32 and up to 1 024 pairs
    const int rc = kernel_type == TRS_OPN_VEC ? opn_vec_pairs_dispatch<true>(x, kernel, batch, fields, embed, out, s)
                                              : opn_vec_pairs_dispatch<false>(x, kernel, batch, fields, embed, out, s);
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  int warps = 8;
  size_t smem;
  for (;; warps >>= 1) {
    smem = ((size_t)warps * embed * (fields | 1) + pairs) * sizeof(float);
    if (smem <= 200 * 1024 || warps == 1) break;
  }
  TRS_UNSUPPORTED(smem > 200 * 1024, "trs_opn_forward: fields*embed tile does not fit shared memory");
  const int grid = grid_for(batch * 32, warps * 32, 4);
  if (kernel_type == TRS_OPN_VEC) {
    TRS_SMEM_OPT_IN(opn_vec_kernel<true>);
    opn_vec_kernel<true><<<grid, warps * 32, smem, s>>>(x, kernel, batch, fields, embed, out);
  } else {
    TRS_SMEM_OPT_IN(opn_vec_kernel<false>);
    opn_vec_kernel<false><<<grid, warps * 32, smem, s>>>(x, kernel, batch, fields, embed, out);
  }
  return check_launch("opn_vec_kernel");
}

extern "C" int64_t trs_senet_workspace_bytes(int64_t batch, int rows_per_sample) {
  if (batch < 0 || rows_per_sample <= 0) return -1;
  return 2 * batch * rows_per_sample * (int64_t)sizeof(float);
}

extern "C" int trs_senet_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                 int activation, int64_t batch, int rows_per_sample, int embed, int reduced,
                                 float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  TRS_REQUIRE(x && w1 && b1 && w2 && b2 && out, "trs_senet_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && rows_per_sample > 0 && embed > 0 && reduced > 0, "trs_senet_forward: bad sizes");
  TRS_REQUIRE(activation >= TRS_ACT_NONE && activation <= TRS_ACT_TANH, "trs_senet_forward: unknown activation");
  if (batch == 0) return TRS_OK;
  TRS_REQUIRE(workspace && workspace_bytes >= trs_senet_workspace_bytes(batch, rows_per_sample),
              "trs_senet_forward: workspace too small (need %lld bytes)",
              (long long)trs_senet_workspace_bytes(batch, rows_per_sample));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* pooled = static_cast<float*>(workspace);
  float* attn = pooled + batch * rows_per_sample;
  const int64_t rows = batch * rows_per_sample;
  const int vec = (embed & 3) == 0 && aligned16(x) && aligned16(out);
  senet_pool_kernel<<<grid_for(rows, 256, 8), 256, 0, s>>>(x, rows, embed, vec, pooled);
  int rc = check_launch("senet_pool_kernel");
  if (rc != TRS_OK) return rc;
  // the two Linears are an MLP [M -> R -> M] with the activation after the first; the second activation is fused into
  // the scaling kernel (trs_mlp_forward applies no activation to its output layer)
  const int dims[3] = {rows_per_sample, reduced, rows_per_sample};
  const float* ws[2] = {w1, w2};
  const float* bs[2] = {b1, b2};
  rc = trs_mlp_forward(pooled, batch, dims, 2, ws, bs, activation, attn, stream);
  if (rc != TRS_OK) return rc;
  senet_scale_kernel<<<grid_for(vec ? rows * embed / 4 : rows * embed, 256, 8), 256, 0, s>>>(x, attn, rows, embed, vec,
                                                                                              activation, out);
  return check_launch("senet_scale_kernel");
}
