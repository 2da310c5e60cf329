// Block-level building blocks shared by the generic (any N, E, layer sizes) kernels:
//   * gather_tile      : rows of the embedding table for a tile of samples -> shared memory
//   * dense_layer_tile : out[s][o] = epi(sum_k in[s][k] * W[o][k] + b[o]) for a tile of rows resident in
//                        shared memory, W/b read through L1 (they are shared by every CTA and stay cached)
// The E=16 / hidden=16 DeepFM fast path (deepfm_fast.cu) does not use these; it keeps rows in registers.
#pragma once

#include "common.cuh"

namespace trs {

// Row pitch (in floats) for a tile whose rows are read with 128-bit loads by consecutive lanes:
// pitch % 8 == 4 makes the eight 16-byte accesses of a quarter-warp land in distinct bank groups.
__host__ __device__ inline int tile_pitch(int k) {
  int p = (k + 3) & ~3;
  return (p % 8 == 4) ? p : p + 4;
}

struct MlpParams {
  static constexpr int kMaxLayers = 8;
  const float* w[kMaxLayers];
  const float* b[kMaxLayers];
  int dims[kMaxLayers + 1];
  int layers;
  int act;
};

// Fused ends of a tcgen05 dense layer (dense_tc_run, cin_tc.cu) -- the wide deep branch of DeepFM-style models:
//   idx != null   : the layer's input rows are gathered by the kernel itself (row m = the `fields` embedding rows of sample
//                   m, k = field * embed + e), and row_base[m] = sum_f w_feat[r_f] + (use_fm ? FM(rows) : 0) + bias
//   dot_w != null : the layer feeds a one-output Linear; its activations are not stored, dot_out[block][m] receives the
//                   partial products (dense_tc_passes() blocks, closed by dense_dot_finish)
struct DenseFuse {
  const void* idx = nullptr;
  int idx_bits = 64;
  const int64_t* offsets = nullptr;
  const float* table = nullptr;     // (table_rows, embed)
  const float* w_feat = nullptr;    // (table_rows,) or null
  const float* bias = nullptr;      // scalar or null
  int32_t* status = nullptr;
  float* row_base = nullptr;        // (rows,)
  int64_t table_rows = 0;
  int fields = 0, embed = 0, use_fm = 0;
  int row_pitch = 0;                // floats between consecutive table rows (0: embed)
  int w_col = -1;                   // >= 0: the first-order value is column w_col of the row's own line (packed [v|w] table)
  const float* dot_w = nullptr;     // (c_dim,)
  float* dot_out = nullptr;         // (passes, rows)
};

#ifdef __CUDACC__

// Gathers `ts` samples x `fields` rows of `embed` floats into tile[s * pitch + n * row_pitch + e].
// Out-of-range rows are zero-filled and reported; samples beyond `valid` are zero-filled.
template <int IdxBits>
__device__ __forceinline__ void gather_tile(const float* __restrict__ w_emb, int64_t rows, int embed,
                                            const void* __restrict__ idx, const int64_t* __restrict__ offsets,
                                            int64_t b0, int ts, int valid, int fields, float* tile, int pitch,
                                            int row_pitch, int32_t* status) {
  if ((embed & 3) == 0) {
    const int chunks = embed >> 2;
    const int per_sample = fields * chunks;
    const int items = ts * per_sample;
    for (int base = threadIdx.x; base < items; base += blockDim.x * 4) {
      float4 v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = base + u * blockDim.x;
        dst[u] = -1;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < items) {
          const int s = item / per_sample;
          const int rem = item - s * per_sample;
          const int n = rem / chunks;
          const int c = rem - n * chunks;
          const int64_t pos = (b0 + s) * fields + n;
          dst[u] = s * pitch + n * row_pitch + c * 4;
          if (s >= valid) continue;
          const int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + n);
          if (r < 0 || r >= rows) {
            if (c == 0) report_oob(status, pos);
          } else {
            v[u] = ldg_stream_f4(reinterpret_cast<const float4*>(w_emb + r * embed) + c);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) *reinterpret_cast<float4*>(tile + dst[u]) = v[u];
    }
  } else {
    const int per_sample = fields * embed;
    const int items = ts * per_sample;
    for (int item = threadIdx.x; item < items; item += blockDim.x) {
      const int s = item / per_sample;
      const int rem = item - s * per_sample;
      const int n = rem / embed;
      const int e = rem - n * embed;
      const int64_t pos = (b0 + s) * fields + n;
      float v = 0.f;
      if (s < valid) {
        const int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + n);
        if (r < 0 || r >= rows) {
          if (e == 0) report_oob(status, pos);
        } else {
          v = ldg_stream_f1(w_emb + r * embed + e);
        }
      }
      tile[s * pitch + n * row_pitch + e] = v;
    }
  }
}

// out = epi(s, o, sum_k in[s][k] W[o][k] + b[o]).  Thread item = (row s, group of 4 outputs); consecutive lanes
// take consecutive rows so the W loads of a warp are (mostly) uniform and the tile reads are conflict-free.
template <class Epi>
__device__ __forceinline__ void dense_layer_tile(const float* in, int in_pitch, int k_dim,
                                                 const float* __restrict__ w, const float* __restrict__ bias,
                                                 int o_dim, int ts, Epi epi) {
  const int groups = (o_dim + 3) >> 2;
  const int items = ts * groups;
  const bool vec = ((k_dim & 3) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15u) == 0);
  if (items * 2 <= static_cast<int>(blockDim.x) && k_dim >= 64) {
    // few rows x few outputs but a long K (e.g. 16 samples x 16 outputs x K = 781 of the fused PNN): with 4 outputs
    // per thread most of the CTA would idle, so every (row, output) pair gets its own thread
    const int singles = ts * o_dim;
    for (int item = threadIdx.x; item < singles; item += blockDim.x) {
      const int o = item / ts;
      const int s = item - o * ts;
      const float* xr = in + s * in_pitch;
      const float* wr = w + (int64_t)o * k_dim;
      float acc0 = 0.f, acc1 = 0.f;
      if (vec) {
        for (int k = 0; k < k_dim; k += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xr + k);
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
          acc0 = fmaf(xv.x, wv.x, acc0); acc1 = fmaf(xv.y, wv.y, acc1);
          acc0 = fmaf(xv.z, wv.z, acc0); acc1 = fmaf(xv.w, wv.w, acc1);
        }
      } else {
        int k = 0;
        for (; k + 1 < k_dim; k += 2) {
          acc0 = fmaf(xr[k], __ldg(wr + k), acc0);
          acc1 = fmaf(xr[k + 1], __ldg(wr + k + 1), acc1);
        }
        if (k < k_dim) acc0 = fmaf(xr[k], __ldg(wr + k), acc0);
      }
      epi(s, o, acc0 + acc1 + (bias ? __ldg(bias + o) : 0.f));
    }
    return;
  }
  for (int item = threadIdx.x; item < items; item += blockDim.x) {
    const int og = item / ts;
    const int s = item - og * ts;
    const int o0 = og << 2;
    const float* xr = in + s * in_pitch;
    const float* wr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wr[u] = w + (int64_t)min(o0 + u, o_dim - 1) * k_dim;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) {
#pragma unroll 2
      for (int k = 0; k < k_dim; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wr[u] + k));
          acc[u] = fmaf(xv.x, wv.x, acc[u]);
          acc[u] = fmaf(xv.y, wv.y, acc[u]);
          acc[u] = fmaf(xv.z, wv.z, acc[u]);
          acc[u] = fmaf(xv.w, wv.w, acc[u]);
        }
      }
    } else {
      for (int k = 0; k < k_dim; ++k) {
        const float xv = xr[k];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(xv, __ldg(wr[u] + k), acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (o0 + u < o_dim) epi(s, o0 + u, acc[u] + (bias ? __ldg(bias + o0 + u) : 0.f));
  }
}

// Runs an MLP (hidden layers with activation, last layer without) over a tile resident in shared memory.
// `in` (ts x dims[0], pitch in_pitch) is preserved; buf0/buf1 are ping-pong buffers with pitch `hpitch`
// (>= tile_pitch(max hidden/out dim)).  Returns the buffer holding the (ts x dims[layers]) result.
__device__ __forceinline__ float* mlp_tile(const MlpParams& mp, const float* in, int in_pitch, int ts, float* buf0,
                                           float* buf1, int hpitch) {
  const float* cur = in;
  int cur_pitch = in_pitch;
  float* dst = buf0;
  for (int l = 0; l < mp.layers; ++l) {
    const bool last = (l == mp.layers - 1);
    const int act = last ? TRS_ACT_NONE : mp.act;
    float* d = dst;
    dense_layer_tile(cur, cur_pitch, mp.dims[l], mp.w[l], mp.b[l], mp.dims[l + 1], ts,
                     [=](int s, int o, float v) { d[s * hpitch + o] = apply_act(v, act); });
    __syncthreads();
    cur = dst;
    cur_pitch = hpitch;
    dst = (dst == buf0) ? buf1 : buf0;
  }
  return const_cast<float*>(cur);
}

#endif  // __CUDACC__

inline int mlp_max_hidden(const int* dims, int layers) {
  int m = 1;
  for (int l = 1; l <= layers; ++l) m = dims[l] > m ? dims[l] : m;
  return m;
}

inline int fill_mlp_params(MlpParams& mp, const int* dims, int layers, const float* const* w, const float* const* b,
                           int act) {
  if (layers < 0 || layers > MlpParams::kMaxLayers) return -1;
  mp.layers = layers;
  mp.act = act;
  for (int l = 0; l < layers; ++l) {
    mp.w[l] = w[l];
    mp.b[l] = b ? b[l] : nullptr;
    if (!w[l]) return -1;
  }
  for (int l = 0; l <= layers; ++l) {
    mp.dims[l] = dims[l];
    if (dims[l] <= 0) return -1;
  }
  return 0;
}

}  // namespace trs
