// Dedicated BACKWARD kernels of the hot path: the ops the a12 models' training steps go through.
// (The remaining layers back-propagate by recomputing their formula with differentiable torch CUDA ops,
// torecsys_b200/autograd.py.)
//
//   trs_embedding_grad   d weight of MultiIndicesEmbedding / SingleIndexEmbedding (nn.Embedding's dense gradient,
//                        torecsys/inputs/base/multi_indices_emb.py:48, sparse=False): grad_weight[idx + off] += grad_out
//   trs_embedding_rows / trs_embedding_grad_segments   the same gradient as a sparse COO tensor (nn.Embedding(sparse=True)):
//                        row ids in lookup order, and the deterministic coalescing of their sorted form
//   trs_fm_backward      d x of FactorizationMachineLayer (factorization_machine.py:46-81):
//                        grad_x[b,n,e] = grad_out[b,e] * (sum_m x[b,m,e] - x[b,n,e])
//   trs_ffm_backward     d v of FieldAwareFactorizationMachineLayer (field_aware_factorization_machine.py:50-94)
//   trs_ipn_backward     d x of InnerProductNetworkLayer (inner_product_network.py:54-79)
//   trs_cross_backward   d x, d W_l, d b_l of CrossNetworkLayer (cross_network.py:52-87), h_0 detached as upstream (:65)
#include "common.cuh"

namespace trs {
namespace {

// One work item = one 16-byte chunk of one looked-up row: grad_weight[row][chunk] += grad_out[pos][chunk] with ONE
// vector reduction (red.global.add.v4.f32, sm_90+): the rows of a batch are random, collisions are rare and resolved
// by the memory system; the sum over colliding lookups is order-dependent in the last bit, like torch's index_add_.
template <int IdxBits>
__global__ void __launch_bounds__(256) embedding_grad_vec_kernel(const float4* __restrict__ grad_out,
                                                                 const void* __restrict__ idx,
                                                                 const int64_t* __restrict__ offsets, int64_t rows,
                                                                 uint32_t chunks, uint32_t fields, int64_t items,
                                                                 int64_t padding_row, float* __restrict__ grad_weight) {
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = item / chunks;
    const uint32_t chunk = static_cast<uint32_t>(item - pos * chunks);
    int64_t r = load_index<IdxBits>(idx, pos);
    if (offsets != nullptr) r += __ldg(offsets + pos % fields);
    if (r < 0 || r >= rows || r == padding_row) continue;   // out-of-range lookups were reported by the forward
    const float4 g = ldg_stream_f4(grad_out + item);
    float* dst = grad_weight + (r * chunks + chunk) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(g.x), "f"(g.y), "f"(g.z), "f"(g.w)
                 : "memory");
  }
}

template <int IdxBits>
__global__ void __launch_bounds__(256) embedding_grad_scalar_kernel(const float* __restrict__ grad_out,
                                                                    const void* __restrict__ idx,
                                                                    const int64_t* __restrict__ offsets, int64_t rows,
                                                                    uint32_t embed, uint32_t fields, int64_t items,
                                                                    int64_t padding_row,
                                                                    float* __restrict__ grad_weight) {
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = item / embed;
    const uint32_t e = static_cast<uint32_t>(item - pos * embed);
    int64_t r = load_index<IdxBits>(idx, pos);
    if (offsets != nullptr) r += __ldg(offsets + pos % fields);
    if (r < 0 || r >= rows || r == padding_row) continue;
    atomicAdd(grad_weight + r * embed + e, ldg_stream_f1(grad_out + item));
  }
}

// Sparse (COO) form of the same gradient, what nn.Embedding(sparse=True) produces: the row ids idx + offsets of every
// lookup, in lookup order; the values are grad_out itself, viewed (B*N, E).
template <int IdxBits>
__global__ void __launch_bounds__(256) embedding_rows_kernel(const void* __restrict__ idx,
                                                             const int64_t* __restrict__ offsets, uint32_t fields,
                                                             int64_t lookups, int64_t* __restrict__ out_rows) {
  for (int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pos < lookups;
       pos += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = load_index<IdxBits>(idx, pos);
    if (offsets != nullptr) r += __ldg(offsets + pos % fields);
    out_rows[pos] = r;
  }
}

// Coalescing of that sparse gradient after its row ids have been sorted (perm = the sort's permutation, starts[s] = first
// sorted position of the s-th distinct row, starts[segments] = number of lookups).  Work item = (segment, 16-byte chunk):
// the rows of a segment are added in sorted order, so the sums are deterministic (unlike the red.global.add path).
template <int VEC>
__global__ void __launch_bounds__(256) embedding_grad_segments_kernel(const float* __restrict__ grad_out,
                                                                      const int64_t* __restrict__ perm,
                                                                      const int64_t* __restrict__ starts,
                                                                      int64_t segments, uint32_t chunks,
                                                                      float* __restrict__ out_values) {
  const int64_t items = segments * chunks;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t seg = item / chunks;
    const uint32_t chunk = static_cast<uint32_t>(item - seg * chunks);
    const int64_t lo = __ldg(starts + seg), hi = __ldg(starts + seg + 1);
    if (VEC == 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t m = lo; m < hi; ++m) {
        const float4 g = ldg_stream_f4(reinterpret_cast<const float4*>(grad_out) + __ldg(perm + m) * chunks + chunk);
        acc.x += g.x, acc.y += g.y, acc.z += g.z, acc.w += g.w;
      }
      reinterpret_cast<float4*>(out_values)[item] = acc;
    } else {
      float acc = 0.f;
      for (int64_t m = lo; m < hi; ++m) acc += ldg_stream_f1(grad_out + __ldg(perm + m) * chunks + chunk);
      out_values[item] = acc;
    }
  }
}

// FM backward: one warp per sample; pass 1 sums the fields per embedding component, pass 2 writes the gradient.
__global__ void __launch_bounds__(256) fm_backward_kernel(const float* __restrict__ x, const float* __restrict__ grad_out,
                                                          int64_t batch, int fields, int embed,
                                                          float* __restrict__ grad_x) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int tile = fields * embed;
  for (int64_t b = warp0; b < batch; b += warps) {
    const float* xb = x + b * tile;
    float* gb = grad_x + b * tile;
    for (int e0 = 0; e0 < embed; e0 += 32) {   // 32 embedding components per pass (one pass for E <= 32)
      const int e = e0 + lane;
      float s = 0.f;
      if (e < embed)
        for (int n = 0; n < fields; ++n) s += __ldg(xb + n * embed + e);
      if (e < embed) {
        const float g = __ldg(grad_out + b * embed + e);
        for (int n = 0; n < fields; ++n) gb[n * embed + e] = g * (s - __ldg(xb + n * embed + e));
      }
    }
  }
}


// FFM backward: out[b,p] = v[b,i*N+j] * v[b,j*N+i]  =>  grad_v[b, a*N+c] = grad_out[b, p(a,c)] * v[b, c*N+a] for a != c
// and 0 on the diagonal (those rows never reach the output).  Work item = one 16-byte chunk (VEC = 4) or one float of
// one row of grad_v: one streaming read of the partner row and of the output gradient, one streaming write.
template <int VEC>
__global__ void __launch_bounds__(256) ffm_backward_kernel(const float* __restrict__ v, const float* __restrict__ grad_out,
                                                           int64_t batch, int fields, int embed,
                                                           float* __restrict__ grad_v) {
  const int chunks = embed / VEC;
  const int slots = fields * fields;
  const int pairs = fields * (fields - 1) / 2;
  const int64_t items = batch * slots * chunks;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int chunk = static_cast<int>(item % chunks);
    const int64_t row = item / chunks;
    const int slot = static_cast<int>(row % slots);
    const int64_t b = row / slots;
    const int a = slot / fields, c = slot - a * fields;
    if (VEC == 4) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a != c) {
        const int i = a < c ? a : c, j = a < c ? c : a;
        const int p = i * (2 * fields - i - 1) / 2 + (j - i - 1);
        const float4 g = ldg_stream_f4(reinterpret_cast<const float4*>(grad_out + (b * pairs + p) * embed) + chunk);
        const float4 w = ldg_stream_f4(reinterpret_cast<const float4*>(v + (b * slots + c * fields + a) * embed) + chunk);
        r = make_float4(g.x * w.x, g.y * w.y, g.z * w.z, g.w * w.w);
      }
      stg_stream_f4(reinterpret_cast<float4*>(grad_v) + item, r);
    } else {
      float r = 0.f;
      if (a != c) {
        const int i = a < c ? a : c, j = a < c ? c : a;
        const int p = i * (2 * fields - i - 1) / 2 + (j - i - 1);
        r = ldg_stream_f1(grad_out + (b * pairs + p) * embed + chunk) *
            ldg_stream_f1(v + (b * slots + c * fields + a) * embed + chunk);
      }
      grad_v[item] = r;
    }
  }
}

// IPN backward: out[b,p] = <x_i, x_j>  =>  grad_x[b,i,:] = sum_{j != i} grad_out[b, p(i,j)] * x[b,j,:].
// One warp per sample: the sample's fields and its pair gradients are staged in the warp's slice of shared memory
// (coalesced reads), then every lane produces (field, component) items with a loop over the other fields.
__global__ void __launch_bounds__(256) ipn_backward_kernel(const float* __restrict__ x, const float* __restrict__ grad_out,
                                                           int64_t batch, int fields, int embed, int warps_per_cta,
                                                           float* __restrict__ grad_x) {
  extern __shared__ __align__(16) float ipn_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= warps_per_cta) return;
  const int tile = fields * embed;
  const int pairs = fields * (fields - 1) / 2;
  float* xs = ipn_smem + (size_t)warp * (tile + pairs);
  float* gs = xs + tile;
  for (int64_t b = (int64_t)blockIdx.x * warps_per_cta + warp; b < batch; b += (int64_t)gridDim.x * warps_per_cta) {
    for (int t = lane; t < tile; t += 32) xs[t] = ldg_stream_f1(x + b * tile + t);
    for (int t = lane; t < pairs; t += 32) gs[t] = ldg_stream_f1(grad_out + b * pairs + t);
    __syncwarp();
    for (int t = lane; t < tile; t += 32) {
      const int i = t / embed, e = t - i * embed;
      float acc = 0.f;
      // pairs (j, i) with j < i sit at p_j + (i - j - 1); pairs (i, j) with j > i are contiguous from p_i
      for (int j = 0; j < i; ++j) acc = fmaf(gs[j * (2 * fields - j - 1) / 2 + (i - j - 1)], xs[j * embed + e], acc);
      const int base = i * (2 * fields - i - 1) / 2 - i - 1;
      for (int j = i + 1; j < fields; ++j) acc = fmaf(gs[base + j], xs[j * embed + e], acc);
      grad_x[b * tile + t] = acc;
    }
    __syncwarp();
  }
}

// Cross-network backward (cross_network.py:52-87).  Per row, with u_l = W_l h_l + b_l and h_{l+1} = x * u_l + x:
//     du_l = g_{l+1} * x ;  grad_x += g_{l+1} * (u_l + 1) ;  grad_W_l += du_l (x) h_l ;  grad_b_l += du_l ;
//     g_l  = W_l^T du_l ;  the chain stops at h_0 = x.detach() (cross_network.py:65: no gradient to x through h_0).
// Persistent CTAs over tiles of R = 8 * (256 / E) rows.  A tile re-runs the forward keeping every u_l in shared
// memory (h_l = x * (u_{l-1} + 1) is rebuilt from it), then walks the layers backwards.  Thread (o = tid % E,
// group = tid / E) owns column o of 8 rows: matrix-vector products read W with one conflict-free request per four k and
// the rows as 16-byte broadcasts.  grad_W / grad_b are accumulated in shared memory by their owner threads over all
// tiles of the CTA (no shared atomics) and added to global memory once per CTA (float atomics: the sum over CTAs is
// order-dependent in the last bits, like any split reduction).
constexpr int kCrossRT = 8;

template <int E>
__global__ void __launch_bounds__(256) cross_backward_kernel(const float* __restrict__ x, const float* __restrict__ weights,
                                                             const float* __restrict__ biases,
                                                             const float* __restrict__ grad_out, int64_t rows, int layers,
                                                             float* __restrict__ grad_x, float* __restrict__ grad_w,
                                                             float* __restrict__ grad_b) {
  constexpr int RG = 256 / E, R = RG * kCrossRT, WP = E + 4;   // WP: pitch of the staged weight matrix
  extern __shared__ __align__(16) float cb_smem[];
  float* hb = cb_smem;                        // [R][E] h_l of the layer at hand
  float* dus = hb + R * E;                    // [R][E] du_l
  float* ws = dus + R * E;                    // [E][WP] W_l
  float* us = ws + E * WP;                    // [layers][R][E] u_l (each entry written and read by its owner thread)
  float* dws = us + (size_t)layers * R * E;   // [layers][E][E]
  float* dbs = dws + (size_t)layers * E * E;  // [layers][E]
  const int tid = threadIdx.x;
  const int o = tid % E, rg = tid / E;
  const int r0 = rg * kCrossRT;

  for (int t = tid; t < layers * E * (E + 1); t += 256) dws[t] = 0.f;   // dws and dbs are adjacent

  const int64_t tiles = (rows + R - 1) / R;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t row_base = tile * R;
    float xr[kCrossRT], g[kCrossRT], dx[kCrossRT];
    __syncthreads();   // the previous tile's readers of hb are done
#pragma unroll
    for (int i = 0; i < kCrossRT; ++i) {
      const int64_t r = row_base + r0 + i;
      xr[i] = r < rows ? ldg_stream_f1(x + r * E + o) : 0.f;
      g[i] = r < rows ? ldg_stream_f1(grad_out + r * E + o) : 0.f;
      dx[i] = 0.f;
      hb[(r0 + i) * E + o] = xr[i];   // h_0 = x
    }
    // ---- forward: u_l for every layer ----
    for (int l = 0; l < layers; ++l) {
      for (int t = tid; t < E * E; t += 256) ws[(t / E) * WP + (t % E)] = __ldg(weights + (size_t)l * E * E + t);
      __syncthreads();
      float acc[kCrossRT];
      const float bias = __ldg(biases + l * E + o);
#pragma unroll
      for (int i = 0; i < kCrossRT; ++i) acc[i] = bias;
#pragma unroll 2
      for (int k = 0; k < E; k += 4) {
        const float4 w = *reinterpret_cast<const float4*>(ws + o * WP + k);
#pragma unroll
        for (int i = 0; i < kCrossRT; ++i) {
          const float4 h = *reinterpret_cast<const float4*>(hb + (r0 + i) * E + k);
          acc[i] = fmaf(w.x, h.x, acc[i]);
          acc[i] = fmaf(w.y, h.y, acc[i]);
          acc[i] = fmaf(w.z, h.z, acc[i]);
          acc[i] = fmaf(w.w, h.w, acc[i]);
        }
      }
      __syncthreads();   // everyone has read hb and ws
      float* ul = us + (size_t)l * R * E;
#pragma unroll
      for (int i = 0; i < kCrossRT; ++i) {
        const int at = (r0 + i) * E + o;
        ul[at] = acc[i];
        if (l + 1 < layers) hb[at] = fmaf(xr[i], acc[i], xr[i]);   // h_{l+1}
      }
    }
    // ---- backward ----
    for (int l = layers - 1; l >= 0; --l) {
      __syncthreads();   // previous readers of ws / hb / dus are done
      for (int t = tid; t < E * E; t += 256) ws[(t / E) * WP + (t % E)] = __ldg(weights + (size_t)l * E * E + t);
#pragma unroll
      for (int i = 0; i < kCrossRT; ++i) {
        const int at = (r0 + i) * E + o;
        dus[at] = g[i] * xr[i];
        dx[i] = fmaf(g[i], us[(size_t)l * R * E + at] + 1.f, dx[i]);
        hb[at] = l == 0 ? xr[i] : fmaf(xr[i], us[(size_t)(l - 1) * R * E + at], xr[i]);   // h_l
      }
      __syncthreads();
      // grad_W_l[o'][k4..k4+3] += sum_r du[r][o'] * h_l[r][k4..]   (owner thread per item, over the tile's rows)
      for (int item = tid; item < E * E / 4; item += 256) {
        const int oo = item / (E / 4), k4 = (item % (E / 4)) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
          const float d = dus[r * E + oo];
          const float4 h = *reinterpret_cast<const float4*>(hb + r * E + k4);
          a.x = fmaf(d, h.x, a.x);
          a.y = fmaf(d, h.y, a.y);
          a.z = fmaf(d, h.z, a.z);
          a.w = fmaf(d, h.w, a.w);
        }
        float4* dst = reinterpret_cast<float4*>(dws + (size_t)l * E * E + oo * E + k4);
        const float4 cur = *dst;
        *dst = make_float4(cur.x + a.x, cur.y + a.y, cur.z + a.z, cur.w + a.w);
      }
      if (tid < E) {
        float s = 0.f;
        for (int r = 0; r < R; ++r) s += dus[r * E + tid];
        dbs[l * E + tid] += s;
      }
      if (l > 0) {
        // g_l[r][k = o] = sum_o' W[o'][k] * du[r][o']
#pragma unroll
        for (int i = 0; i < kCrossRT; ++i) g[i] = 0.f;
#pragma unroll 2
        for (int q = 0; q < E; q += 4) {
          const float w0 = ws[(q + 0) * WP + o], w1 = ws[(q + 1) * WP + o], w2 = ws[(q + 2) * WP + o],
                      w3 = ws[(q + 3) * WP + o];
#pragma unroll
          for (int i = 0; i < kCrossRT; ++i) {
            const float4 d = *reinterpret_cast<const float4*>(dus + (r0 + i) * E + q);
            g[i] = fmaf(w0, d.x, g[i]);
            g[i] = fmaf(w1, d.y, g[i]);
            g[i] = fmaf(w2, d.z, g[i]);
            g[i] = fmaf(w3, d.w, g[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kCrossRT; ++i) {
      const int64_t r = row_base + r0 + i;
      if (r < rows) grad_x[r * E + o] = dx[i];
    }
  }
  __syncthreads();
  for (int t = tid; t < layers * E * E; t += 256) atomicAdd(grad_w + t, dws[t]);
  for (int t = tid; t < layers * E; t += 256) atomicAdd(grad_b + t, dbs[t]);
}

template <int E>
int cross_backward_run(const float* x, const float* w, const float* b, const float* g, int64_t rows, int layers,
                       float* gx, float* gw, float* gb, cudaStream_t s) {
  constexpr int R = (256 / E) * kCrossRT;
  const size_t floats = (size_t)2 * R * E + (size_t)E * (E + 4) + (size_t)layers * (R * E + E * E + E);
  const size_t smem = floats * sizeof(float);
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_cross_backward: %d layers of width %d need %zu bytes of shared memory",
                  layers, E, smem);
  TRS_SMEM_OPT_IN(cross_backward_kernel<E>);
  const int per_sm = static_cast<int>((size_t)kMaxDynSmem / smem) > 2 ? 2 : static_cast<int>((size_t)kMaxDynSmem / smem);
  const int64_t tiles = (rows + R - 1) / R;
  const int64_t cap = (int64_t)kNumSMs * (per_sm < 1 ? 1 : per_sm);
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  cross_backward_kernel<E><<<grid, 256, smem, s>>>(x, w, b, g, rows, layers, gx, gw, gb);
  return check_launch("cross_backward_kernel");
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_embedding_grad(const float* grad_out, const void* idx, int idx_bits, const int64_t* offsets,
                                  int64_t batch, int fields, int64_t rows, int embed, int64_t padding_row,
                                  float* grad_weight, void* stream) {
  TRS_REQUIRE(grad_out && idx && grad_weight, "trs_embedding_grad: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_embedding_grad: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0 && embed > 0, "trs_embedding_grad: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t lookups = batch * fields;
  if ((embed & 3) == 0 && aligned16(grad_out) && aligned16(grad_weight)) {
    const uint32_t chunks = embed / 4;
    const int64_t items = lookups * chunks;
    const int grid = grid_for(items, 256, 16);
    if (idx_bits == 64)
      embedding_grad_vec_kernel<64><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(grad_out), idx, offsets, rows,
                                                         chunks, fields, items, padding_row, grad_weight);
    else
      embedding_grad_vec_kernel<32><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(grad_out), idx, offsets, rows,
                                                         chunks, fields, items, padding_row, grad_weight);
    return check_launch("embedding_grad_vec_kernel");
  }
  const int64_t items = lookups * embed;
  const int grid = grid_for(items, 256, 16);
  if (idx_bits == 64)
    embedding_grad_scalar_kernel<64><<<grid, 256, 0, s>>>(grad_out, idx, offsets, rows, embed, fields, items, padding_row,
                                                          grad_weight);
  else
    embedding_grad_scalar_kernel<32><<<grid, 256, 0, s>>>(grad_out, idx, offsets, rows, embed, fields, items, padding_row,
                                                          grad_weight);
  return check_launch("embedding_grad_scalar_kernel");
}

extern "C" int trs_embedding_rows(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                  int64_t* out_rows, void* stream) {
  TRS_REQUIRE(idx && out_rows, "trs_embedding_rows: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_embedding_rows: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0, "trs_embedding_rows: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t lookups = batch * fields;
  const int grid = grid_for(lookups, 256, 8);
  if (idx_bits == 64)
    embedding_rows_kernel<64><<<grid, 256, 0, s>>>(idx, offsets, fields, lookups, out_rows);
  else
    embedding_rows_kernel<32><<<grid, 256, 0, s>>>(idx, offsets, fields, lookups, out_rows);
  return check_launch("embedding_rows_kernel");
}

extern "C" int trs_embedding_grad_segments(const float* grad_out, const int64_t* perm, const int64_t* starts,
                                           int64_t segments, int embed, float* out_values, void* stream) {
  TRS_REQUIRE(grad_out && perm && starts && out_values, "trs_embedding_grad_segments: null pointer");
  TRS_REQUIRE(segments >= 0 && embed > 0, "trs_embedding_grad_segments: bad sizes");
  if (segments == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((embed & 3) == 0 && aligned16(grad_out) && aligned16(out_values)) {
    const uint32_t chunks = embed / 4;
    embedding_grad_segments_kernel<4><<<grid_for(segments * chunks, 256, 16), 256, 0, s>>>(grad_out, perm, starts,
                                                                                            segments, chunks, out_values);
  } else {
    embedding_grad_segments_kernel<1><<<grid_for(segments * embed, 256, 16), 256, 0, s>>>(grad_out, perm, starts, segments,
                                                                                         embed, out_values);
  }
  return check_launch("embedding_grad_segments_kernel");
}

extern "C" int trs_fm_backward(const float* x, const float* grad_out, int64_t batch, int fields, int embed,
                               float* grad_x, void* stream) {
  TRS_REQUIRE(x && grad_out && grad_x, "trs_fm_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 0 && embed > 0, "trs_fm_backward: bad sizes");
  if (batch == 0) return TRS_OK;
  fm_backward_kernel<<<grid_for(batch * 32, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, grad_out, batch,
                                                                                                  fields, embed, grad_x);
  return check_launch("fm_backward_kernel");
}

extern "C" int trs_ffm_backward(const float* v, const float* grad_out, int64_t batch, int fields, int embed,
                                float* grad_v, void* stream) {
  TRS_REQUIRE(v && grad_out && grad_v, "trs_ffm_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_ffm_backward: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows = batch * fields * fields;
  if ((embed & 3) == 0 && aligned16(v) && aligned16(grad_out) && aligned16(grad_v)) {
    ffm_backward_kernel<4><<<grid_for(rows * (embed / 4), 256, 16), 256, 0, s>>>(v, grad_out, batch, fields, embed, grad_v);
  } else {
    ffm_backward_kernel<1><<<grid_for(rows * embed, 256, 16), 256, 0, s>>>(v, grad_out, batch, fields, embed, grad_v);
  }
  return check_launch("ffm_backward_kernel");
}

extern "C" int trs_ipn_backward(const float* x, const float* grad_out, int64_t batch, int fields, int embed,
                                float* grad_x, void* stream) {
  TRS_REQUIRE(x && grad_out && grad_x, "trs_ipn_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_ipn_backward: bad sizes");
  if (batch == 0) return TRS_OK;
  const size_t per_warp = ((size_t)fields * embed + (size_t)fields * (fields - 1) / 2) * sizeof(float);
  TRS_UNSUPPORTED(per_warp > (size_t)kMaxDynSmem, "trs_ipn_backward: %d fields x %d do not fit shared memory", fields,
                  embed);
  int warps = static_cast<int>((size_t)(96 * 1024) / per_warp);   // two CTAs per SM when the sample is small
  if (warps < 1) warps = static_cast<int>((size_t)kMaxDynSmem / per_warp);
  if (warps > 8) warps = 8;
  TRS_SMEM_OPT_IN(ipn_backward_kernel);
  const int64_t ctas = (batch + warps - 1) / warps;
  const int64_t cap = (int64_t)kNumSMs * 2;
  ipn_backward_kernel<<<static_cast<int>(ctas < cap ? ctas : cap), 256, per_warp * warps,
                        static_cast<cudaStream_t>(stream)>>>(x, grad_out, batch, fields, embed, warps, grad_x);
  return check_launch("ipn_backward_kernel");
}

extern "C" int trs_cross_backward(const float* x, const float* weights, const float* biases, const float* grad_out,
                                  int layers, int64_t rows, int embed, float* grad_x, float* grad_weights,
                                  float* grad_biases, void* stream) {
  TRS_REQUIRE(x && weights && biases && grad_out && grad_x && grad_weights && grad_biases,
              "trs_cross_backward: null pointer");
  TRS_REQUIRE(layers > 0 && rows >= 0 && embed > 0, "trs_cross_backward: bad sizes");
  TRS_UNSUPPORTED(embed != 8 && embed != 16 && embed != 32 && embed != 64,
                  "trs_cross_backward: embed must be 8, 16, 32 or 64 (got %d)", embed);
  TRS_UNSUPPORTED(!aligned16(x) || !aligned16(grad_out), "trs_cross_backward: x and grad_out must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TRS_CUDA(cudaMemsetAsync(grad_weights, 0, (size_t)layers * embed * embed * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_biases, 0, (size_t)layers * embed * sizeof(float), s));
  if (rows == 0) return TRS_OK;
  switch (embed) {
    case 8: return cross_backward_run<8>(x, weights, biases, grad_out, rows, layers, grad_x, grad_weights, grad_biases, s);
    case 16: return cross_backward_run<16>(x, weights, biases, grad_out, rows, layers, grad_x, grad_weights, grad_biases, s);
    case 32: return cross_backward_run<32>(x, weights, biases, grad_out, rows, layers, grad_x, grad_weights, grad_biases, s);
    default: return cross_backward_run<64>(x, weights, biases, grad_out, rows, layers, grad_x, grad_weights, grad_biases, s);
  }
}
