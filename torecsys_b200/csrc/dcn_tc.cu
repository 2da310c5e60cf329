// DeepAndCrossNetworkModel forward, indices -> logits, with the per-row dense chains on the tensor pipe.
//
// Work per (sample, field) row of E floats (SURVEY.md 8a rows a7 + a12, cfg 3: E = 32, 6 cross layers, MLP 32-32-16-8-4):
//     cross: h <- x * (W_l h + b_l) + x  (6 x 32x32 mat-vec)      deep: relu MLP per row      fc: dot with fc_w[n, :]
// = 7 840 MACs per row, 636.8 KFLOP per sample against 5 308 B -> compute bound; FP32 FFMA tops out near 110 M
// samples/s, so the chains run as mma.sync.m16n8k8 TF32 with the 3xTF32 split (fp32-accurate), fully in registers:
//   * a warp owns 16 rows; lane (g,t) holds columns {8j+2t, 8j+2t+1} of rows g and g+8 = the mma ACCUMULATOR layout.
//     With the k order permuted accordingly (baked into the B fragments) the same registers are the A operand of the
//     next layer, so a 6-layer cross chain + 4-layer MLP needs no shuffles and no shared-memory round trips;
//   * weights of all layers live in shared memory as pre-split (hi, lo) B fragments, one conflict-free LDS.64 each;
//   * a CTA takes 16 whole samples (= N warp tiles of 16 rows, any N), rows are staged global -> shared with one
//     cp.async request per row (double buffered per warp), per-row fc partials go to shared memory and 16 threads
//     add the N partials of each sample in a fixed order: deterministic, no atomics.
// Shapes outside (E in {8,16,32,64}, layer widths <= max(E,32), ReLU, one output) use the generic kernel in
// fused_models.cu.
#include <stdlib.h>

#include "tile_ops.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;
constexpr int kSamples = 16;   // samples per CTA step

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

struct DenseDesc {   // one dense layer in the shared-memory fragment store
  int frag_off;      // float2 index of frag[ks][nt][hi|lo][lane]
  int bias_off;      // float index of the (padded) bias
  int ks, nt;        // k-steps (in/8) and n-tiles (out/8), padded
};

struct DcnTcArgs {
  const void* idx;
  const int64_t* offsets;
  const float* w_emb;
  const float* cross_w;   // (L, E, E)
  const float* cross_b;   // (L, E)
  const float* fc_w;      // (1, N*(E+Od))
  const float* fc_b;
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, cross_layers, od;
  MlpParams mp;
};

constexpr int kMaxDense = 16;

template <int E>
struct Cfg {
  static constexpr int D = E < 32 ? 32 : E;   // widest activation held in registers
  static constexpr int T = D / 8;             // n-tiles / k-steps of the widest layer
  static constexpr int kPitch = E + 8;        // staging row pitch (floats): conflict-free LDS.64 in accumulator layout
};

// h (accumulator layout, `ks` k-steps wide) -> out (accumulator layout, `nt` tiles) = W h, 3xTF32
template <int T>
__device__ __forceinline__ void dense_mma(const float (&h)[T][4], int ks, int nt, const float2* __restrict__ frag,
                                          int lane, float (&out)[T][4]) {
#pragma unroll
  for (int j = 0; j < T; ++j) out[j][0] = out[j][1] = out[j][2] = out[j][3] = 0.f;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    if (k < ks) {
      uint32_t ah[4], al[4];
      split_fast(h[k][0], ah[0], al[0]);   // row g,   k = 8k + 2t
      split_fast(h[k][2], ah[1], al[1]);   // row g+8
      split_fast(h[k][1], ah[2], al[2]);   // row g,   k = 8k + 2t + 1
      split_fast(h[k][3], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < T; ++j) {
        if (j < nt) {
          const float2 wh = frag[((k * nt + j) * 2 + 0) * 32 + lane];
          const float2 wl = frag[((k * nt + j) * 2 + 1) * 32 + lane];
          mma_tf32(out[j], al, __float_as_uint(wh.x), __float_as_uint(wh.y));
          mma_tf32(out[j], ah, __float_as_uint(wl.x), __float_as_uint(wl.y));
          mma_tf32(out[j], ah, __float_as_uint(wh.x), __float_as_uint(wh.y));
        }
      }
    }
  }
}

template <int E, int IdxBits>
__global__ void __launch_bounds__(kWarps * 32, E <= 32 ? 2 : 1) dcn_tc_kernel(DcnTcArgs a) {
  using C = Cfg<E>;
  constexpr int T = C::T, KE = E / 8, kPitch = C::kPitch;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ DenseDesc descs[kMaxDense];
  const int n_fields = a.fields;
  const int n_dense = a.cross_layers + a.mp.layers;
  const int cat = E + a.od;

  // ---- carve: fragments | biases | fc_w | per-warp staging (2 buffers) | per-row partials ---------------------------
  int frag_total = 0, bias_total = 0;
  for (int l = 0; l < n_dense; ++l) {
    const int in = l < a.cross_layers ? E : a.mp.dims[l - a.cross_layers];
    const int out = l < a.cross_layers ? E : a.mp.dims[l - a.cross_layers + 1];
    const int ks = (in + 7) / 8, nt = (out + 7) / 8;
    if (threadIdx.x == 0) descs[l] = DenseDesc{frag_total, bias_total, ks, nt};
    frag_total += ks * nt * 2 * 32;
    bias_total += nt * 8;
  }
  float2* frags = reinterpret_cast<float2*>(smem_raw);
  float* bias_s = reinterpret_cast<float*>(frags + frag_total);
  float* fcw_s = bias_s + bias_total;                              // [N][cat]
  float* stage_all = fcw_s + ((n_fields * cat + 3) & ~3);          // [warps][2][16][kPitch]
  float* part = stage_all + kWarps * 2 * 16 * kPitch;              // [16 * N]
  __syncthreads();

  // ---- one-time: B fragments (hi/lo) of every dense layer; b0 = W[8nt+g][8ks+2t], b1 = W[8nt+g][8ks+2t+1] ---------
  for (int l = 0; l < n_dense; ++l) {
    const bool is_cross = l < a.cross_layers;
    const int in = is_cross ? E : a.mp.dims[l - a.cross_layers];
    const int out = is_cross ? E : a.mp.dims[l - a.cross_layers + 1];
    const float* w = is_cross ? a.cross_w + (size_t)l * E * E : a.mp.w[l - a.cross_layers];
    const float* b = is_cross ? a.cross_b + (size_t)l * E : a.mp.b[l - a.cross_layers];
    const DenseDesc d = descs[l];
    for (int i = threadIdx.x; i < d.ks * d.nt * 32; i += blockDim.x) {
      const int ln = i & 31, j = (i >> 5) % d.nt, k = (i >> 5) / d.nt;
      const int o = 8 * j + (ln >> 2), c = 8 * k + 2 * (ln & 3);
      const float w0 = (o < out && c < in) ? __ldg(w + (size_t)o * in + c) : 0.f;
      const float w1 = (o < out && c + 1 < in) ? __ldg(w + (size_t)o * in + c + 1) : 0.f;
      const uint32_t h0 = tf32_rna(w0), h1 = tf32_rna(w1);
      const uint32_t l0 = tf32_rna(w0 - __uint_as_float(h0)), l1 = tf32_rna(w1 - __uint_as_float(h1));
      frags[d.frag_off + ((k * d.nt + j) * 2 + 0) * 32 + ln] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
      frags[d.frag_off + ((k * d.nt + j) * 2 + 1) * 32 + ln] = make_float2(__uint_as_float(l0), __uint_as_float(l1));
    }
    for (int i = threadIdx.x; i < d.nt * 8; i += blockDim.x) bias_s[d.bias_off + i] = (i < out && b) ? __ldg(b + i) : 0.f;
  }
  for (int i = threadIdx.x; i < n_fields * cat; i += blockDim.x) fcw_s[i] = __ldg(a.fc_w + i);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* stage = stage_all + warp * 2 * 16 * kPitch;
  constexpr int kChunks = E / 4;                 // 16-byte chunks per row
  constexpr int kCopies = 16 * kChunks / 32;     // cp.async per lane per tile (E=32: 4)
  static_assert(16 * kChunks % 32 == 0, "tile copies must divide over the warp");
  const int64_t groups = (a.batch + kSamples - 1) / kSamples;
  const float fc_bias = __ldg(a.fc_b);

  // Every warp walks its own sequence of 16-row tiles: tile w, w+8, ... of group blockIdx.x, then of the next group
  // of this CTA, and so on.  The raw indices of tile q+2 and the rows of tile q+1 are in flight (cp.async) while tile q
  // is computed, ACROSS group boundaries -- a warp never waits for an index load it has just issued.
  // The raw index of row r of tile q waits in the 32 unused bytes behind row r of staging buffer q & 1 (the row pitch
  // is E + 8 floats for conflict-free fragment loads): two tiles of indices are alive at a time, no extra shared memory.
  constexpr int kIdxBytes = IdxBits / 8;
  auto idx_slot = [&](int64_t q, int r) -> unsigned char* {
    return reinterpret_cast<unsigned char*>(stage + ((q & 1) * 16 + r) * kPitch + E);
  };
  const int64_t all_rows = a.batch * n_fields;
  const int cnt_w = warp < n_fields ? (n_fields - warp + kWarps - 1) / kWarps : 0;   // tiles of this warp per group
  const int64_t my_groups = blockIdx.x < groups ? (groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t my_tiles = my_groups * cnt_w;
  auto tile_row0 = [&](int64_t q) -> int64_t {
    const int64_t ord = q / cnt_w;
    const int j = static_cast<int>(q - ord * cnt_w);
    return ((int64_t)blockIdx.x + ord * gridDim.x) * kSamples * n_fields + (int64_t)(warp + kWarps * j) * 16;
  };
  // raw indices of tile q (lanes 0-15, one index each); always commits a group
  auto issue_idx = [&](int64_t q) {
    if (q < my_tiles && lane < 16) {
      const int64_t row = tile_row0(q) + lane;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(idx_slot(q, lane)));
      const unsigned char* src = static_cast<const unsigned char*>(a.idx) + (row < all_rows ? row : 0) * kIdxBytes;
      const int bytes = row < all_rows ? kIdxBytes : 0;
      if (IdxBits == 64)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
      else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // rows of tile q (its indices have landed) -> staging buffer q & 1; always commits a group
  auto issue_rows = [&](int64_t q) {
    if (q < my_tiles) {
      const int64_t row0 = tile_row0(q);
      const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(stage + (q & 1) * 16 * kPitch));
#pragma unroll
      for (int k = 0; k < kCopies; ++k) {
        const int c = lane + 32 * k;
        const int r = c / kChunks, ch = c - r * kChunks;
        const int64_t row = row0 + r;
        int src_bytes = 0;
        const float* src = a.w_emb;
        if (row < all_rows) {
          const int n = static_cast<int>(row % n_fields);
          const unsigned char* slot = idx_slot(q, r);
          const int64_t raw = IdxBits == 64 ? *reinterpret_cast<const long long*>(slot)
                                            : static_cast<int64_t>(*reinterpret_cast<const int*>(slot));
          const int64_t ix = raw + __ldg(a.offsets + n);
          if (ix >= 0 && ix < a.rows) {
            src = a.w_emb + ix * E + 4 * ch;
            src_bytes = 16;
          } else if (ch == 0) {
            report_oob(a.status, row);
          }
        }
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (r * kPitch + 4 * ch) * 4), "l"(src),
                     "r"(src_bytes) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue_idx(0);
  issue_idx(1);
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncwarp();
  issue_rows(0);
  int64_t seq = 0;
  for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    const int64_t b0 = grp * kSamples;
    const int nsamp = static_cast<int>(a.batch - b0 < kSamples ? a.batch - b0 : kSamples);
    for (int tile = warp; tile < n_fields; tile += kWarps, ++seq) {
      const int buf = static_cast<int>(seq & 1);
      issue_idx(seq + 2);
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // indices of tile seq+1 and rows of tile seq have landed
      __syncwarp();
      issue_rows(seq + 1);
      // ---- x in accumulator layout ---------------------------------------------------------------------------------
      const float* sx = stage + buf * 16 * kPitch;
      float x[T][4], h[T][4], acc[T][4];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        if (j < KE) {
          const float2 lo = *reinterpret_cast<const float2*>(sx + g * kPitch + 8 * j + 2 * t);
          const float2 hi = *reinterpret_cast<const float2*>(sx + (g + 8) * kPitch + 8 * j + 2 * t);
          x[j][0] = lo.x; x[j][1] = lo.y; x[j][2] = hi.x; x[j][3] = hi.y;
        } else {
          x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) h[j][q] = x[j][q];
      }
      __syncwarp();
      // ---- cross network ---------------------------------------------------------------------------------------------
      for (int l = 0; l < a.cross_layers; ++l) {
        const DenseDesc d = descs[l];
        dense_mma<T>(h, KE, KE, frags + d.frag_off, lane, acc);
#pragma unroll
        for (int j = 0; j < T; ++j) {
          if (j < KE) {
            const float bb0 = bias_s[d.bias_off + 8 * j + 2 * t], bb1 = bias_s[d.bias_off + 8 * j + 2 * t + 1];
            h[j][0] = fmaf(x[j][0], acc[j][0] + bb0, x[j][0]);
            h[j][1] = fmaf(x[j][1], acc[j][1] + bb1, x[j][1]);
            h[j][2] = fmaf(x[j][2], acc[j][2] + bb0, x[j][2]);
            h[j][3] = fmaf(x[j][3], acc[j][3] + bb1, x[j][3]);
          }
        }
      }
      // fc over the cross half: rows r_a = tile*16 + g, r_b = r_a + 8 (local to the group); field = row % N
      const int ra = tile * 16 + g, rb = ra + 8;
      const float* fa = fcw_s + (ra % n_fields) * cat;
      const float* fb = fcw_s + (rb % n_fields) * cat;
      float dot_a = 0.f, dot_b = 0.f;
#pragma unroll
      for (int j = 0; j < T; ++j) {
        if (j < KE) {
          const int c = 8 * j + 2 * t;
          dot_a = fmaf(h[j][0], fa[c], dot_a);
          dot_a = fmaf(h[j][1], fa[c + 1], dot_a);
          dot_b = fmaf(h[j][2], fb[c], dot_b);
          dot_b = fmaf(h[j][3], fb[c + 1], dot_b);
        }
      }
      // ---- per-field MLP on the same rows ------------------------------------------------------------------------------
#pragma unroll
      for (int j = 0; j < T; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) h[j][q] = x[j][q];
      for (int l = 0; l < a.mp.layers; ++l) {
        const DenseDesc d = descs[a.cross_layers + l];
        dense_mma<T>(h, d.ks, d.nt, frags + d.frag_off, lane, acc);
        const bool last = l == a.mp.layers - 1;
#pragma unroll
        for (int j = 0; j < T; ++j) {
          const float bb0 = j < d.nt ? bias_s[d.bias_off + 8 * j + 2 * t] : 0.f;
          const float bb1 = j < d.nt ? bias_s[d.bias_off + 8 * j + 2 * t + 1] : 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v = j < d.nt ? acc[j][q] + ((q & 1) ? bb1 : bb0) : 0.f;
            h[j][q] = last ? v : fmaxf(v, 0.f);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < T; ++j) {
        const int c = 8 * j + 2 * t;
        if (c < a.od) {
          dot_a = fmaf(h[j][0], fa[E + c], dot_a);
          dot_b = fmaf(h[j][2], fb[E + c], dot_b);
        }
        if (c + 1 < a.od) {
          dot_a = fmaf(h[j][1], fa[E + c + 1], dot_a);
          dot_b = fmaf(h[j][3], fb[E + c + 1], dot_b);
        }
      }
      dot_a += __shfl_xor_sync(0xffffffffu, dot_a, 1);
      dot_b += __shfl_xor_sync(0xffffffffu, dot_b, 1);
      dot_a += __shfl_xor_sync(0xffffffffu, dot_a, 2);
      dot_b += __shfl_xor_sync(0xffffffffu, dot_b, 2);
      if (t == 0) {
        part[ra] = dot_a;
        part[rb] = dot_b;
      }
    }
    __syncthreads();
    if (threadIdx.x < nsamp) {
      float s = fc_bias;
      const float* p = part + threadIdx.x * n_fields;
      for (int n = 0; n < n_fields; ++n) s += p[n];
      a.logits[b0 + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

// ---- a7 as an L1 op: CrossNetworkLayer on materialised rows (rows, E) -> (rows, E), same register-resident chain ----
struct CrossTcArgs {
  const float* x;
  const float* w;   // (L, E, E)
  const float* b;   // (L, E)
  float* out;
  int64_t rows;
  int layers;
};

template <int E>
__global__ void __launch_bounds__(kWarps * 32, E <= 32 ? 2 : 1) cross_tc_kernel(CrossTcArgs a) {
  using C = Cfg<E>;
  constexpr int KE = E / 8, kPitch = C::kPitch;
  constexpr int kChunks = E / 4, kCopies = 16 * kChunks / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* frags = reinterpret_cast<float2*>(smem_raw);                       // [L][KE][KE][hi|lo][32]
  float* bias_s = reinterpret_cast<float*>(frags + (size_t)a.layers * KE * KE * 2 * 32);   // [L][E]
  float* stage_all = bias_s + ((a.layers * E + 3) & ~3);                      // [warps][2][16][kPitch]
  for (int i = threadIdx.x; i < a.layers * KE * KE * 32; i += blockDim.x) {
    const int ln = i & 31, j = (i >> 5) % KE, k = ((i >> 5) / KE) % KE, l = (i >> 5) / (KE * KE);
    const float* w = a.w + (size_t)l * E * E + (size_t)(8 * j + (ln >> 2)) * E + 8 * k + 2 * (ln & 3);
    const float w0 = __ldg(w), w1 = __ldg(w + 1);
    const uint32_t h0 = tf32_rna(w0), h1 = tf32_rna(w1);
    const uint32_t l0 = tf32_rna(w0 - __uint_as_float(h0)), l1 = tf32_rna(w1 - __uint_as_float(h1));
    frags[(((l * KE + k) * KE + j) * 2 + 0) * 32 + ln] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
    frags[(((l * KE + k) * KE + j) * 2 + 1) * 32 + ln] = make_float2(__uint_as_float(l0), __uint_as_float(l1));
  }
  for (int i = threadIdx.x; i < a.layers * E; i += blockDim.x) bias_s[i] = __ldg(a.b + i);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  float* stage = stage_all + warp * 2 * 16 * kPitch;
  const int64_t tiles = (a.rows + 15) / 16;
  const int64_t wstride = (int64_t)gridDim.x * kWarps;
  auto issue_tile = [&](int64_t tile, int buf) {
    const uint32_t dst0 = static_cast<uint32_t>(__cvta_generic_to_shared(stage + buf * 16 * kPitch));
#pragma unroll
    for (int k = 0; k < kCopies; ++k) {
      const int c = lane + 32 * k;
      const int r = c / kChunks, ch = c - r * kChunks;
      const int64_t row = tile * 16 + r;
      const bool ok = row < a.rows;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + (r * kPitch + 4 * ch) * 4),
                   "l"(a.x + (ok ? row * E + 4 * ch : 0)), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  int64_t tile = (int64_t)blockIdx.x * kWarps + warp;
  if (tile < tiles) issue_tile(tile, 0);
  for (; tile < tiles; tile += wstride) {
    if (tile + wstride < tiles) {
      issue_tile(tile + wstride, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    float* sx = stage + buf * 16 * kPitch;
    float x[KE][4], h[KE][4], acc[KE][4];
#pragma unroll
    for (int j = 0; j < KE; ++j) {
      const float2 lo = *reinterpret_cast<const float2*>(sx + g * kPitch + 8 * j + 2 * t);
      const float2 hi = *reinterpret_cast<const float2*>(sx + (g + 8) * kPitch + 8 * j + 2 * t);
      x[j][0] = lo.x; x[j][1] = lo.y; x[j][2] = hi.x; x[j][3] = hi.y;
#pragma unroll
      for (int q = 0; q < 4; ++q) h[j][q] = x[j][q];
    }
    for (int l = 0; l < a.layers; ++l) {
      dense_mma<KE>(h, KE, KE, frags + (size_t)l * KE * KE * 2 * 32, lane, acc);
#pragma unroll
      for (int j = 0; j < KE; ++j) {
        const float bb0 = bias_s[l * E + 8 * j + 2 * t], bb1 = bias_s[l * E + 8 * j + 2 * t + 1];
        h[j][0] = fmaf(x[j][0], acc[j][0] + bb0, x[j][0]);
        h[j][1] = fmaf(x[j][1], acc[j][1] + bb1, x[j][1]);
        h[j][2] = fmaf(x[j][2], acc[j][2] + bb0, x[j][2]);
        h[j][3] = fmaf(x[j][3], acc[j][3] + bb1, x[j][3]);
      }
    }
    // back through the staging buffer so that the global stores are whole 16-byte chunks of contiguous rows
    __syncwarp();
#pragma unroll
    for (int j = 0; j < KE; ++j) {
      *reinterpret_cast<float2*>(sx + g * kPitch + 8 * j + 2 * t) = make_float2(h[j][0], h[j][1]);
      *reinterpret_cast<float2*>(sx + (g + 8) * kPitch + 8 * j + 2 * t) = make_float2(h[j][2], h[j][3]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kCopies; ++k) {
      const int c = lane + 32 * k;
      const int r = c / kChunks, ch = c - r * kChunks;
      const int64_t row = tile * 16 + r;
      if (row < a.rows)
        *reinterpret_cast<float4*>(a.out + row * E + 4 * ch) = *reinterpret_cast<const float4*>(sx + r * kPitch + 4 * ch);
    }
    __syncwarp();
    buf ^= 1;
  }
}

template <int E>
int cross_tc_dispatch(const CrossTcArgs& a, cudaStream_t s) {
  constexpr int KE = E / 8;
  const size_t smem = (size_t)a.layers * KE * KE * 2 * 32 * sizeof(float2) + (size_t)((a.layers * E + 3) & ~3) * 4 +
                      (size_t)kWarps * 2 * 16 * Cfg<E>::kPitch * sizeof(float);
  if (smem > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    TRS_CUDA(cudaFuncSetAttribute(cross_tc_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t tiles = (a.rows + 15) / 16;
  const int64_t ctas = (tiles + kWarps - 1) / kWarps;
  const int grid = static_cast<int>(ctas < 2 * kNumSMs ? ctas : 2 * kNumSMs);
  cross_tc_kernel<E><<<grid, kWarps * 32, smem, s>>>(a);
  return check_launch("cross_tc_kernel");
}

template <int E>
size_t dcn_tc_smem(const DcnTcArgs& a) {
  size_t frag = 0, bias = 0;
  for (int l = 0; l < a.cross_layers + a.mp.layers; ++l) {
    const int in = l < a.cross_layers ? E : a.mp.dims[l - a.cross_layers];
    const int out = l < a.cross_layers ? E : a.mp.dims[l - a.cross_layers + 1];
    const int ks = (in + 7) / 8, nt = (out + 7) / 8;
    frag += (size_t)ks * nt * 2 * 32 * sizeof(float2);
    bias += (size_t)nt * 8 * sizeof(float);
  }
  const size_t fcw = (size_t)((a.fields * (E + a.od) + 3) & ~3) * sizeof(float);
  const size_t stage = (size_t)kWarps * 2 * 16 * Cfg<E>::kPitch * sizeof(float);
  const size_t part = (size_t)kSamples * a.fields * sizeof(float);
  return frag + bias + fcw + stage + part;
}

template <int E>
int dcn_tc_dispatch(const DcnTcArgs& a, int idx_bits, cudaStream_t s) {
  const size_t smem = dcn_tc_smem<E>(a);
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_dcn_forward: layer stack does not fit shared memory");
  const int64_t groups = (a.batch + kSamples - 1) / kSamples;
  const int grid = static_cast<int>(groups < 2 * kNumSMs ? groups : 2 * kNumSMs);
  // the kernel also has static shared memory (layer descriptors), so opt in to what is needed, not to the maximum
  static size_t configured[2] = {0, 0};
  if (idx_bits == 64) {
    if (smem > configured[0]) {
      TRS_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<E, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[0] = smem;
    }
    dcn_tc_kernel<E, 64><<<grid, kWarps * 32, smem, s>>>(a);
  } else {
    if (smem > configured[1]) {
      TRS_CUDA(cudaFuncSetAttribute(dcn_tc_kernel<E, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[1] = smem;
    }
    dcn_tc_kernel<E, 32><<<grid, kWarps * 32, smem, s>>>(a);
  }
  return check_launch("dcn_tc_kernel");
}

}  // namespace

// CrossNetworkLayer (L1) on the tensor pipe; returns TRS_ERR_UNSUPPORTED when the shape is not covered
int cross_tc_launch(const float* x, const float* w, const float* b, int layers, int64_t rows, int embed, float* out,
                    cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || layers < 1 || !aligned16(x) || !aligned16(out)) return TRS_ERR_UNSUPPORTED;
  CrossTcArgs a{x, w, b, out, rows, layers};
  switch (embed) {
    case 8: return cross_tc_dispatch<8>(a, s);
    case 16: return cross_tc_dispatch<16>(a, s);
    case 32: return cross_tc_dispatch<32>(a, s);
    case 64: return cross_tc_dispatch<64>(a, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

int dcn_tc_supported(int embed, int cross_layers, const int* mlp_dims, int mlp_layers, int activation) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled) return 0;
  if (!(embed == 8 || embed == 16 || embed == 32 || embed == 64)) return 0;
  if (activation != TRS_ACT_RELU || cross_layers + mlp_layers > kMaxDense || mlp_layers < 1) return 0;
  const int dmax = embed < 32 ? 32 : embed;
  for (int l = 0; l <= mlp_layers; ++l)
    if (mlp_dims[l] > dmax) return 0;
  return 1;
}

int dcn_tc_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                  const float* w_emb, int64_t rows, int embed, const float* cross_w, const float* cross_b,
                  int cross_layers, const MlpParams& mp, const float* fc_w, const float* fc_b, float* logits,
                  int32_t* status, cudaStream_t s) {
  DcnTcArgs a{};
  a.idx = idx; a.offsets = offsets; a.w_emb = w_emb; a.cross_w = cross_w; a.cross_b = cross_b; a.fc_w = fc_w;
  a.fc_b = fc_b; a.logits = logits; a.status = status; a.batch = batch; a.rows = rows; a.fields = fields;
  a.cross_layers = cross_layers; a.od = mp.dims[mp.layers]; a.mp = mp;
  TRS_REQUIRE(aligned16(w_emb), "trs_dcn_forward: w_emb must be 16-byte aligned");
  switch (embed) {
    case 8: return dcn_tc_dispatch<8>(a, idx_bits, s);
    case 16: return dcn_tc_dispatch<16>(a, idx_bits, s);
    case 32: return dcn_tc_dispatch<32>(a, idx_bits, s);
    case 64: return dcn_tc_dispatch<64>(a, idx_bits, s);
  }
  set_error("trs_dcn_forward: unsupported embed %d for the tensor-core path", embed);
  return TRS_ERR_UNSUPPORTED;
}

}  // namespace trs
