// DeepFM forward, indices -> logits, for the Criteo-shaped configuration (embed = 16, hidden widths = 16).
//
// HBM-gather bound (cfg 2): per sample 39 x 8 B of indices (coalesced), 39 random 64 B table
// rows, 39 random 4 B first-order weights, 4 B out = 2 968 algorithmic bytes.  Everything else must hide under
// those loads, so nothing is staged through shared memory except the row ids:
//   * a warp owns 16 samples; lane (g = lane/4, t = lane%4) loads the 16-byte chunk t of the rows of samples g
//     and g+8 with 128-bit no-allocate loads, 4 fields per step = 8 independent LDG.128 in flight per lane;
//   * FM (sum, sum of squares) accumulates in the lane's registers, reduced over t with two shuffles at the end;
//   * that (sample, column) ownership IS the A-fragment layout of mma.sync.m16n8k8 (row = g / g+8, col = t / t+4),
//     so the first MLP layer (K = 16 N, 16 outputs) runs on the tensor pipe straight from the loaded registers,
//     FP32-accurate through the 3xTF32 split (hi*hi + lo*hi + hi*lo, fp32 accumulate);  W1 is pre-split into
//     hi/lo B fragments in shared memory once per CTA.  The accumulator layout of one layer is again the
//     A layout of the next (with a permuted k order baked into the B fragments), so the 16x16 hidden layers and
//     the output dot product need no shuffles either.
// Tensor cores are used here only because FFMA for the 9 984 MACs/sample of layer 1 would sit too close to the
// gather time; the kernel's roofline is HBM, not the tensor pipe.
#include <stdlib.h>

#include "common.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;
constexpr int kTile = 16;       // samples per warp tile
constexpr int kUF = 4;          // fields per load step
constexpr int kMaxHidden = 4;   // 16x16 hidden layers after the first one

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// d += A * B with A, B given as (hi, lo) TF32 pairs: lo terms first, then hi*hi
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al[0], al[1], al[2], al[3], bh0, bh1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma_tf32(d, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

struct FastArgs {
  const void* idx;
  const int64_t* offsets;
  const float* w_feat;
  const float* w_emb;
  const float* w1;              // (16, 16 N)
  const float* b1;              // (16)
  const float* wh[kMaxHidden];  // (16, 16) each
  const float* bh[kMaxHidden];  // (16)
  const float* w_out;           // (1, 16)
  const float* b_out;           // (1)
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, hidden_layers;
  float* x_out;        // null, or (B, N, 16): the gathered rows, for a consumer behind this kernel (xDeepFM's CIN)
  const float* bias;   // null, or a scalar added to every logit (xDeepFM's model bias)
  int use_fm;          // 0: no second-order FM term (xDeepFM: first-order + MLP only)
};

// shared memory carve-up (bytes): [W1 frags: N*2*2*32 float4][hidden frags: L*2*2*2*32 float2][biases][rid]
template <int IdxBits>
__global__ void __launch_bounds__(kWarps * 32, 2) deepfm_fast_kernel(FastArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_fields = a.fields;
  float4* w1s = reinterpret_cast<float4*>(smem_raw);
  float2* whs = reinterpret_cast<float2*>(w1s + (size_t)n_fields * 4 * 32);
  float* bias_s = reinterpret_cast<float*>(whs + (size_t)a.hidden_layers * 8 * 32);  // [(1+L)*16 + 16 + 1]
  int* rid_all = reinterpret_cast<int*>(bias_s + (1 + kMaxHidden) * 16 + 16 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int kdim = 16 * n_fields;

  // ---- one-time per CTA: B fragments of W1 (hi/lo), of the hidden layers, biases --------------------------------
  for (int i = threadIdx.x; i < n_fields * 2 * 32; i += blockDim.x) {
    const int l = i & 31, j = (i >> 5) & 1, n = i >> 6;
    const int lg = l >> 2, lt = l & 3;
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.w1 + (size_t)(8 * j + lg) * kdim + 16 * n + 4 * lt));
    uint32_t h[4], lo[4];
    split_tf32(w.x, h[0], lo[0]);
    split_tf32(w.y, h[1], lo[1]);
    split_tf32(w.z, h[2], lo[2]);
    split_tf32(w.w, h[3], lo[3]);
    w1s[((n * 2 + j) * 2 + 0) * 32 + l] =
        make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
    w1s[((n * 2 + j) * 2 + 1) * 32 + l] =
        make_float4(__uint_as_float(lo[0]), __uint_as_float(lo[1]), __uint_as_float(lo[2]), __uint_as_float(lo[3]));
  }
  for (int i = threadIdx.x; i < a.hidden_layers * 4 * 32; i += blockDim.x) {
    const int l = i & 31, jk = (i >> 5) & 1, jn = (i >> 6) & 1, layer = i >> 7;
    const int lg = l >> 2, lt = l & 3;
    const float* w = a.wh[layer] + (8 * jn + lg) * 16 + 8 * jk + 2 * lt;
    uint32_t h0, l0, h1, l1;
    split_tf32(__ldg(w), h0, l0);
    split_tf32(__ldg(w + 1), h1, l1);
    whs[(((layer * 2 + jn) * 2 + jk) * 2 + 0) * 32 + l] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
    whs[(((layer * 2 + jn) * 2 + jk) * 2 + 1) * 32 + l] = make_float2(__uint_as_float(l0), __uint_as_float(l1));
  }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) {
    bias_s[i] = __ldg(a.b1 + i);
    for (int l = 0; l < a.hidden_layers; ++l) bias_s[(1 + l) * 16 + i] = __ldg(a.bh[l] + i);
    bias_s[(1 + kMaxHidden) * 16 + i] = __ldg(a.w_out + i);
  }
  if (threadIdx.x == 0) bias_s[(1 + kMaxHidden) * 16 + 16] = __ldg(a.b_out);
  __syncthreads();

  int* rid = rid_all + warp * (kTile * n_fields);
  const float4* emb4 = reinterpret_cast<const float4*>(a.w_emb);
  const int64_t tiles = (a.batch + kTile - 1) / kTile;

  for (int64_t tile = (int64_t)blockIdx.x * kWarps + warp; tile < tiles; tile += (int64_t)gridDim.x * kWarps) {
    const int64_t b0 = tile * kTile;
    // ---- stage row ids of the 16 samples (coalesced index loads, offsets added, range checked) ----------------
    for (int k = lane; k < kTile * n_fields; k += 32) {
      const int s = k / n_fields, n = k - s * n_fields;
      const bool live = b0 + s < a.batch;
      const int64_t pos = (live ? b0 + s : a.batch - 1) * n_fields + n;
      int64_t r = load_index<IdxBits>(a.idx, pos) + __ldg(a.offsets + n);
      if (r < 0 || r >= a.rows) {
        if (live) report_oob(a.status, pos);
        r = -1;
      }
      rid[k] = static_cast<int>(r);
    }
    __syncwarp();

    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), qa = sa, sb = sa, qb = sa;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float first_a = 0.f, first_b = 0.f;
    const int* rid_a = rid + g * n_fields;
    const int* rid_b = rid + (g + 8) * n_fields;

    for (int n0 = 0; n0 < n_fields; n0 += kUF) {
      int ra[kUF], rb[kUF];
#pragma unroll
      for (int u = 0; u < kUF; ++u) {
        const bool in = n0 + u < n_fields;
        ra[u] = in ? rid_a[n0 + u] : -1;
        rb[u] = in ? rid_b[n0 + u] : -1;
      }
      float4 xa[kUF], xb[kUF];
#pragma unroll
      for (int u = 0; u < kUF; ++u) {
        xa[u] = ra[u] >= 0 ? ldg_stream_f4(emb4 + (int64_t)ra[u] * 4 + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        xb[u] = rb[u] >= 0 ? ldg_stream_f4(emb4 + (int64_t)rb[u] * 4 + t) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // first-order weights: lane t takes field n0 + t of its two samples
      const int rfa = t == 0 ? ra[0] : (t == 1 ? ra[1] : (t == 2 ? ra[2] : ra[3]));
      const int rfb = t == 0 ? rb[0] : (t == 1 ? rb[1] : (t == 2 ? rb[2] : rb[3]));
      const float fa = rfa >= 0 ? ldg_stream_f1(a.w_feat + rfa) : 0.f;
      const float fb = rfb >= 0 ? ldg_stream_f1(a.w_feat + rfb) : 0.f;

#pragma unroll
      for (int u = 0; u < kUF; ++u) {
        const int n = n0 + u;
        if (n < n_fields) {
          const float4 va = xa[u], vb = xb[u];
          if (a.x_out != nullptr) {   // lanes t = 0..3 write the 64 bytes of a row together
            if (b0 + g < a.batch)
              *reinterpret_cast<float4*>(a.x_out + ((b0 + g) * n_fields + n) * 16 + 4 * t) = va;
            if (b0 + g + 8 < a.batch)
              *reinterpret_cast<float4*>(a.x_out + ((b0 + g + 8) * n_fields + n) * 16 + 4 * t) = vb;
          }
          sa.x += va.x; sa.y += va.y; sa.z += va.z; sa.w += va.w;
          qa.x = fmaf(va.x, va.x, qa.x); qa.y = fmaf(va.y, va.y, qa.y);
          qa.z = fmaf(va.z, va.z, qa.z); qa.w = fmaf(va.w, va.w, qa.w);
          sb.x += vb.x; sb.y += vb.y; sb.z += vb.z; sb.w += vb.w;
          qb.x = fmaf(vb.x, vb.x, qb.x); qb.y = fmaf(vb.y, vb.y, qb.y);
          qb.z = fmaf(vb.z, vb.z, qb.z); qb.w = fmaf(vb.w, vb.w, qb.w);
          // A fragments of the two k-steps of this field: {row g col t, row g+8 col t, row g col t+4, row g+8 col t+4}
          uint32_t ah0[4], al0[4], ah1[4], al1[4];
          split_tf32(va.x, ah0[0], al0[0]); split_tf32(vb.x, ah0[1], al0[1]);
          split_tf32(va.y, ah0[2], al0[2]); split_tf32(vb.y, ah0[3], al0[3]);
          split_tf32(va.z, ah1[0], al1[0]); split_tf32(vb.z, ah1[1], al1[1]);
          split_tf32(va.w, ah1[2], al1[2]); split_tf32(vb.w, ah1[3], al1[3]);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float4 wh = w1s[((n * 2 + j) * 2 + 0) * 32 + lane];
            const float4 wl = w1s[((n * 2 + j) * 2 + 1) * 32 + lane];
            mma_3xtf32(acc[j], ah0, al0, __float_as_uint(wh.x), __float_as_uint(wh.y), __float_as_uint(wl.x),
                       __float_as_uint(wl.y));
            mma_3xtf32(acc[j], ah1, al1, __float_as_uint(wh.z), __float_as_uint(wh.w), __float_as_uint(wl.z),
                       __float_as_uint(wl.w));
          }
        }
      }
      first_a += fa;
      first_b += fb;
    }

    // ---- FM second order + first order, reduced over the 4 lanes of a row ------------------------------------
    float fm_a = 0.5f * ((sa.x * sa.x - qa.x) + (sa.y * sa.y - qa.y) + (sa.z * sa.z - qa.z) + (sa.w * sa.w - qa.w));
    float fm_b = 0.5f * ((sb.x * sb.x - qb.x) + (sb.y * sb.y - qb.y) + (sb.z * sb.z - qb.z) + (sb.w * sb.w - qb.w));
    if (!a.use_fm) fm_a = fm_b = 0.f;
    float side_a = fm_a + first_a, side_b = fm_b + first_b;

    // ---- layer 1 epilogue: bias + ReLU.  acc[j] = {(g, 8j+2t), (g, 8j+2t+1), (g+8, 8j+2t), (g+8, 8j+2t+1)} ----
    float h[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float b0v = bias_s[8 * j + 2 * t], b1v = bias_s[8 * j + 2 * t + 1];
      h[j][0] = fmaxf(acc[j][0] + b0v, 0.f);
      h[j][1] = fmaxf(acc[j][1] + b1v, 0.f);
      h[j][2] = fmaxf(acc[j][2] + b0v, 0.f);
      h[j][3] = fmaxf(acc[j][3] + b1v, 0.f);
    }
    // ---- 16x16 hidden layers on the tensor pipe; k order (2t, 2t+1) matches the accumulator layout -------------
    for (int layer = 0; layer < a.hidden_layers; ++layer) {
      float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int jk = 0; jk < 2; ++jk) {
        uint32_t ah[4], al[4];
        split_tf32(h[jk][0], ah[0], al[0]);  // row g,   k = 8jk + 2t
        split_tf32(h[jk][2], ah[1], al[1]);  // row g+8, k = 8jk + 2t
        split_tf32(h[jk][1], ah[2], al[2]);  // row g,   k = 8jk + 2t + 1
        split_tf32(h[jk][3], ah[3], al[3]);  // row g+8, k = 8jk + 2t + 1
#pragma unroll
        for (int jn = 0; jn < 2; ++jn) {
          const float2 wh = whs[(((layer * 2 + jn) * 2 + jk) * 2 + 0) * 32 + lane];
          const float2 wl = whs[(((layer * 2 + jn) * 2 + jk) * 2 + 1) * 32 + lane];
          mma_3xtf32(o[jn], ah, al, __float_as_uint(wh.x), __float_as_uint(wh.y), __float_as_uint(wl.x),
                     __float_as_uint(wl.y));
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float b0v = bias_s[(1 + layer) * 16 + 8 * j + 2 * t], b1v = bias_s[(1 + layer) * 16 + 8 * j + 2 * t + 1];
        h[j][0] = fmaxf(o[j][0] + b0v, 0.f);
        h[j][1] = fmaxf(o[j][1] + b1v, 0.f);
        h[j][2] = fmaxf(o[j][2] + b0v, 0.f);
        h[j][3] = fmaxf(o[j][3] + b1v, 0.f);
      }
    }
    // ---- output layer (16 -> 1) + reduction over t ---------------------------------------------------------------
    const float* wo = bias_s + (1 + kMaxHidden) * 16;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float w0 = wo[8 * j + 2 * t], w1v = wo[8 * j + 2 * t + 1];
      side_a = fmaf(h[j][0], w0, side_a);
      side_a = fmaf(h[j][1], w1v, side_a);
      side_b = fmaf(h[j][2], w0, side_b);
      side_b = fmaf(h[j][3], w1v, side_b);
    }
    side_a += __shfl_xor_sync(0xffffffffu, side_a, 1);
    side_b += __shfl_xor_sync(0xffffffffu, side_b, 1);
    side_a += __shfl_xor_sync(0xffffffffu, side_a, 2);
    side_b += __shfl_xor_sync(0xffffffffu, side_b, 2);
    if (t == 0) {
      const float bo = wo[16] + (a.bias != nullptr ? __ldg(a.bias) : 0.f);
      if (b0 + g < a.batch) a.logits[b0 + g] = side_a + bo;
      if (b0 + g + 8 < a.batch) a.logits[b0 + g + 8] = side_b + bo;
    }
    __syncwarp();
  }
}

size_t fast_smem_bytes(int fields, int hidden_layers) {
  return (size_t)fields * 4 * 32 * sizeof(float4) + (size_t)hidden_layers * 8 * 32 * sizeof(float2) +
         ((1 + kMaxHidden) * 16 + 16 + 4) * sizeof(float) + (size_t)kWarps * kTile * fields * sizeof(int);
}

}  // namespace

// Shape gate of the fast path: embed 16, every hidden width 16, output 1, ReLU, row ids fit int32.
int deepfm_fast_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation, int64_t rows) {
  static const bool disabled = getenv("TRS_DISABLE_FAST") != nullptr;
  if (disabled) return 0;
  if (embed != 16 || activation != TRS_ACT_RELU || rows >= (int64_t(1) << 31)) return 0;
  if (mlp_layers < 2 || mlp_layers - 2 > kMaxHidden) return 0;
  if (mlp_dims[0] != fields * 16 || mlp_dims[mlp_layers] != 1) return 0;
  for (int l = 1; l < mlp_layers; ++l)
    if (mlp_dims[l] != 16) return 0;
  if (fast_smem_bytes(fields, mlp_layers - 2) > (size_t)kMaxDynSmem / 2) return 0;
  return 1;
}

int deepfm_fast_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                       const float* w_feat, const float* w_emb, int64_t rows, const float* const* mlp_w,
                       const float* const* mlp_b, int mlp_layers, float* logits, int32_t* status, cudaStream_t s,
                       float* x_out, const float* bias, int use_fm) {
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_deepfm_forward: idx_bits must be 32 or 64");
  TRS_REQUIRE(aligned16(w_emb) && aligned16(mlp_w[0]), "trs_deepfm_forward: w_emb / W1 must be 16-byte aligned");
  FastArgs a{};
  a.idx = idx; a.offsets = offsets; a.w_feat = w_feat; a.w_emb = w_emb; a.logits = logits; a.status = status;
  a.batch = batch; a.rows = rows; a.fields = fields;
  a.x_out = x_out; a.bias = bias; a.use_fm = use_fm;
  a.hidden_layers = mlp_layers - 2;
  a.w1 = mlp_w[0];
  a.b1 = mlp_b[0];
  for (int l = 0; l < a.hidden_layers; ++l) {
    a.wh[l] = mlp_w[1 + l];
    a.bh[l] = mlp_b[1 + l];
  }
  a.w_out = mlp_w[mlp_layers - 1];
  a.b_out = mlp_b[mlp_layers - 1];
  for (int l = 0; l < mlp_layers; ++l) TRS_REQUIRE(mlp_w[l] && mlp_b[l], "trs_deepfm_forward: null MLP parameter");
  const size_t smem = fast_smem_bytes(fields, a.hidden_layers);
  const int64_t tiles = (batch + kTile - 1) / kTile;
  const int64_t ctas = (tiles + kWarps - 1) / kWarps;
  const int grid = static_cast<int>(ctas < 2 * kNumSMs ? ctas : 2 * kNumSMs);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN(deepfm_fast_kernel<64>);
    deepfm_fast_kernel<64><<<grid, kWarps * 32, smem, s>>>(a);
  } else {
    TRS_SMEM_OPT_IN(deepfm_fast_kernel<32>);
    deepfm_fast_kernel<32><<<grid, kWarps * 32, smem, s>>>(a);
  }
  return check_launch("deepfm_fast_kernel");
}

}  // namespace trs
