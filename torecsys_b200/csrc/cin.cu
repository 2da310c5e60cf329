// a8: Compress Interaction Network, generic fp32 FFMA path (any N, H_l, E % 4 == 0).
//
// GEMM view per layer: rows m = (b, e), K = N * H (k = xf*H + y, x-major),
// columns = output channels;  A[m][k] = x0[b,xf,e] * h[b,y,e] is never materialised: each thread rebuilds its
// 4-row A fragment from two 128-bit shared-memory loads (4 consecutive e of x0[b,xf,:] and of h[b,y,:]).
// Epilogue fuses Conv1d bias + eval BatchNorm (folded to scale/shift by the caller) + activation, the
// direct/hidden split and the sum over e of the direct half.  The dead hidden half of the LAST layer
// (computed and discarded by the reference, compress_interaction_network.py:151-156) is not computed.
#include "common.cuh"

namespace trs {
namespace {

constexpr int BM = 64;   // rows (b,e) per CTA, as 16 groups of 4 consecutive e
constexpr int BN = 64;   // channels per CTA
constexpr int BK = 32;   // k values staged per step
constexpr int WP = BN + 4;

struct CinLayerArgs {
  const float* x0;      // (B, N, E)
  const float* h;       // (B, H, E)
  const float* w;       // (C, N*H)
  const float* scale;   // (C)
  const float* shift;   // (C)
  float* h_next;        // (B, hid_count, E) or null
  float* pooled;        // (B, pooled_width)
  int64_t m_rows;       // B * E
  int n_fields, h_prev, embed;
  int c_count;          // channels to compute
  int n_direct;         // channels [0, n_direct) are summed over e into pooled[:, pool_off + c]
  int hid_begin, hid_count;
  int pool_off, pooled_width;
  int act;
};

__global__ void __launch_bounds__(256) cin_layer_kernel(CinLayerArgs a) {
  extern __shared__ __align__(16) float smem[];
  float4* xs = reinterpret_cast<float4*>(smem);           // [16][N]
  float4* hs = xs + 16 * a.n_fields;                      // [16][H]
  float* ws = reinterpret_cast<float*>(hs + 16 * a.h_prev);  // [BK][WP]
  float* red = ws + BK * WP;                              // [16][BN]

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int c0 = blockIdx.y * BN;
  const int e_dim = a.embed;
  const int kdim = a.n_fields * a.h_prev;

  // stage the x0 / h fragments of the 16 row groups
  for (int t = tid; t < 16 * a.n_fields; t += blockDim.x) {
    const int g = t / a.n_fields, xf = t - g * a.n_fields;
    const int64_t m = m0 + 4 * g;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < a.m_rows) {
      const int64_t b = m / e_dim;
      const int e = static_cast<int>(m - b * e_dim);
      v = __ldg(reinterpret_cast<const float4*>(a.x0 + (b * a.n_fields + xf) * e_dim + e));
    }
    xs[g * a.n_fields + xf] = v;
  }
  for (int t = tid; t < 16 * a.h_prev; t += blockDim.x) {
    const int g = t / a.h_prev, y = t - g * a.h_prev;
    const int64_t m = m0 + 4 * g;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < a.m_rows) {
      const int64_t b = m / e_dim;
      const int e = static_cast<int>(m - b * e_dim);
      v = __ldg(reinterpret_cast<const float4*>(a.h + (b * a.h_prev + y) * e_dim + e));
    }
    hs[g * a.h_prev + y] = v;
  }

  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  const float4* xg = xs + ty * a.n_fields;
  const float4* hg = hs + ty * a.h_prev;

  for (int k0 = 0; k0 < kdim; k0 += BK) {
    __syncthreads();
    // stage W[c0 .. c0+BN)[k0 .. k0+BK) transposed into ws[kk][c]
    for (int t = tid; t < BK * BN; t += blockDim.x) {
      const int kk = t & (BK - 1), c = t >> 5;
      const int k = k0 + kk, ch = c0 + c;
      ws[kk * WP + c] = (k < kdim && ch < a.c_count) ? __ldg(a.w + (int64_t)ch * kdim + k) : 0.f;
    }
    __syncthreads();
    int xf = k0 / a.h_prev;
    int y = k0 - xf * a.h_prev;
    const int kend = min(BK, kdim - k0);
    for (int kk = 0; kk < kend; ++kk) {
      const float4 xv = xg[xf];
      const float4 hv = hg[y];
      const float4 wv = *reinterpret_cast<const float4*>(ws + kk * WP + 4 * tx);
      const float av[4] = {xv.x * hv.x, xv.y * hv.y, xv.z * hv.z, xv.w * hv.w};
      const float bv[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      if (++y == a.h_prev) {
        y = 0;
        ++xf;
      }
    }
  }

  // epilogue
  const int64_t m = m0 + 4 * ty;
  const bool row_ok = m < a.m_rows;
  const int64_t b = row_ok ? m / e_dim : 0;
  const int e = row_ok ? static_cast<int>(m - b * e_dim) : 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int ch = c0 + 4 * tx + c;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch < a.c_count) {
      const float sc = __ldg(a.scale + ch), sh = __ldg(a.shift + ch);
#pragma unroll
      for (int r = 0; r < 4; ++r) o[r] = apply_act(fmaf(acc[r][c], sc, sh), a.act);
      if (row_ok && a.h_next != nullptr && ch >= a.hid_begin && ch < a.hid_begin + a.hid_count) {
        *reinterpret_cast<float4*>(a.h_next + (b * a.hid_count + (ch - a.hid_begin)) * e_dim + e) =
            make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    red[ty * BN + 4 * tx + c] = row_ok ? (o[0] + o[1]) + (o[2] + o[3]) : 0.f;
  }
  __syncthreads();
  // sum the row groups that belong to the same sample (fixed order), one atomic per (sample-in-CTA, channel)
  for (int t = tid; t < 16 * BN; t += blockDim.x) {
    const int g = t / BN, c = t - g * BN;
    const int ch = c0 + c;
    const int64_t mg = m0 + 4 * g;
    if (ch >= a.n_direct || mg >= a.m_rows) continue;
    const int64_t bg = mg / e_dim;
    // only the first group of each sample inside this CTA does the reduction
    if (g > 0 && (m0 + 4 * (g - 1)) / e_dim == bg) continue;
    float tot = 0.f;
    for (int g2 = g; g2 < 16; ++g2) {
      const int64_t m2 = m0 + 4 * g2;
      if (m2 >= a.m_rows || m2 / e_dim != bg) break;
      tot += red[g2 * BN + c];
    }
    atomicAdd(a.pooled + bg * a.pooled_width + a.pool_off + ch, tot);
  }
}

__global__ void __launch_bounds__(256) cin_fc_kernel(const float* __restrict__ pooled, int pooled_width,
                                                     const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                                     int out_features, int64_t batch, float* __restrict__ out,
                                                     int accumulate) {
  const int64_t items = batch * out_features;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < items; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / out_features;
    const int o = static_cast<int>(t - b * out_features);
    float acc = fc_b ? __ldg(fc_b + o) : 0.f;
    const float* p = pooled + b * pooled_width;
    const float* w = fc_w + (int64_t)o * pooled_width;
    for (int k = 0; k < pooled_width; ++k) acc = fmaf(p[k], __ldg(w + k), acc);
    if (accumulate) out[t] += acc;
    else out[t] = acc;
  }
}

}  // namespace

int cin_tc_supported(int fields, int embed, const int* sizes, int layers, int is_direct);
int64_t cin_tc_workspace_bytes(int64_t batch, int fields, int embed, const int* sizes, int layers, int is_direct);
int cin_tc_run(const float* x, const float* const* conv_w, const float* const* scale, const float* const* shift,
               const int* sizes, int layers, int is_direct, int activation, const float* fc_w, const float* fc_b,
               int out_features, int64_t batch, int fields, int embed, float* out, int accumulate, void* workspace,
               int64_t workspace_bytes, cudaStream_t s);

int cin_fc_launch(const float* pooled, int pooled_width, const float* fc_w, const float* fc_b, int out_features,
                  int64_t batch, float* out, int accumulate, cudaStream_t s) {
  const int grid = grid_for(batch * out_features, 256, 8);
  cin_fc_kernel<<<grid, 256, 0, s>>>(pooled, pooled_width, fc_w, fc_b, out_features, batch, out, accumulate);
  return check_launch("cin_fc_kernel");
}

struct CinPlan {
  int64_t pooled_floats, h_floats;
  int pooled_width, h_max;
};

static CinPlan cin_plan(int64_t batch, int embed, const int* layer_sizes, int layers) {
  CinPlan p{};
  for (int l = 0; l < layers; ++l) {
    p.pooled_width += layer_sizes[l];
    if (l < layers - 1 && layer_sizes[l] > p.h_max) p.h_max = layer_sizes[l];
  }
  p.pooled_floats = ((batch * p.pooled_width + 63) / 64) * 64;
  p.h_floats = ((batch * (int64_t)p.h_max * embed + 63) / 64) * 64;
  return p;
}

// shared by trs_cin_forward and trs_xdeepfm_forward; `accumulate` adds the fc output onto `out`
int cin_run(const float* x, const float* const* conv_w, const float* const* scale, const float* const* shift,
            const int* layer_sizes, int layers, int is_direct, int activation, const float* fc_w, const float* fc_b,
            int out_features, int64_t batch, int fields, int embed, float* out, int accumulate, void* workspace,
            int64_t workspace_bytes, cudaStream_t s) {
  TRS_REQUIRE(x && conv_w && scale && shift && layer_sizes && fc_w && out && workspace, "cin: null pointer");
  TRS_REQUIRE(layers >= 1 && batch >= 0 && fields > 0 && embed > 0 && out_features > 0, "cin: bad sizes");
  TRS_UNSUPPORTED(embed % 4 != 0, "cin: embed_size must be a multiple of 4 (got %d)", embed);
  if (batch == 0) return TRS_OK;
  // dense GEMM -> tensor cores (tcgen05 / TMEM, 3xTF32) whenever the shape allows; FP32 FFMA tiles otherwise
  if (cin_tc_supported(fields, embed, layer_sizes, layers, is_direct) && aligned16(x))
    return cin_tc_run(x, conv_w, scale, shift, layer_sizes, layers, is_direct, activation, fc_w, fc_b, out_features,
                      batch, fields, embed, out, accumulate, workspace, workspace_bytes, s);
  const CinPlan plan = cin_plan(batch, embed, layer_sizes, layers);
  const int64_t need = (plan.pooled_floats + 2 * plan.h_floats) * (int64_t)sizeof(float);
  TRS_REQUIRE(workspace_bytes >= need, "cin: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
              (long long)need);
  TRS_REQUIRE(aligned16(workspace) && aligned16(x), "cin: x and workspace must be 16-byte aligned");
  float* pooled = static_cast<float*>(workspace);
  float* hbuf[2] = {pooled + plan.pooled_floats, pooled + plan.pooled_floats + plan.h_floats};
  TRS_CUDA(cudaMemsetAsync(pooled, 0, (size_t)batch * plan.pooled_width * sizeof(float), s));
  TRS_SMEM_OPT_IN(cin_layer_kernel);

  const float* h = x;
  int h_prev = fields;
  int pool_off = 0;
  for (int l = 0; l < layers; ++l) {
    const bool last = (l == layers - 1);
    const int hl = layer_sizes[l];
    TRS_REQUIRE(hl > 0 && conv_w[l] && scale[l] && shift[l], "cin: bad layer %d", l);
    CinLayerArgs a{};
    a.x0 = x;
    a.h = h;
    a.w = conv_w[l];
    a.scale = scale[l];
    a.shift = shift[l];
    a.h_next = last ? nullptr : hbuf[l & 1];
    a.pooled = pooled;
    a.m_rows = batch * embed;
    a.n_fields = fields;
    a.h_prev = h_prev;
    a.embed = embed;
    a.n_direct = hl;
    a.hid_begin = is_direct ? 0 : hl;
    a.hid_count = last ? 0 : hl;
    a.c_count = (is_direct || last) ? hl : 2 * hl;
    a.pool_off = pool_off;
    a.pooled_width = plan.pooled_width;
    a.act = activation;
    const size_t smem = (size_t)16 * (fields + h_prev) * sizeof(float4) + (size_t)(BK * WP + 16 * BN) * sizeof(float);
    TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "cin: num_fields + layer size too large for shared memory");
    dim3 grid((unsigned)((a.m_rows + BM - 1) / BM), (unsigned)((a.c_count + BN - 1) / BN));
    cin_layer_kernel<<<grid, 256, smem, s>>>(a);
    int rc = check_launch("cin_layer_kernel");
    if (rc != TRS_OK) return rc;
    h = a.h_next;
    h_prev = hl;
    pool_off += hl;
  }
  return cin_fc_launch(pooled, plan.pooled_width, fc_w, fc_b, out_features, batch, out, accumulate, s);
}

}  // namespace trs

using namespace trs;

extern "C" int64_t trs_cin_workspace_bytes(int64_t batch, int fields, int embed, const int* layer_sizes, int layers,
                                           int is_direct) {
  if (!layer_sizes || layers < 1 || batch < 0 || embed <= 0) return -1;
  const CinPlan plan = cin_plan(batch, embed, layer_sizes, layers);
  int64_t need = (plan.pooled_floats + 2 * plan.h_floats) * (int64_t)sizeof(float) + 256;
  if (cin_tc_supported(fields, embed, layer_sizes, layers, is_direct)) {
    const int64_t tc = cin_tc_workspace_bytes(batch, fields, embed, layer_sizes, layers, is_direct);
    if (tc > need) need = tc;
  }
  return need;
}

extern "C" int trs_cin_forward(const float* x, const float* const* conv_w, const float* const* scale,
                               const float* const* shift, const int* layer_sizes, int layers, int is_direct,
                               int activation, const float* fc_w, const float* fc_b, int out_features, int64_t batch,
                               int fields, int embed, float* out, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  return cin_run(x, conv_w, scale, shift, layer_sizes, layers, is_direct, activation, fc_w, fc_b, out_features, batch,
                 fields, embed, out, 0, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}
