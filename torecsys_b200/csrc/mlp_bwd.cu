// Backward of MultilayerPerceptionLayer.forward (torecsys/layers/ctr/multilayer_perceptron.py:63-84, eval mode) for the
// NARROW MLPs of the CTR models on this path -- DeepFM / xDeepFM / NFM / FNN's deep branch (in = fields x embed, hidden
// widths 16 .. 32), DCN's per-field MLP: every width behind the first Linear <= 32.
//     h_0 = x;  z_l = h_{l-1} W_l^T + b_l;  h_l = act(z_l) for the hidden Linears, h_L = z_L (LinearOutput)
//     d_L = grad_out;  d_{l-1} = (d_l W_l) * act'(z_{l-1});  grad_W_l = d_l^T h_{l-1};  grad_b_l = sum_rows d_l;  grad_x = d_1 W_1
// One kernel: persistent CTAs walk tiles of kRows rows.  Per tile the rows of x are staged in shared memory (coalesced),
// the forward is recomputed (all activations of the tile stay in shared memory), the layers are walked backwards in
// FP32 FFMA, grad_x leaves as full rows.  The parameter gradients are accumulated IN SHARED MEMORY over all tiles of the
// CTA (grad_W_1 is (d_1, in): 40 KB at 624 x 16) with one owner thread per element -- no atomics in the loop -- and
// added to global memory once per CTA.  No library GEMM.
#include "common.cuh"
#include "tile_ops.cuh"

namespace trs {
namespace {

constexpr int kRows = 32;       // rows per tile
constexpr int kThreads = 256;
constexpr int kMaxW = 32;       // widest layer behind the first Linear
constexpr int kHp = kMaxW + 1;  // pitch of the small per-layer tiles

struct MlpBwdArgs {
  const float* x;
  const float* grad_out;
  float* grad_x;
  float* grad_w[MlpParams::kMaxLayers];
  float* grad_b[MlpParams::kMaxLayers];
  int64_t rows;
  int xp;            // pitch of the x tile / W_1 / grad_W_1 rows in shared memory (floats, 4 mod 32)
  MlpParams mp;
};

__device__ __forceinline__ float act_grad(float h, float z, int act) {   // act'(z) from the stored h = act(z)
  switch (act) {
    case TRS_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case TRS_ACT_SIGMOID: return h * (1.f - h);
    case TRS_ACT_TANH: return 1.f - h * h;
    default: return 1.f;
  }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_backward_kernel(const MlpBwdArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.mp.layers, d0 = a.mp.dims[0], d1 = a.mp.dims[1], xp = a.xp;
  float* x_s = sm;                                   // [kRows][xp]       rows of x, later rows of grad_x
  float* w1_s = x_s + kRows * xp;                    // [d1][xp]
  float* gw1_s = w1_s + d1 * xp;                     // [d1][xp]          grad_W_1 of this CTA
  float* h_s = gw1_s + d1 * xp;                      // [L + 1][kRows][kHp]   h_1 .. h_L (index l), z kept for ReLU via h > 0
  float* d_s = h_s + (L + 1) * kRows * kHp;          // [2][kRows][kHp]   d_l ping-pong
  float* wl_s = d_s + 2 * kRows * kHp;               // [L][kMaxW][kHp]   W_2 .. W_L  (index l - 1), row o = outputs
  float* gwl_s = wl_s + L * kMaxW * kHp;             // [L][kMaxW][kHp]   grad_W_2 .. grad_W_L
  float* b_s = gwl_s + L * kMaxW * kHp;              // [L][kMaxW]
  float* gb_s = b_s + L * kMaxW;                     // [L][kMaxW]
  const int tid = threadIdx.x;

  for (int i = tid; i < d1 * xp; i += kThreads) {
    const int o = i / xp, k = i - o * xp;
    w1_s[i] = k < d0 ? __ldg(a.mp.w[0] + (size_t)o * d0 + k) : 0.f;
    gw1_s[i] = 0.f;
  }
  for (int l = 1; l < L; ++l) {
    const int din = a.mp.dims[l], dout = a.mp.dims[l + 1];
    for (int i = tid; i < kMaxW * kHp; i += kThreads) {
      const int o = i / kHp, k = i - o * kHp;
      wl_s[l * kMaxW * kHp + i] = (o < dout && k < din) ? __ldg(a.mp.w[l] + (size_t)o * din + k) : 0.f;
      gwl_s[l * kMaxW * kHp + i] = 0.f;
    }
  }
  for (int i = tid; i < L * kMaxW; i += kThreads) {
    const int l = i / kMaxW, o = i - l * kMaxW;
    b_s[i] = (o < a.mp.dims[l + 1] && a.mp.b[l] != nullptr) ? __ldg(a.mp.b[l] + o) : 0.f;
    gb_s[i] = 0.f;
  }
  __syncthreads();

  const int64_t tiles = (a.rows + kRows - 1) / kRows;
  const int d0v = d0 >> 2;   // float4 per row (d0 % 4 == 0)
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kRows;
    const int valid = static_cast<int>(a.rows - r0 < kRows ? a.rows - r0 : kRows);
    // ---- stage x (zero rows beyond the batch)
    for (int i = tid; i < kRows * d0v; i += kThreads) {
      const int r = i / d0v, c = i - r * d0v;
      const float4 v = r < valid ? __ldg(reinterpret_cast<const float4*>(a.x + (r0 + r) * d0) + c)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(x_s + r * xp + 4 * c) = v;
    }
    __syncthreads();
    // ---- forward, layer 1: thread = (row, two outputs o, o + 16) ; 8 lanes of a row group share the x row (broadcast)
    {
      const int r = tid >> 3, og = tid & 7;   // 32 rows x 8 lanes; lane og owns outputs og, og + 8, og + 16, og + 24
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* xr = x_s + r * xp;
      for (int k = 0; k < d0; k += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = og + 8 * j;
          if (o < d1) {
            const float4 wv = *reinterpret_cast<const float4*>(w1_s + o * xp + k);
            acc[j] = fmaf(xv.x, wv.x, acc[j]); acc[j] = fmaf(xv.y, wv.y, acc[j]);
            acc[j] = fmaf(xv.z, wv.z, acc[j]); acc[j] = fmaf(xv.w, wv.w, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = og + 8 * j;
        if (o < d1) {
          const float z = acc[j] + b_s[o];
          h_s[(1 * kRows + r) * kHp + o] = L > 1 ? apply_act(z, a.mp.act) : z;
        }
      }
    }
    __syncthreads();
    // ---- forward, layers 2 .. L: thread = (row, output) with the 8 lanes striding the outputs
    for (int l = 1; l < L; ++l) {
      const int din = a.mp.dims[l], dout = a.mp.dims[l + 1];
      const int r = tid >> 3, og = tid & 7;
      for (int o = og; o < dout; o += 8) {
        float z = b_s[l * kMaxW + o];
        const float* hr = h_s + (l * kRows + r) * kHp;
        const float* wr = wl_s + (l * kMaxW + o) * kHp;
        for (int k = 0; k < din; ++k) z = fmaf(hr[k], wr[k], z);
        h_s[((l + 1) * kRows + r) * kHp + o] = l < L - 1 ? apply_act(z, a.mp.act) : z;
      }
      __syncthreads();
    }
    // ---- backward: d_L = grad_out
    {
      const int dl = a.mp.dims[L];
      for (int i = tid; i < kRows * dl; i += kThreads) {
        const int r = i / dl, o = i - r * dl;
        d_s[(0 * kRows + r) * kHp + o] = r < valid ? __ldg(a.grad_out + (r0 + r) * dl + o) : 0.f;
      }
    }
    __syncthreads();
    int cur = 0;
    for (int l = L - 1; l >= 1; --l) {   // layer l + 1 in 1-based terms: W = wl_s[l], input h_l, output width dims[l + 1]
      const int din = a.mp.dims[l], dout = a.mp.dims[l + 1];
      const float* dl_s = d_s + cur * kRows * kHp;
      // grad_W, grad_b: one owner thread per (o, k)
      for (int i = tid; i < dout * din; i += kThreads) {
        const int o = i / din, k = i - o * din;
        float g = 0.f;
        for (int r = 0; r < kRows; ++r) g = fmaf(dl_s[r * kHp + o], h_s[(l * kRows + r) * kHp + k], g);
        gwl_s[(l * kMaxW + o) * kHp + k] += g;
      }
      for (int o = tid; o < dout; o += kThreads) {
        float g = 0.f;
        for (int r = 0; r < kRows; ++r) g += dl_s[r * kHp + o];
        gb_s[l * kMaxW + o] += g;
      }
      // d_{l} = (d_{l+1} W) * act'(h_l)
      float* dn_s = d_s + (cur ^ 1) * kRows * kHp;
      for (int i = tid; i < kRows * din; i += kThreads) {
        const int r = i / din, k = i - r * din;
        float g = 0.f;
        for (int o = 0; o < dout; ++o) g = fmaf(dl_s[r * kHp + o], wl_s[(l * kMaxW + o) * kHp + k], g);
        const float h = h_s[(l * kRows + r) * kHp + k];
        dn_s[r * kHp + k] = g * act_grad(h, h, a.mp.act);   // ReLU: h > 0 <=> z > 0
      }
      cur ^= 1;
      __syncthreads();
    }
    // ---- layer 1: grad_W_1 (owner per element), grad_b_1, grad_x (written over the x tile after grad_W_1 has read it)
    {
      const float* d1_s = d_s + cur * kRows * kHp;
      for (int i = tid; i < d1 * d0v; i += kThreads) {
        const int o = i / d0v, c = i - o * d0v;
        float4 g = *reinterpret_cast<float4*>(gw1_s + o * xp + 4 * c);
        for (int r = 0; r < kRows; ++r) {
          const float dv = d1_s[r * kHp + o];
          const float4 xv = *reinterpret_cast<const float4*>(x_s + r * xp + 4 * c);
          g.x = fmaf(dv, xv.x, g.x); g.y = fmaf(dv, xv.y, g.y); g.z = fmaf(dv, xv.z, g.z); g.w = fmaf(dv, xv.w, g.w);
        }
        *reinterpret_cast<float4*>(gw1_s + o * xp + 4 * c) = g;
      }
      for (int o = tid; o < d1; o += kThreads) {
        float g = 0.f;
        for (int r = 0; r < kRows; ++r) g += d1_s[r * kHp + o];
        gb_s[o] += g;
      }
      __syncthreads();
      if (a.grad_x != nullptr) {
        for (int i = tid; i < kRows * d0v; i += kThreads) {
          const int r = i / d0v, c = i - r * d0v;
          float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int o = 0; o < d1; ++o) {
            const float dv = d1_s[r * kHp + o];
            const float4 wv = *reinterpret_cast<const float4*>(w1_s + o * xp + 4 * c);
            g.x = fmaf(dv, wv.x, g.x); g.y = fmaf(dv, wv.y, g.y); g.z = fmaf(dv, wv.z, g.z); g.w = fmaf(dv, wv.w, g.w);
          }
          if (r < valid) *reinterpret_cast<float4*>(a.grad_x + (r0 + r) * d0 + 4 * c) = g;
        }
      }
    }
    __syncthreads();
  }
  // ---- this CTA's parameter gradients -> global memory (one float atomic per element and CTA)
  for (int i = tid; i < d1 * d0; i += kThreads) {
    const int o = i / d0, k = i - o * d0;
    atomicAdd(a.grad_w[0] + i, gw1_s[o * xp + k]);
  }
  for (int l = 1; l < L; ++l) {
    const int din = a.mp.dims[l], dout = a.mp.dims[l + 1];
    for (int i = tid; i < dout * din; i += kThreads) {
      const int o = i / din, k = i - o * din;
      atomicAdd(a.grad_w[l] + i, gwl_s[(l * kMaxW + o) * kHp + k]);
    }
  }
  for (int l = 0; l < L; ++l)
    if (a.grad_b[l] != nullptr)
      for (int o = tid; o < a.mp.dims[l + 1]; o += kThreads) atomicAdd(a.grad_b[l] + o, gb_s[l * kMaxW + o]);
}

size_t smem_for(const MlpParams& mp, int xp) {
  const int L = mp.layers;
  size_t f = (size_t)kRows * xp + 2 * (size_t)mp.dims[1] * xp + (size_t)(L + 1) * kRows * kHp + 2 * kRows * kHp +
             2 * (size_t)L * kMaxW * kHp + 2 * (size_t)L * kMaxW;
  return f * sizeof(float);
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_mlp_backward_supported(const int* dims, int layers) {
  if (dims == nullptr || layers < 1 || layers > MlpParams::kMaxLayers) return 0;
  if (dims[0] < 4 || dims[0] % 4 != 0) return 0;
  for (int l = 1; l <= layers; ++l)
    if (dims[l] < 1 || dims[l] > kMaxW) return 0;
  MlpParams mp{};
  mp.layers = layers;
  for (int l = 0; l <= layers; ++l) mp.dims[l] = dims[l];
  const int xp = ((dims[0] + 31) / 32) * 32 + 4;
  return smem_for(mp, xp) <= (size_t)kMaxDynSmem ? 1 : 0;
}

extern "C" int trs_mlp_backward(const float* x, int64_t rows, const int* dims, int layers,
                                const float* const* weights, const float* const* biases, int activation,
                                const float* grad_out, float* grad_x, float* const* grad_weights,
                                float* const* grad_biases, void* stream) {
  TRS_REQUIRE(x && dims && weights && grad_out && grad_weights, "trs_mlp_backward: null pointer");
  TRS_REQUIRE(rows >= 0, "trs_mlp_backward: bad sizes");
  TRS_UNSUPPORTED(!trs_mlp_backward_supported(dims, layers),
                  "trs_mlp_backward: needs in %% 4 == 0, every other width <= %d and the tile in shared memory", kMaxW);
  TRS_UNSUPPORTED(!aligned16(x) || (grad_x != nullptr && !aligned16(grad_x)), "trs_mlp_backward: x / grad_x must be 16-byte aligned");
  MlpBwdArgs a{};
  TRS_REQUIRE(fill_mlp_params(a.mp, dims, layers, weights, biases, activation) == 0, "trs_mlp_backward: bad MLP description");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int l = 0; l < layers; ++l) {
    TRS_REQUIRE(grad_weights[l] != nullptr, "trs_mlp_backward: null weight gradient");
    a.grad_w[l] = grad_weights[l];
    a.grad_b[l] = grad_biases != nullptr ? grad_biases[l] : nullptr;
    TRS_CUDA(cudaMemsetAsync(a.grad_w[l], 0, (size_t)dims[l + 1] * dims[l] * sizeof(float), s));
    if (a.grad_b[l] != nullptr) TRS_CUDA(cudaMemsetAsync(a.grad_b[l], 0, (size_t)dims[l + 1] * sizeof(float), s));
  }
  if (rows == 0) return TRS_OK;
  a.x = x; a.grad_out = grad_out; a.grad_x = grad_x; a.rows = rows;
  a.xp = ((dims[0] + 31) / 32) * 32 + 4;
  const size_t smem = smem_for(a.mp, a.xp);
  const int64_t tiles = (rows + kRows - 1) / kRows;
  const int grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  TRS_SMEM_OPT_IN(mlp_backward_kernel);
  mlp_backward_kernel<<<grid, kThreads, smem, s>>>(a);
  return check_launch("mlp_backward_kernel");
}

// ---- backward of ComposeExcitationNetworkLayer.forward (compose_excitation_network.py:72-109), SENET of FiBiNET ---------
//     p[m] = mean_e x[m,e];  u = act(W1 p + b1) (R);  s = act(W2 u + b2) (M);  out[m,e] = x[m,e] * s[m]
//     ds[m] = sum_e g[m,e] x[m,e];  dz2 = ds * act'(s);  du = W2^T dz2;  dz1 = du * act'(u);  dp = W1^T dz1
//     grad_x[m,e] = g[m,e] * s[m] + dp[m] / E;  grad_W2 += dz2 u^T, grad_b2 += dz2, grad_W1 += dz1 p^T, grad_b1 += dz1
// One warp per sample (M <= 64 fields, R <= 32): lane = field (two per lane), the small vectors through shuffles /
// shared memory; the CTA accumulates the parameter gradients of its samples in shared memory (one owner thread per
// element, fixed order over the CTA's warps) and adds them to global memory once.
namespace trs {
namespace {

constexpr int kSeWarps = 8, kSeMaxM = 64, kSeMaxR = 32;

__global__ void __launch_bounds__(kSeWarps * 32) senet_backward_kernel(
    const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ w2, const float* __restrict__ b2, int act, const float* __restrict__ g, int64_t batch,
    int m_dim, int embed, int r_dim, float* __restrict__ gx, float* __restrict__ gw1, float* __restrict__ gb1,
    float* __restrict__ gw2, float* __restrict__ gb2) {
  __shared__ float w1_s[kSeMaxR][kSeMaxM + 1], w2_s[kSeMaxM][kSeMaxR + 1], b1_s[kSeMaxR], b2_s[kSeMaxM];
  __shared__ float gw1_s[kSeMaxR][kSeMaxM + 1], gw2_s[kSeMaxM][kSeMaxR + 1], gb1_s[kSeMaxR], gb2_s[kSeMaxM];
  __shared__ float p_s[kSeWarps][kSeMaxM], u_s[kSeWarps][kSeMaxR], dz1_s[kSeWarps][kSeMaxR], dz2_s[kSeWarps][kSeMaxM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  for (int i = tid; i < kSeMaxR * (kSeMaxM + 1); i += blockDim.x) {
    const int r = i / (kSeMaxM + 1), m = i - r * (kSeMaxM + 1);
    w1_s[r][m] = (r < r_dim && m < m_dim) ? __ldg(w1 + r * m_dim + m) : 0.f;
    gw1_s[r][m] = 0.f;
  }
  for (int i = tid; i < kSeMaxM * (kSeMaxR + 1); i += blockDim.x) {
    const int m = i / (kSeMaxR + 1), r = i - m * (kSeMaxR + 1);
    w2_s[m][r] = (m < m_dim && r < r_dim) ? __ldg(w2 + m * r_dim + r) : 0.f;
    gw2_s[m][r] = 0.f;
  }
  for (int i = tid; i < kSeMaxR; i += blockDim.x) { b1_s[i] = i < r_dim ? __ldg(b1 + i) : 0.f; gb1_s[i] = 0.f; }
  for (int i = tid; i < kSeMaxM; i += blockDim.x) { b2_s[i] = i < m_dim ? __ldg(b2 + i) : 0.f; gb2_s[i] = 0.f; }
  __syncthreads();
  const float inv_e = 1.f / static_cast<float>(embed);
  const int64_t rounds = (batch + (int64_t)gridDim.x * kSeWarps - 1) / ((int64_t)gridDim.x * kSeWarps);
  for (int64_t it = 0; it < rounds; ++it) {
    const int64_t b = (it * gridDim.x + blockIdx.x) * kSeWarps + warp;
    const bool live = b < batch;
    // p, ds per field (lane owns fields lane and lane + 32)
    float pm[2] = {0.f, 0.f}, ds[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      if (live && m < m_dim) {
        const float* xr = x + (b * m_dim + m) * embed;
        const float* gr = g + (b * m_dim + m) * embed;
        float sx = 0.f, sg = 0.f;
        for (int e = 0; e < embed; ++e) {
          const float xv = __ldg(xr + e);
          sx += xv;
          sg = fmaf(__ldg(gr + e), xv, sg);
        }
        pm[h] = sx * inv_e;
        ds[h] = sg;
      }
      p_s[warp][m] = pm[h];
    }
    __syncwarp();
    // u[r] (lane = r)
    float z1 = b1_s[lane];
    for (int m = 0; m < m_dim; ++m) z1 = fmaf(w1_s[lane][m], p_s[warp][m], z1);
    const float u = lane < r_dim ? apply_act(z1, act) : 0.f;
    u_s[warp][lane] = u;
    __syncwarp();
    // s[m], dz2[m]
    float sv[2], dz2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      float z2 = b2_s[m];
      for (int r = 0; r < r_dim; ++r) z2 = fmaf(w2_s[m][r], u_s[warp][r], z2);
      sv[h] = m < m_dim ? apply_act(z2, act) : 0.f;
      dz2[h] = (live && m < m_dim) ? ds[h] * act_grad(sv[h], z2, act) : 0.f;
      dz2_s[warp][m] = dz2[h];
    }
    __syncwarp();
    // du[r], dz1[r]
    float du = 0.f;
    for (int m = 0; m < m_dim; ++m) du = fmaf(w2_s[m][lane], dz2_s[warp][m], du);
    const float dz1 = lane < r_dim ? du * act_grad(u, z1, act) : 0.f;
    dz1_s[warp][lane] = dz1;
    __syncwarp();
    // dp[m], grad_x
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      if (live && m < m_dim) {
        float dp = 0.f;
        for (int r = 0; r < r_dim; ++r) dp = fmaf(w1_s[r][m], dz1_s[warp][r], dp);
        dp *= inv_e;
        const float* gr = g + (b * m_dim + m) * embed;
        float* o = gx + (b * m_dim + m) * embed;
        for (int e = 0; e < embed; ++e) o[e] = fmaf(__ldg(gr + e), sv[h], dp);
      }
    }
    __syncthreads();
    // parameter gradients of this round's kSeWarps samples: one owner thread per element, warps in fixed order
    for (int i = tid; i < r_dim * m_dim; i += blockDim.x) {
      const int r = i / m_dim, m = i - r * m_dim;
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int w = 0; w < kSeWarps; ++w) {
        a1 = fmaf(dz1_s[w][r], p_s[w][m], a1);
        a2 = fmaf(dz2_s[w][m], u_s[w][r], a2);
      }
      gw1_s[r][m] += a1;
      gw2_s[m][r] += a2;
    }
    if (tid < r_dim) {
      float a1 = 0.f;
#pragma unroll
      for (int w = 0; w < kSeWarps; ++w) a1 += dz1_s[w][tid];
      gb1_s[tid] += a1;
    }
    if (tid >= 64 && tid - 64 < m_dim) {
      float a2 = 0.f;
#pragma unroll
      for (int w = 0; w < kSeWarps; ++w) a2 += dz2_s[w][tid - 64];
      gb2_s[tid - 64] += a2;
    }
    __syncthreads();
  }
  for (int i = tid; i < r_dim * m_dim; i += blockDim.x) {
    const int r = i / m_dim, m = i - r * m_dim;
    atomicAdd(gw1 + i, gw1_s[r][m]);
    atomicAdd(gw2 + m * r_dim + r, gw2_s[m][r]);
  }
  if (tid < r_dim) atomicAdd(gb1 + tid, gb1_s[tid]);
  if (tid >= 64 && tid - 64 < m_dim) atomicAdd(gb2 + tid - 64, gb2_s[tid - 64]);
}

}  // namespace
}  // namespace trs

extern "C" int trs_senet_backward_supported(int rows_per_sample, int reduced) {
  return rows_per_sample >= 1 && rows_per_sample <= kSeMaxM && reduced >= 1 && reduced <= kSeMaxR ? 1 : 0;
}

extern "C" int trs_senet_backward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                                  int activation, const float* grad_out, int64_t batch, int rows_per_sample, int embed,
                                  int reduced, float* grad_x, float* grad_w1, float* grad_b1, float* grad_w2,
                                  float* grad_b2, void* stream) {
  TRS_REQUIRE(x && w1 && b1 && w2 && b2 && grad_out && grad_x && grad_w1 && grad_b1 && grad_w2 && grad_b2,
              "trs_senet_backward: null pointer");
  TRS_REQUIRE(batch >= 0 && embed > 0, "trs_senet_backward: bad sizes");
  TRS_REQUIRE(activation >= TRS_ACT_NONE && activation <= TRS_ACT_TANH, "trs_senet_backward: unknown activation");
  TRS_UNSUPPORTED(!trs_senet_backward_supported(rows_per_sample, reduced),
                  "trs_senet_backward: at most %d rows per sample and %d reduced units", kSeMaxM, kSeMaxR);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TRS_CUDA(cudaMemsetAsync(grad_w1, 0, (size_t)reduced * rows_per_sample * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_w2, 0, (size_t)reduced * rows_per_sample * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_b1, 0, (size_t)reduced * sizeof(float), s));
  TRS_CUDA(cudaMemsetAsync(grad_b2, 0, (size_t)rows_per_sample * sizeof(float), s));
  if (batch == 0) return TRS_OK;
  const int64_t want = (batch + kSeWarps - 1) / kSeWarps;
  const int grid = static_cast<int>(want < kNumSMs * 2 ? want : kNumSMs * 2);
  senet_backward_kernel<<<grid, kSeWarps * 32, 0, s>>>(x, w1, b1, w2, b2, activation, grad_out, batch, rows_per_sample,
                                                     embed, reduced, grad_x, grad_w1, grad_b1, grad_w2, grad_b2);
  return check_launch("senet_backward_kernel");
}
