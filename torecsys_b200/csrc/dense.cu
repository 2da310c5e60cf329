// MLP (DNNLayer, adjacent to the hot path) and a7 Cross network as L1 ops on materialised rows.
//
// Both treat the input as a (rows, width) matrix, stage a tile of rows in shared memory and run the small dense
// layers on it in plain fp32 FFMA (parity bar 1e-5: no TF32/bf16).
#include <stdlib.h>

#include "tile_ops.cuh"

namespace trs {

int cross_tc_launch(const float* x, const float* w, const float* b, int layers, int64_t rows, int embed, float* out,
                    cudaStream_t s);
int cross_tc5_launch(const float* x, const float* w, const float* b, int layers, int64_t rows, int embed, float* out,
                     cudaStream_t s);
int dense_tc_supported(int k_dim, int c_dim, const void* x, const void* out);
int dense_tc_run(const float* x, int64_t rows, int k_dim, const float* w, const float* bias, int c_dim, int activation,
                 float* out, cudaStream_t s, const DenseFuse* fz);
int dense_tc_passes(int c_dim, int sel, int gather);
int dense_dot_finish(const float* dot, int passes, int64_t rows, const float* bias, float* out, int accumulate,
                     cudaStream_t s);

namespace {

// Narrow output layers (the logit layer of every model): out[r, o] (+)= act(b[o] + <x[r, :], w[o, :]>), one warp per row.
__global__ void __launch_bounds__(256) narrow_dense_kernel(const float* __restrict__ x, int64_t rows, int k,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           int c, int act, float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (k & 3) == 0 && aligned16(x) && aligned16(w);
  for (int64_t r = warp0; r < rows; r += warps) {
    const float* xr = x + r * k;
    for (int o = 0; o < c; ++o) {
      const float* wr = w + (int64_t)o * k;
      float acc = 0.f;
      if (vec) {
        for (int i = lane; i < (k >> 2); i += 32) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(xr) + i);
          const float4 wv = __ldg(reinterpret_cast<const float4*>(wr) + i);
          acc = fmaf(xv.x, wv.x, acc); acc = fmaf(xv.y, wv.y, acc);
          acc = fmaf(xv.z, wv.z, acc); acc = fmaf(xv.w, wv.w, acc);
        }
      } else {
        for (int i = lane; i < k; i += 32) acc = fmaf(__ldg(xr + i), __ldg(wr + i), acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        const float v = apply_act(acc + (b ? __ldg(b + o) : 0.f), act);
        if (accumulate) out[r * c + o] += v;
        else out[r * c + o] = v;
      }
    }
  }
}

// Tall layers (K in the thousands, few outputs): out[r, o] = act(b[o] + <x[r, :], w[o, :]>) in plain FP32 FFMA with a
// K-chunked shared-memory tile: 64 rows x (16 CO) outputs per CTA, a thread owns 4 rows x CO outputs.  The first Linear
// of FiBiNET / DeepFFM / PNN-style MLPs (K = pairs x embed).  Why not the tensor pipe: tcgen05 accumulates K in the
// thousands inside the tensor core's own fp32 accumulator and drifts past the 1e-5 bar (2e-5 at K = 4 104, 1.5e-4 at
// K = 11 856, tests/test_gpu_more.py); FFMA keeps the parity and the layer is HBM-bound on x either way.
constexpr int kTallRows = 64, kTallKC = 64, kTallPitch = kTallKC + 4;
template <int CO>
__global__ void __launch_bounds__(256) tall_dense_kernel(const float* __restrict__ x, int64_t rows, int k,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         int c0, int c, int c_total, int act, float* __restrict__ out) {
  __shared__ __align__(16) float xs[kTallRows * kTallPitch];
  __shared__ __align__(16) float ws[16 * CO * kTallPitch];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool vec = (k & 3) == 0 && aligned16(x) && aligned16(w);
  for (int64_t r0 = (int64_t)blockIdx.x * kTallRows; r0 < rows; r0 += (int64_t)gridDim.x * kTallRows) {
    float acc[4][CO];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < k; k0 += kTallKC) {
      __syncthreads();
      for (int t = threadIdx.x; t < (kTallRows + 16 * CO) * (kTallKC / 4); t += blockDim.x) {
        const int row = t / (kTallKC / 4), kk = 4 * (t - row * (kTallKC / 4));
        const bool is_x = row < kTallRows;
        const int wr = row - kTallRows;                      // weight row within this block (when !is_x)
        const bool live = is_x ? (r0 + row < rows) : (wr < c);
        const float* src = is_x ? x + (r0 + row) * k : w + (int64_t)(c0 + wr) * k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
          if (vec && k0 + kk + 3 < k) {
            v = is_x ? ldg_stream_f4(reinterpret_cast<const float4*>(src + k0 + kk))
                     : __ldg(reinterpret_cast<const float4*>(src + k0 + kk));
          } else {
            if (k0 + kk < k) v.x = __ldg(src + k0 + kk);
            if (k0 + kk + 1 < k) v.y = __ldg(src + k0 + kk + 1);
            if (k0 + kk + 2 < k) v.z = __ldg(src + k0 + kk + 2);
            if (k0 + kk + 3 < k) v.w = __ldg(src + k0 + kk + 3);
          }
        }
        float* dst = is_x ? xs + row * kTallPitch + kk : ws + wr * kTallPitch + kk;
        *reinterpret_cast<float4*>(dst) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int kk = 0; kk < kTallKC; kk += 4) {
        float4 xv[4], wv[CO];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (4 * ty + i) * kTallPitch + kk);
#pragma unroll
        for (int j = 0; j < CO; ++j) wv[j] = *reinterpret_cast<const float4*>(ws + (tx + 16 * j) * kTallPitch + kk);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < CO; ++j) {
            acc[i][j] = fmaf(xv[i].x, wv[j].x, acc[i][j]);
            acc[i][j] = fmaf(xv[i].y, wv[j].y, acc[i][j]);
            acc[i][j] = fmaf(xv[i].z, wv[j].z, acc[i][j]);
            acc[i][j] = fmaf(xv[i].w, wv[j].w, acc[i][j]);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + 4 * ty + i;
      if (r >= rows) continue;
#pragma unroll
      for (int j = 0; j < CO; ++j) {
        const int o = tx + 16 * j;
        if (o < c) out[r * c_total + c0 + o] = apply_act(acc[i][j] + (b ? __ldg(b + c0 + o) : 0.f), act);
      }
    }
  }
}

// Tensor-core form of the tall layer (K in the thousands, <= 32 outputs; K a multiple of 4): the layer is a stream of x
// (8 FLOP per byte at 16 outputs), so the kernel reads x ONCE, straight from global memory into mma.sync fragments -- no
// shared-memory staging.  The sum over k is order-free, so the k slots of an m16n8k8 fragment are assigned to suit the
// loads: lane (g, t) reads the 16 bytes x[row g][k0 + 4t .. 4t + 3] (and row g + 8, and the same columns of W rows
// n = g of every 8-output tile); the first mma of a 16-k step takes components (.x, .y) as fragment columns (t, t + 4),
// the second (.z, .w).  3xTF32 (hi/lo split of both operands, small terms first) with FP32 accumulators that are
// flushed into FP32 master sums every 128 k, so the error does not grow with K the way a single long tensor-core
// accumulation does.  CTA = 4 warps on one 16-row tile, each warp a quarter of the k range; the quarters are added in a
// fixed order through shared memory (deterministic).
__device__ __forceinline__ void tall_mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void tall_split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(v - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

constexpr int kTallTcWarps = 4, kTallTcFlush = 8;   // flush the mma accumulators every 8 steps of 16 k

template <int NT, int U>
__global__ void __launch_bounds__(kTallTcWarps * 32, U == 2 ? 4 : 3) tall_dense_tc_kernel(const float* __restrict__ x, int64_t rows, int k,
                                                                          const float* __restrict__ w,
                                                                          const float* __restrict__ b, int c, int act,
                                                                          float* __restrict__ out) {
  __shared__ float red[kTallTcWarps][16][NT * 8 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int steps = (k + 15) / 16;
  const int s_begin = static_cast<int>((int64_t)steps * warp / kTallTcWarps);
  const int s_end = static_cast<int>((int64_t)steps * (warp + 1) / kTallTcWarps);
  const int64_t tiles = (rows + 15) / 16;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t r_lo = tile * 16 + g, r_hi = r_lo + 8;
    const float* x_lo = x + r_lo * k;
    const float* x_hi = x + r_hi * k;
    const bool lo_ok = r_lo < rows, hi_ok = r_hi < rows;
    float master[NT][4], acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) master[nt][q] = acc[nt][q] = 0.f;
    int since_flush = 0;
    for (int s = s_begin; s < s_end; s += U) {
      // U 16-k steps per iteration, all their loads issued before the first mma (U KB of x in flight per warp)
      float4 xa[U], xb[U], wv[U][NT];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = 16 * (s + u) + 4 * t;
        const bool in_k = s + u < s_end && kk < k;   // k is a multiple of 4: the 16 bytes are entirely inside or outside
        xa[u] = xb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_k && lo_ok) xa[u] = ldg_stream_f4(reinterpret_cast<const float4*>(x_lo + kk));
        if (in_k && hi_ok) xb[u] = ldg_stream_f4(reinterpret_cast<const float4*>(x_hi + kk));
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int n = nt * 8 + g;
          wv[u][nt] = (in_k && n < c) ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)n * k + kk))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t ah[8], al[8];   // (row g: x y z w, row g+8: x y z w)
        tall_split_tf32(xa[u].x, ah[0], al[0]);
        tall_split_tf32(xa[u].y, ah[1], al[1]);
        tall_split_tf32(xa[u].z, ah[2], al[2]);
        tall_split_tf32(xa[u].w, ah[3], al[3]);
        tall_split_tf32(xb[u].x, ah[4], al[4]);
        tall_split_tf32(xb[u].y, ah[5], al[5]);
        tall_split_tf32(xb[u].z, ah[6], al[6]);
        tall_split_tf32(xb[u].w, ah[7], al[7]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t bh[4], bl[4];
          tall_split_tf32(wv[u][nt].x, bh[0], bl[0]);
          tall_split_tf32(wv[u][nt].y, bh[1], bl[1]);
          tall_split_tf32(wv[u][nt].z, bh[2], bl[2]);
          tall_split_tf32(wv[u][nt].w, bh[3], bl[3]);
          // fragment columns (t, t+4) = components (.x, .y) in the first mma, (.z, .w) in the second
          tall_mma_tf32(acc[nt], al[0], al[4], al[1], al[5], bh[0], bh[1]);
          tall_mma_tf32(acc[nt], ah[0], ah[4], ah[1], ah[5], bl[0], bl[1]);
          tall_mma_tf32(acc[nt], ah[0], ah[4], ah[1], ah[5], bh[0], bh[1]);
          tall_mma_tf32(acc[nt], al[2], al[6], al[3], al[7], bh[2], bh[3]);
          tall_mma_tf32(acc[nt], ah[2], ah[6], ah[3], ah[7], bl[2], bl[3]);
          tall_mma_tf32(acc[nt], ah[2], ah[6], ah[3], ah[7], bh[2], bh[3]);
        }
      }
      since_flush += U;
      if (since_flush >= kTallTcFlush) {
        since_flush = 0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            master[nt][q] += acc[nt][q];
            acc[nt][q] = 0.f;
          }
      }
    }
    // accumulator layout: (row g, cols 2t, 2t+1), (row g+8, cols 2t, 2t+1) of every 8-output tile
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      red[warp][g][nt * 8 + 2 * t] = master[nt][0] + acc[nt][0];
      red[warp][g][nt * 8 + 2 * t + 1] = master[nt][1] + acc[nt][1];
      red[warp][g + 8][nt * 8 + 2 * t] = master[nt][2] + acc[nt][2];
      red[warp][g + 8][nt * 8 + 2 * t + 1] = master[nt][3] + acc[nt][3];
    }
    __syncthreads();
    for (int el = threadIdx.x; el < 16 * NT * 8; el += kTallTcWarps * 32) {
      const int r = el / (NT * 8), o = el - r * (NT * 8);
      const int64_t row = tile * 16 + r;
      if (row < rows && o < c) {
        float v = red[0][r][o];
#pragma unroll
        for (int q = 1; q < kTallTcWarps; ++q) v += red[q][r][o];
        out[row * c + o] = apply_act(v + (b ? __ldg(b + o) : 0.f), act);
      }
    }
    __syncthreads();
  }
}

static bool tall_dense_tc_ok(const float* x, int k, const float* w, int c) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  return !disabled && (k & 3) == 0 && c <= 32 && aligned16(x) && aligned16(w);
}

int tall_dense_run(const float* x, int64_t rows, int k, const float* w, const float* b, int c, int act, float* out,
                   cudaStream_t s) {
  if (tall_dense_tc_ok(x, k, w, c)) {
    const int64_t tiles16 = (rows + 15) / 16;
    const int grid16 = static_cast<int>(tiles16 < kNumSMs * 8 ? tiles16 : kNumSMs * 8);
    static const int unroll = getenv("TRS_TALL_U") ? atoi(getenv("TRS_TALL_U")) : 4;   // measured: 474 us (4) vs 520 us (2)
#define TRS_TALL_TC(NT_)                                                                                      \
  if (unroll == 4)                                                                                            \
    tall_dense_tc_kernel<NT_, 4><<<grid16, kTallTcWarps * 32, 0, s>>>(x, rows, k, w, b, c, act, out);          \
  else                                                                                                        \
    tall_dense_tc_kernel<NT_, 2><<<grid16, kTallTcWarps * 32, 0, s>>>(x, rows, k, w, b, c, act, out);
    switch ((c + 7) / 8) {
      case 1: TRS_TALL_TC(1) break;
      case 2: TRS_TALL_TC(2) break;
      case 3: TRS_TALL_TC(3) break;
      default: TRS_TALL_TC(4) break;
    }
#undef TRS_TALL_TC
    return check_launch("tall_dense_tc_kernel");
  }
  const int64_t tiles = (rows + kTallRows - 1) / kTallRows;
  const int grid = static_cast<int>(tiles < kNumSMs * 4 ? tiles : kNumSMs * 4);
  for (int c0 = 0; c0 < c; c0 += 64) {   // blocks of at most 64 outputs (x is re-read per block; C <= 64 is the case)
    const int cc = c - c0 < 64 ? c - c0 : 64;
    switch ((cc + 15) / 16) {
      case 1: tall_dense_kernel<1><<<grid, 256, 0, s>>>(x, rows, k, w, b, c0, cc, c, act, out); break;
      case 2: tall_dense_kernel<2><<<grid, 256, 0, s>>>(x, rows, k, w, b, c0, cc, c, act, out); break;
      case 3: tall_dense_kernel<3><<<grid, 256, 0, s>>>(x, rows, k, w, b, c0, cc, c, act, out); break;
      default: tall_dense_kernel<4><<<grid, 256, 0, s>>>(x, rows, k, w, b, c0, cc, c, act, out); break;
    }
    const int rc = check_launch("tall_dense_kernel");
    if (rc != TRS_OK) return rc;
  }
  return TRS_OK;
}

__global__ void __launch_bounds__(256) mlp_kernel(const float* __restrict__ x, int64_t rows, MlpParams mp, int ts,
                                                  int in_pitch, int hpitch, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* in = smem;
  float* buf0 = in + (size_t)ts * in_pitch;
  float* buf1 = buf0 + (size_t)ts * hpitch;
  const int k = mp.dims[0];
  const int o_dim = mp.dims[mp.layers];
  for (int64_t r0 = (int64_t)blockIdx.x * ts; r0 < rows; r0 += (int64_t)gridDim.x * ts) {
    const int n = static_cast<int>(rows - r0 < ts ? rows - r0 : ts);
    __syncthreads();
    for (int t = threadIdx.x; t < ts * k; t += blockDim.x) {
      const int s = t / k, c = t - s * k;
      in[s * in_pitch + c] = (s < n) ? ldg_stream_f1(x + (r0 + s) * k + c) : 0.f;
    }
    __syncthreads();
    const float* res = mlp_tile(mp, in, in_pitch, ts, buf0, buf1, hpitch);
    for (int t = threadIdx.x; t < n * o_dim; t += blockDim.x) {
      const int s = t / o_dim, o = t - s * o_dim;
      out[(r0 + s) * o_dim + o] = res[s * hpitch + o];
    }
  }
}

// h_{l+1} = x * (W_l h_l + b_l) + x on a tile of rows; x stays in shared memory, h ping-pongs.
__global__ void __launch_bounds__(256) cross_kernel(const float* __restrict__ x, const float* __restrict__ weights,
                                                    const float* __restrict__ biases, int layers, int64_t rows,
                                                    int embed, int ts, int pitch, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* h0 = xs + (size_t)ts * pitch;
  float* h1 = h0 + (size_t)ts * pitch;
  for (int64_t r0 = (int64_t)blockIdx.x * ts; r0 < rows; r0 += (int64_t)gridDim.x * ts) {
    const int n = static_cast<int>(rows - r0 < ts ? rows - r0 : ts);
    __syncthreads();
    for (int t = threadIdx.x; t < ts * embed; t += blockDim.x) {
      const int s = t / embed, c = t - s * embed;
      xs[s * pitch + c] = (s < n) ? ldg_stream_f1(x + (r0 + s) * embed + c) : 0.f;
    }
    __syncthreads();
    const float* cur = xs;
    float* dst = h0;
    for (int l = 0; l < layers; ++l) {
      float* d = dst;
      dense_layer_tile(cur, pitch, embed, weights + (int64_t)l * embed * embed, biases + (int64_t)l * embed, embed,
                       ts, [=](int s, int o, float v) {
                         const float x0 = xs[s * pitch + o];
                         d[s * pitch + o] = fmaf(x0, v, x0);  // x0 * v + x0
                       });
      __syncthreads();
      cur = dst;
      dst = (dst == h0) ? h1 : h0;
    }
    for (int t = threadIdx.x; t < n * embed; t += blockDim.x) {
      const int s = t / embed, c = t - s * embed;
      out[(r0 + s) * embed + c] = cur[s * pitch + c];
    }
  }
}

}  // namespace

// The MLP as a chain of kernels for WIDE layers: every layer with >= 32 outputs runs on tcgen05 (cin_tc.cu, 3xTF32,
// fp32-accurate), narrower ones (the logit layer) one warp per row; activations ping-pong through stream-ordered
// scratch.  Taken when some layer is at least 64 x 64 -- below that the one-kernel shared-memory tile MLP wins.
// `accumulate` adds the last layer's output onto `out` (the deep branch of DeepFM / xDeepFM joining the other terms).
// A layer of the chain runs on tcgen05 when it has >= 32 outputs; a TALL layer with few outputs (the first layer of
// FiBiNET / DeepFFM / PNN-style MLPs: K = pairs x embed in the thousands, 16 outputs) runs tall_dense_kernel -- the
// one-kernel shared-memory tile MLP would hold only a couple of such rows per CTA and re-read the whole weight matrix
// for each of them.
constexpr int kTallK = 1024;
static bool chain_layer_tall(int k_dim, int c_dim) { return k_dim >= kTallK && c_dim <= 64; }
static bool chain_layer_on_tc(int k_dim, int c_dim) { return !chain_layer_tall(k_dim, c_dim) && c_dim >= 32; }

int mlp_chain_supported(const int* dims, int layers, int64_t rows, const void* x, const void* out, int accumulate) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled || rows < 1024 || !aligned16(x) || !aligned16(out)) return 0;
  if (accumulate && dims[layers] >= 32) return 0;
  bool wide = false;
  for (int l = 0; l < layers; ++l) {
    if (chain_layer_tall(dims[l], dims[l + 1])) {
      wide = true;
      continue;
    }
    if (!chain_layer_on_tc(dims[l], dims[l + 1])) continue;
    if (!dense_tc_supported(dims[l], dims[l + 1], x, out)) return 0;
    if (dims[l] >= 64 && dims[l + 1] >= 64) wide = true;
  }
  return wide ? 1 : 0;
}

// The layer before a one-output Linear runs with that Linear folded into its epilogue (its activations are never stored)
// when it is a tensor-core layer; TRS_MLP_NO_DOT=1 switches that off (A/B measurements).
static bool chain_dot_fusable(const MlpParams& mp) {
  static const bool off = getenv("TRS_MLP_NO_DOT") != nullptr;
  const int L = mp.layers;
  return !off && L >= 2 && mp.dims[L] == 1 && chain_layer_on_tc(mp.dims[L - 2], mp.dims[L - 1]);
}

// Can the chain read its input rows straight from an embedding table (DenseFuse gather, cin_tc.cu)?  The first layer must
// be a tensor-core layer, the rows whole 16-column chunks, and the FM selector channels must fit one channel block.
int mlp_chain_gather_supported(const int* dims, int layers, int fields, int embed, int use_fm) {
  static const bool off = getenv("TRS_MLP_NO_GATHER") != nullptr;
  if (off || layers < 2 || embed % 16 != 0 || dims[0] != fields * embed || dims[layers] != 1) return 0;
  if (!chain_layer_on_tc(dims[0], dims[1]) || dims[0] < 16 || fields > 128) return 0;
  if (use_fm && embed > 64) return 0;
  return dense_tc_passes(dims[1], use_fm ? embed : 0, 1) > 0 ? 1 : 0;
}

// `gather` != null (mlp_chain_gather_supported): x is ignored, layer 0 gathers its rows itself and writes the row base
// (first-order + FM + bias) to out, onto which the last layer accumulates.
int mlp_chain_run(const float* x, int64_t rows, const MlpParams& mp, float* out, int accumulate, cudaStream_t s,
                  const DenseFuse* gather) {
  const int L = mp.layers;
  int hmax = 0;
  for (int l = 1; l < L; ++l) hmax = mp.dims[l] > hmax ? mp.dims[l] : hmax;
  const bool dot = chain_dot_fusable(mp);
  const int sel_last = (gather != nullptr && L == 2 && gather->use_fm) ? gather->embed : 0;
  const int dot_passes = dot ? dense_tc_passes(mp.dims[L - 1], sel_last, (gather != nullptr && L == 2) ? 1 : 0) : 0;
  const size_t half = ((size_t)rows * hmax + 63) / 64 * 64;
  const size_t dot_floats = ((size_t)rows * dot_passes + 63) / 64 * 64;
  float* buf = nullptr;
  if (L > 1) TRS_CUDA(scratch_alloc(reinterpret_cast<void**>(&buf), (2 * half + dot_floats) * sizeof(float), s));
  float* dot_buf = buf + 2 * half;
  if (gather != nullptr) accumulate = 1;
  const float* cur = x;
  int rc = TRS_OK;
  for (int l = 0; l < L && rc == TRS_OK; ++l) {
    const bool last = l == L - 1;
    if (last && dot) {
      rc = dense_dot_finish(dot_buf, dot_passes, rows, mp.b[l], out, accumulate, s);
      break;
    }
    float* dst = last ? out : buf + (l & 1) * half;
    const int act = last ? TRS_ACT_NONE : mp.act;
    if (chain_layer_tall(mp.dims[l], mp.dims[l + 1]) && !(last && accumulate)) {
      rc = tall_dense_run(cur, rows, mp.dims[l], mp.w[l], mp.b[l], mp.dims[l + 1], act, dst, s);
    } else if (chain_layer_on_tc(mp.dims[l], mp.dims[l + 1])) {
      DenseFuse fz;
      bool fused = false;
      if (l == 0 && gather != nullptr) {
        fz = *gather;
        fz.row_base = out;
        fused = true;
      }
      if (dot && l == L - 2) {
        fz.dot_w = mp.w[L - 1];
        fz.dot_out = dot_buf;
        fused = true;
      }
      rc = dense_tc_run(cur, rows, mp.dims[l], mp.w[l], mp.b[l], mp.dims[l + 1], act, dst, s, fused ? &fz : nullptr);
    } else {
      narrow_dense_kernel<<<grid_for(rows * 32, 256, 8), 256, 0, s>>>(cur, rows, mp.dims[l], mp.w[l], mp.b[l],
                                                                      mp.dims[l + 1], act, dst,
                                                                      last && accumulate ? 1 : 0);
      rc = check_launch("narrow_dense_kernel");
    }
    cur = dst;
  }
  if (buf) cudaFreeAsync(buf, s);
  return rc;
}

}  // namespace trs

using namespace trs;

extern "C" int trs_mlp_forward(const float* x, int64_t rows, const int* dims, int layers,
                               const float* const* weights, const float* const* biases, int activation, float* out,
                               void* stream) {
  TRS_REQUIRE(x && out && dims && weights, "trs_mlp_forward: null pointer");
  TRS_REQUIRE(rows >= 0 && layers >= 1, "trs_mlp_forward: bad sizes");
  MlpParams mp;
  TRS_REQUIRE(fill_mlp_params(mp, dims, layers, weights, biases, activation) == 0,
              "trs_mlp_forward: bad layer description (at most %d layers)", MlpParams::kMaxLayers);
  if (rows == 0) return TRS_OK;
  if (mlp_chain_supported(dims, layers, rows, x, out, 0))
    return mlp_chain_run(x, rows, mp, out, 0, static_cast<cudaStream_t>(stream), nullptr);
  const int in_pitch = tile_pitch(dims[0]);
  const int hpitch = tile_pitch(mlp_max_hidden(dims, layers));
  int ts = 64;
  size_t smem;
  for (;; ts >>= 1) {
    smem = ((size_t)ts * in_pitch + 2 * (size_t)ts * hpitch) * sizeof(float);
    if (smem <= (size_t)kMaxDynSmem - 1024 || ts == 1) break;
  }
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem - 1024, "trs_mlp_forward: layer widths do not fit shared memory");
  TRS_SMEM_OPT_IN(mlp_kernel);
  const int64_t tiles = (rows + ts - 1) / ts;
  const int grid = static_cast<int>(tiles < kNumSMs * 2 ? tiles : kNumSMs * 2);
  mlp_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, rows, mp, ts, in_pitch, hpitch, out);
  return check_launch("mlp_kernel");
}

extern "C" int trs_cross_forward(const float* x, const float* weights, const float* biases, int layers, int64_t rows,
                                 int embed, float* out, void* stream) {
  TRS_REQUIRE(x && out && (layers == 0 || (weights && biases)), "trs_cross_forward: null pointer");
  TRS_REQUIRE(rows >= 0 && embed > 0 && layers >= 0, "trs_cross_forward: bad sizes");
  if (rows == 0) return TRS_OK;
  // Large inputs, E in {16, 32}: the chain of a 128-row tile in tensor memory on tcgen05 (cross_tc5.cu: 141-155 us for
  // 1.28 M rows x 6 layers of width 32 against 241 us for the mma.sync chain, once its MMAs are issued under elect.sync).
  // Else E in {8,16,32,64}: register-resident 3xTF32 mma.sync chain (dcn_tc.cu); FP32 FFMA tiles otherwise.
  static const bool no_tc5 = getenv("TRS_DISABLE_TC5") != nullptr || getenv("TRS_DISABLE_TC") != nullptr;
  if (!no_tc5 && layers >= 1 && rows >= 32768) {
    const int rc = cross_tc5_launch(x, weights, biases, layers, rows, embed, out, static_cast<cudaStream_t>(stream));
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  {
    const int rc = cross_tc_launch(x, weights, biases, layers, rows, embed, out, static_cast<cudaStream_t>(stream));
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  const int pitch = tile_pitch(embed);
  int ts = 128;
  size_t smem;
  for (;; ts >>= 1) {
    smem = 3 * (size_t)ts * pitch * sizeof(float);
    if (smem <= 96 * 1024 || ts == 1) break;
  }
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem - 1024, "trs_cross_forward: embed too large");
  TRS_SMEM_OPT_IN(cross_kernel);
  const int64_t tiles = (rows + ts - 1) / ts;
  const int grid = static_cast<int>(tiles < kNumSMs * 2 ? tiles : kNumSMs * 2);
  cross_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, weights, biases, layers, rows, embed, ts,
                                                                        pitch, out);
  return check_launch("cross_kernel");
}
