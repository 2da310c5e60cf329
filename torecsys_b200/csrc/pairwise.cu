// a5/a6/a9/a10/a11: the pairwise interaction layers on a materialised (B, N, E) tile.
//
// These are the L1 drop-ins (module boundary = a (B,N,E) tensor in HBM).  All of them are HBM-bound
// (FM, FFM, bilinear 'all': one read or one write of a big tensor) or FFMA-bound (IPN, AFM, bilinear 'each')
// in plain fp32 -- no reduced precision anywhere, the parity bar is 1e-5.
#include "common.cuh"

namespace trs {

int ipn_tc_launch(const float* x, int64_t batch, int fields, int embed, float* out, cudaStream_t s);
int bilinear_tc_launch(const float* x, const float* w, const float* bias, int each_type, int64_t batch, int fields,
                       int embed, int64_t out_stride, float* out, cudaStream_t s);
int afm_tc_launch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch,
                  int fields, int embed, int attn, float* out, float* scores, cudaStream_t s);
int afm_tc5_launch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch,
                   int fields, int embed, int attn, float* out, float* scores, cudaStream_t s);

namespace {

// ------------------------------------------------------------------------------------------------ FM (a5)
// thread = (sample, 4-wide column chunk); N independent 128-bit loads per thread, fully unrolled by 4.
template <bool Vec>
__global__ void __launch_bounds__(256) fm_kernel(const float* __restrict__ x, int64_t batch, int fields, int embed,
                                                 float* __restrict__ out) {
  constexpr int W = Vec ? 4 : 1;
  const int per_row = embed / W;
  const int64_t items = batch * per_row;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = item / per_row;
    const int c = static_cast<int>(item - b * per_row);
    const float* base = x + b * fields * embed + c * W;
    if (Vec) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 4
      for (int n = 0; n < fields; ++n) {
        float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(base + (int64_t)n * embed));
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
      }
      float4 o;
      o.x = 0.5f * (s.x * s.x - q.x); o.y = 0.5f * (s.y * s.y - q.y);
      o.z = 0.5f * (s.z * s.z - q.z); o.w = 0.5f * (s.w * s.w - q.w);
      reinterpret_cast<float4*>(out)[item] = o;
    } else {
      float s = 0.f, q = 0.f;
#pragma unroll 4
      for (int n = 0; n < fields; ++n) {
        float v = ldg_stream_f1(base + (int64_t)n * embed);
        s += v;
        q = fmaf(v, v, q);
      }
      out[item] = 0.5f * (s * s - q);
    }
  }
}

// ------------------------------------------------------------------------------------------------ FFM (a6)
// item = (sample, pair, chunk): two 128-bit loads, one multiply, one streaming store.  The (i,j) of a pair comes
// from a table built once per CTA in shared memory.
template <bool Vec>
__global__ void __launch_bounds__(256) ffm_kernel(const float* __restrict__ v, int64_t batch, int fields, int embed,
                                                  float* __restrict__ out) {
  extern __shared__ int pair_tab[];  // pair p -> (i*N+j) << 16 | (j*N+i)
  constexpr int W = Vec ? 4 : 1;
  const int pairs = fields * (fields - 1) / 2;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    pair_tab[p] = ((i * fields + j) << 16) | (j * fields + i);
  }
  __syncthreads();
  const int per_row = embed / W;
  const int64_t per_sample = (int64_t)pairs * per_row;
  const int64_t items = batch * per_sample;
  const int64_t in_sample = (int64_t)fields * fields * embed;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = item / per_sample;
    const int rem = static_cast<int>(item - b * per_sample);
    const int p = rem / per_row;
    const int c = rem - p * per_row;
    const int t = pair_tab[p];
    const float* a = v + b * in_sample + (int64_t)(t >> 16) * embed + c * W;
    const float* bb = v + b * in_sample + (int64_t)(t & 0xffff) * embed + c * W;
    if (Vec) {
      float4 u = ldg_stream_f4(reinterpret_cast<const float4*>(a));
      float4 w = ldg_stream_f4(reinterpret_cast<const float4*>(bb));
      stg_stream_f4(reinterpret_cast<float4*>(out) + item, make_float4(u.x * w.x, u.y * w.y, u.z * w.z, u.w * w.w));
    } else {
      out[item] = ldg_stream_f1(a) * ldg_stream_f1(bb);
    }
  }
}

// ------------------------------------------------------------------------------------------------ IPN (a9)
// One warp per sample.  The sample's (N, E) tile is staged TRANSPOSED in shared memory (xt[e][n], row pitch
// padded to an odd number of words), then lane <-> pair: consecutive pairs share i (broadcast) and have
// consecutive j (conflict-free).  Generic in N and E.
__global__ void __launch_bounds__(256) ipn_kernel(const float* __restrict__ x, int64_t batch, int fields, int embed,
                                                  float* __restrict__ out) {
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
      int n = t / embed, e = t - n * embed;
      xt[e * pitch + n] = ldg_stream_f1(src + t);
    }
    __syncwarp();
    float* dst = out + b * pairs;
    for (int p = lane; p < pairs; p += 32) {
      const int ij = ptab[p];
      const int i = ij >> 16, j = ij & 0xffff;
      float acc = 0.f;
#pragma unroll 4
      for (int e = 0; e < embed; ++e) acc = fmaf(xt[e * pitch + i], xt[e * pitch + j], acc);
      dst[p] = acc;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ bilinear (a10)
// 'all':  CTA stages W (E,E) once; per sample a warp computes y_n = x_n @ W for all fields into shared memory, then
//         streams out[p,:] = y_i * x_j + bias.  Output-write bound (P*E*4 bytes per sample).
__global__ void __launch_bounds__(256) bilinear_all_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int64_t batch, int fields,
                                                           int embed, float* __restrict__ out) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = fields * embed;
  const int pairs = fields * (fields - 1) / 2;
  float* ws = smem;                                  // (E, E)
  float* bs = ws + embed * embed;                    // (E)
  float* xs = bs + embed + (size_t)warp * 2 * tile;  // per warp: x tile then y tile
  float* ys = xs + tile;
  int* ptab = reinterpret_cast<int*>(smem + embed * embed + embed + (size_t)warps * 2 * tile);
  for (int t = threadIdx.x; t < embed * embed; t += blockDim.x) ws[t] = __ldg(w + t);
  for (int t = threadIdx.x; t < embed; t += blockDim.x) bs[t] = bias ? __ldg(bias + t) : 0.f;
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  __syncthreads();
  for (int64_t b = (int64_t)blockIdx.x * warps + warp; b < batch; b += (int64_t)gridDim.x * warps) {
    const float* src = x + b * tile;
    for (int t = lane; t < tile; t += 32) xs[t] = ldg_stream_f1(src + t);
    __syncwarp();
    for (int t = lane; t < tile; t += 32) {  // y[n][o] = sum_k x[n][k] * W[k][o]
      const int n = t / embed, o = t - n * embed;
      float acc = 0.f;
#pragma unroll 4
      for (int k = 0; k < embed; ++k) acc = fmaf(xs[n * embed + k], ws[k * embed + o], acc);
      ys[t] = acc;
    }
    __syncwarp();
    float* dst = out + b * (int64_t)pairs * embed;
    const int total = pairs * embed;
    for (int t = lane; t < total; t += 32) {
      const int p = t / embed, e = t - p * embed;
      const int ij = ptab[p];
      dst[t] = fmaf(ys[(ij >> 16) * embed + e], xs[(ij & 0xffff) * embed + e], bs[e]);
    }
    __syncwarp();
  }
}

// 'each': one CTA per (pair, batch tile): W_p (E,E) staged once in shared memory and reused over the tile, so the
//         759 KB of pair weights are read from L2 once per 256 samples instead of once per sample.
constexpr int kEachTile = 256;
__global__ void __launch_bounds__(256) bilinear_each_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int64_t batch,
                                                            int fields, int embed, float* __restrict__ out) {
  extern __shared__ float smem[];
  float* ws = smem;                // (E, E) of this pair
  float* bs = ws + embed * embed;  // (E)
  const int pairs = fields * (fields - 1) / 2;
  const int p = blockIdx.x;
  int i, j;
  pair_from_index(p, fields, i, j);
  const float* wp = w + (int64_t)p * embed * embed;
  for (int t = threadIdx.x; t < embed * embed; t += blockDim.x) ws[t] = __ldg(wp + t);
  for (int t = threadIdx.x; t < embed; t += blockDim.x) bs[t] = bias ? __ldg(bias + (int64_t)p * embed + t) : 0.f;
  __syncthreads();
  const int64_t b0 = (int64_t)blockIdx.y * kEachTile;
  const int nb = static_cast<int>(batch - b0 < kEachTile ? batch - b0 : kEachTile);
  const int total = nb * embed;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int s = t / embed, o = t - s * embed;
    const float* xi = x + ((b0 + s) * fields + i) * embed;
    const float* xj = x + ((b0 + s) * fields + j) * embed;
    float acc = 0.f;
#pragma unroll 4
    for (int k = 0; k < embed; ++k) acc = fmaf(__ldg(xi + k), ws[k * embed + o], acc);
    out[((b0 + s) * pairs + p) * embed + o] = fmaf(acc, __ldg(xj + o), bs[o]);
  }
}

// ------------------------------------------------------------------------------------------------ AFM (a11)
// One CTA per sample.  Stage x (N,E) and W1/b1/w2 in shared memory; each thread owns pairs p = tid, tid+T, ...:
// score_p = w2 . relu(W1 (x_i*x_j) + b1) + b2 kept in shared memory; block softmax over the P scores; then
// out[e] = sum_p s_p x_i[e] x_j[e] with threads striding over e and a shared-memory reduction over pair slices.
__global__ void __launch_bounds__(256) afm_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                  const float* __restrict__ b1, const float* __restrict__ w2,
                                                  const float* __restrict__ b2, int64_t batch, int fields, int embed,
                                                  int attn, float* __restrict__ out, float* __restrict__ scores) {
  extern __shared__ float smem[];
  const int pairs = fields * (fields - 1) / 2;
  const int tile = fields * embed;
  const int epitch = embed | 1;
  float* xs = smem;                    // (N, epitch)
  float* w1s = xs + fields * epitch;   // (A, E)
  float* b1s = w1s + attn * embed;     // (A)
  float* w2s = b1s + attn;             // (A)
  float* sc = w2s + attn;              // (P)
  float* red = sc + pairs;             // (blockDim.x)  reduction scratch
  int* ptab = reinterpret_cast<int*>(red + blockDim.x);
  for (int t = threadIdx.x; t < attn * embed; t += blockDim.x) w1s[t] = __ldg(w1 + t);
  for (int t = threadIdx.x; t < attn; t += blockDim.x) {
    b1s[t] = __ldg(b1 + t);
    w2s[t] = __ldg(w2 + t);
  }
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  const float bias2 = __ldg(b2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  for (int64_t b = blockIdx.x; b < batch; b += gridDim.x) {
    __syncthreads();
    const float* src = x + b * tile;
    for (int t = threadIdx.x; t < tile; t += blockDim.x) {
      int n = t / embed, e = t - n * embed;
      xs[n * epitch + e] = ldg_stream_f1(src + t);
    }
    __syncthreads();
    // scores
    float lmax = -INFINITY;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      const int ij = ptab[p];
      const float* xi = xs + (ij >> 16) * epitch;
      const float* xj = xs + (ij & 0xffff) * epitch;
      float s = bias2;
      for (int a = 0; a < attn; ++a) {
        float h = b1s[a];
        const float* wr = w1s + a * embed;
#pragma unroll 4
        for (int e = 0; e < embed; ++e) h = fmaf(wr[e], xi[e] * xj[e], h);
        s = fmaf(w2s[a], fmaxf(h, 0.f), s);
      }
      sc[p] = s;
      lmax = fmaxf(lmax, s);
    }
    lmax = warp_max(lmax);
    if (lane == 0) red[warp] = lmax;
    __syncthreads();
    float gmax = -INFINITY;
    for (int w = 0; w < warps; ++w) gmax = fmaxf(gmax, red[w]);
    __syncthreads();
    float lsum = 0.f;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      float ev = expf(sc[p] - gmax);
      sc[p] = ev;
      lsum += ev;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    float gsum = 0.f;
    for (int w = 0; w < warps; ++w) gsum += red[w];
    const float inv = 1.0f / gsum;
    __syncthreads();
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      float sv = sc[p] * inv;
      sc[p] = sv;
      scores[b * pairs + p] = sv;
    }
    __syncthreads();
    // weighted sum: thread (slice, e): slices of pairs reduced through shared memory
    const int slices = blockDim.x / embed > 0 ? blockDim.x / embed : 1;
    float acc = 0.f;
    const int e = threadIdx.x % embed, slice = threadIdx.x / embed;
    if (slice < slices) {
      for (int p = slice; p < pairs; p += slices) {
        const int ij = ptab[p];
        acc = fmaf(sc[p], xs[(ij >> 16) * epitch + e] * xs[(ij & 0xffff) * epitch + e], acc);
      }
    }
    red[threadIdx.x] = (slice < slices) ? acc : 0.f;
    __syncthreads();
    for (int e2 = threadIdx.x; e2 < embed; e2 += blockDim.x) {
      float tot = 0.f;
      if (embed <= (int)blockDim.x) {
        for (int sl = 0; sl < slices; ++sl) tot += red[sl * embed + e2];
      }
      out[b * embed + e2] = tot;
    }
    if (embed > (int)blockDim.x) {  // very wide embeddings: one thread per column, all pairs
      __syncthreads();
      for (int e2 = threadIdx.x; e2 < embed; e2 += blockDim.x) {
        float tot = 0.f;
        for (int p = 0; p < pairs; ++p) {
          const int ij = ptab[p];
          tot = fmaf(sc[p], xs[(ij >> 16) * epitch + e2] * xs[(ij & 0xffff) * epitch + e2], tot);
        }
        out[b * embed + e2] = tot;
      }
    }
  }
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_fm_forward(const float* x, int64_t batch, int fields, int embed, float* out, void* stream) {
  TRS_REQUIRE(x && out, "trs_fm_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 0 && embed > 0, "trs_fm_forward: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = embed % 4 == 0 && aligned16(x) && aligned16(out);
  const int64_t items = batch * (vec ? embed / 4 : embed);
  const int grid = grid_for(items, 256, 8);
  if (vec) fm_kernel<true><<<grid, 256, 0, s>>>(x, batch, fields, embed, out);
  else fm_kernel<false><<<grid, 256, 0, s>>>(x, batch, fields, embed, out);
  return check_launch("fm_kernel");
}

extern "C" int trs_ffm_forward(const float* v, int64_t batch, int fields, int embed, float* out, void* stream) {
  TRS_REQUIRE(v && out, "trs_ffm_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_ffm_forward: bad sizes");
  TRS_REQUIRE(fields <= 255, "trs_ffm_forward: at most 255 fields");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = embed % 4 == 0 && aligned16(v) && aligned16(out);
  const int pairs = fields * (fields - 1) / 2;
  const int64_t items = batch * pairs * (vec ? embed / 4 : embed);
  const int grid = grid_for(items, 256, 8);
  const size_t smem = pairs * sizeof(int);
  if (vec) ffm_kernel<true><<<grid, 256, smem, s>>>(v, batch, fields, embed, out);  // <= 32 KB: no opt-in needed
  else ffm_kernel<false><<<grid, 256, smem, s>>>(v, batch, fields, embed, out);
  return check_launch("ffm_kernel");
}

extern "C" int trs_ipn_forward(const float* x, int64_t batch, int fields, int embed, float* out, void* stream) {
  TRS_REQUIRE(x && out, "trs_ipn_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_ipn_forward: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {  // Gram matrix on the tensor pipe (ipn_tc.cu) when the shape allows
    const int rc = ipn_tc_launch(x, batch, fields, embed, out, s);
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  const int pairs = fields * (fields - 1) / 2;
  const int pitch = fields | 1;
  int warps = 8;
  size_t smem;
  for (;; warps >>= 1) {
    smem = ((size_t)warps * embed * pitch + pairs) * sizeof(float);
    if (smem <= 200 * 1024 || warps == 1) break;
  }
  TRS_UNSUPPORTED(smem > 200 * 1024, "trs_ipn_forward: fields*embed tile does not fit shared memory");
  TRS_SMEM_OPT_IN(ipn_kernel);
  const int grid = grid_for(batch * 32, warps * 32, 4);
  ipn_kernel<<<grid, warps * 32, smem, s>>>(x, batch, fields, embed, out);
  return check_launch("ipn_kernel");
}

extern "C" int trs_bilinear_forward_strided(const float* x, const float* weight, const float* bias, int each_type,
                                            int64_t batch, int fields, int embed, int64_t out_stride, float* out,
                                            void* stream) {
  TRS_REQUIRE(x && weight && out, "trs_bilinear_forward_strided: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_bilinear_forward_strided: bad sizes");
  TRS_REQUIRE(out_stride >= (int64_t)fields * (fields - 1) / 2 * embed,
              "trs_bilinear_forward_strided: out_stride smaller than one sample's output");
  if (batch == 0) return TRS_OK;
  const int rc = bilinear_tc_launch(x, weight, bias, each_type, batch, fields, embed, out_stride, out,
                                    static_cast<cudaStream_t>(stream));
  if (rc == TRS_ERR_UNSUPPORTED)
    set_error("trs_bilinear_forward_strided: embed must be 8, 16 or 32, x / out 16-byte aligned, out_stride a multiple "
              "of 4 (got embed %d)", embed);
  return rc;
}

extern "C" int trs_bilinear_forward(const float* x, const float* weight, const float* bias, int each_type,
                                    int64_t batch, int fields, int embed, float* out, void* stream) {
  TRS_REQUIRE(x && weight && out, "trs_bilinear_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0, "trs_bilinear_forward: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int pairs = fields * (fields - 1) / 2;
  {   // tensor-pipe kernel (bilinear_tc.cu) for embed 8 / 16 / 32; anything else takes the generic kernels below
    const int rc = bilinear_tc_launch(x, weight, bias, each_type, batch, fields, embed, (int64_t)pairs * embed, out, s);
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  if (!each_type) {
    int warps = 8;
    size_t smem;
    for (;; warps >>= 1) {
      smem = ((size_t)embed * embed + embed + (size_t)warps * 2 * fields * embed + pairs) * sizeof(float);
      if (smem <= 200 * 1024 || warps == 1) break;
    }
    TRS_UNSUPPORTED(smem > 200 * 1024, "trs_bilinear_forward: tile does not fit shared memory");
    TRS_SMEM_OPT_IN(bilinear_all_kernel);
    const int grid = grid_for(batch * 32, warps * 32, 4);
    bilinear_all_kernel<<<grid, warps * 32, smem, s>>>(x, weight, bias, batch, fields, embed, out);
    return check_launch("bilinear_all_kernel");
  }
  const size_t smem = ((size_t)embed * embed + embed) * sizeof(float);
  TRS_UNSUPPORTED(smem > 200 * 1024, "trs_bilinear_forward: embed too large for the 'each' kernel");
  TRS_SMEM_OPT_IN(bilinear_each_kernel);
  const int64_t tiles = (batch + kEachTile - 1) / kEachTile;
  TRS_UNSUPPORTED(tiles > 65535, "trs_bilinear_forward: batch too large for one launch of the 'each' kernel");
  dim3 grid(pairs, (unsigned)tiles);
  bilinear_each_kernel<<<grid, 256, smem, s>>>(x, weight, bias, batch, fields, embed, out);
  return check_launch("bilinear_each_kernel");
}

extern "C" int trs_afm_forward(const float* x, const float* w1, const float* b1, const float* w2, const float* b2,
                               int64_t batch, int fields, int embed, int attn, float* out, float* scores,
                               void* stream) {
  TRS_REQUIRE(x && w1 && b1 && w2 && b2 && out && scores, "trs_afm_forward: null pointer");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0 && attn > 0, "trs_afm_forward: bad sizes");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {  // large batches, embed 16 / 32, attn <= 32: tcgen05 with the pair products in tensor memory (afm_tc5.cu)
    const int rc = afm_tc5_launch(x, w1, b1, w2, b2, batch, fields, embed, attn, out, scores, s);
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  {  // the pair x embed x attn contraction on the tensor pipe (afm_tc.cu) when the shape allows
    const int rc = afm_tc_launch(x, w1, b1, w2, b2, batch, fields, embed, attn, out, scores, s);
    if (rc != TRS_ERR_UNSUPPORTED) return rc;
  }
  const int pairs = fields * (fields - 1) / 2;
  const int threads = 256;
  const size_t smem = ((size_t)fields * (embed | 1) + (size_t)attn * embed + 2 * attn + pairs + threads + pairs) *
                      sizeof(float);
  TRS_UNSUPPORTED(smem > 200 * 1024, "trs_afm_forward: tile does not fit shared memory");
  TRS_SMEM_OPT_IN(afm_kernel);
  const int grid = static_cast<int>(batch < (int64_t)kNumSMs * 8 ? batch : (int64_t)kNumSMs * 8);
  afm_kernel<<<grid, threads, smem, s>>>(x, w1, b1, w2, b2, batch, fields, embed, attn, out, scores);
  return check_launch("afm_kernel");
}
