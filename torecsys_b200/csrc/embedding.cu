// a1/a2/a3: embedding row gathers (bit-exact copies of fp32 rows).
//
// HBM-bound: per looked-up row the kernel reads one index (8 B, coalesced), E*4 B of table row (random,
// 128-bit loads, L1 no-allocate) and writes E*4 B (coalesced).  Each thread keeps kUnroll independent
// row loads in flight (index loads first, then all row loads, then the stores) so that with 8 resident
// CTAs of 256 threads an SM has far more than the ~35 KB outstanding that 148 SMs need to cover the
// HBM latency-bandwidth product (SURVEY.md section 7 "Hard parts").
#include "common.cuh"

namespace trs {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// One work item = one 16-byte chunk of one output row.  chunks = embed/4.
template <int IdxBits>
__global__ void __launch_bounds__(kThreads) gather_rows_vec_kernel(
    const float4* __restrict__ weight, int64_t rows, uint32_t chunks, const void* __restrict__ idx,
    const int64_t* __restrict__ offsets, uint32_t fields, uint32_t items, int64_t pos_base,
    float4* __restrict__ out, int32_t* status) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < items; base += stride * kUnroll) {
    int64_t row[kUnroll];
    uint32_t chunk[kUnroll];
    bool live[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint32_t item = base + u * stride;
      live[u] = item < items;
      uint32_t pos = live[u] ? item / chunks : 0;
      chunk[u] = live[u] ? item - pos * chunks : 0;
      uint32_t n = pos % fields;
      int64_t r = live[u] ? load_index<IdxBits>(idx, pos_base + pos) : 0;
      if (offsets != nullptr) r += __ldg(offsets + n);
      if (live[u] && (r < 0 || r >= rows)) {
        if (chunk[u] == 0) report_oob(status, pos_base + pos);
        r = -1;
      }
      row[u] = r;
    }
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live[u] && row[u] >= 0) v[u] = ldg_stream_f4(weight + row[u] * chunks + chunk[u]);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint32_t item = base + u * stride;
      if (live[u]) out[item] = v[u];
    }
  }
}

// Generic element-wise variant (embed not a multiple of 4, e.g. the first-order table with embed = 1).
template <int IdxBits>
__global__ void __launch_bounds__(kThreads) gather_rows_scalar_kernel(
    const float* __restrict__ weight, int64_t rows, uint32_t embed, const void* __restrict__ idx,
    const int64_t* __restrict__ offsets, uint32_t fields, uint32_t items, int64_t pos_base,
    float* __restrict__ out, int32_t* status) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < items; base += stride * kUnroll) {
    int64_t src[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint32_t item = base + u * stride;
      src[u] = -1;
      if (item < items) {
        uint32_t pos = item / embed;
        uint32_t e = item - pos * embed;
        int64_t r = load_index<IdxBits>(idx, pos_base + pos);
        if (offsets != nullptr) r += __ldg(offsets + pos % fields);
        if (r < 0 || r >= rows) {
          if (e == 0) report_oob(status, pos_base + pos);
        } else {
          src[u] = r * embed + e;
        }
      }
    }
    float v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = src[u] >= 0 ? ldg_stream_f1(weight + src[u]) : 0.0f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint32_t item = base + u * stride;
      if (item < items) out[item] = v[u];
    }
  }
}

// Field-aware: out[b, t*N + f, :] = tables[t][idx[b,f] + off[f], :].  item -> (b, t, f, chunk).
template <int IdxBits, bool Vec>
__global__ void __launch_bounds__(kThreads) gather_field_aware_kernel(
    const float* const* __restrict__ tables, int64_t rows, uint32_t embed, const void* __restrict__ idx,
    const int64_t* __restrict__ offsets, uint32_t fields, uint32_t items, int64_t batch_base,
    float* __restrict__ out, int32_t* status) {
  constexpr int W = Vec ? 4 : 1;
  const uint32_t per_row = embed / W;
  const uint32_t nn = fields * fields;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x; base < items; base += stride * kUnroll) {
    const float* src[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      uint32_t item = base + u * stride;
      src[u] = nullptr;
      if (item < items) {
        uint32_t orow = item / per_row;
        uint32_t c = item - orow * per_row;
        uint32_t b = orow / nn;
        uint32_t tf = orow - b * nn;
        uint32_t t = tf / fields;
        uint32_t f = tf - t * fields;
        int64_t pos = (batch_base + b) * fields + f;
        int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + f);
        if (r < 0 || r >= rows) {
          if (c == 0 && t == 0) report_oob(status, pos);
        } else {
          src[u] = tables[t] + r * embed + c * W;
        }
      }
    }
    if (Vec) {
      float4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
        v[u] = src[u] ? ldg_stream_f4(reinterpret_cast<const float4*>(src[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        uint32_t item = base + u * stride;
        if (item < items) stg_stream_f4(reinterpret_cast<float4*>(out) + item, v[u]);
      }
    } else {
      float v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = src[u] ? ldg_stream_f1(src[u]) : 0.0f;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        uint32_t item = base + u * stride;
        if (item < items) out[item] = v[u];
      }
    }
  }
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_embedding_gather(const float* weight, int64_t rows, int embed, const void* idx, int idx_bits,
                                    const int64_t* offsets, int64_t batch, int fields, float* out, int32_t* status,
                                    void* stream) {
  TRS_REQUIRE(weight && idx && out, "trs_embedding_gather: null pointer");
  TRS_REQUIRE(rows > 0 && embed > 0 && batch >= 0 && fields > 0, "trs_embedding_gather: bad sizes");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_embedding_gather: idx_bits must be 32 or 64");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (embed % 4 == 0) && aligned16(weight) && aligned16(out);
  const int64_t per_row = vec ? embed / 4 : embed;
  // keep every launch below 2^31 work items so the kernels can use 32-bit div/mod
  const int64_t max_pos = ((int64_t(1) << 31) - 1) / per_row;
  const int64_t total_pos = batch * fields;
  const int64_t pos_step = (max_pos / fields) * fields > 0 ? (max_pos / fields) * fields : fields;
  for (int64_t p0 = 0; p0 < total_pos; p0 += pos_step) {
    const int64_t npos = (total_pos - p0 < pos_step) ? total_pos - p0 : pos_step;
    const uint32_t items = static_cast<uint32_t>(npos * per_row);
    const int grid = grid_for((items + kUnroll - 1) / kUnroll, kThreads, 8);
    if (vec) {
      auto w4 = reinterpret_cast<const float4*>(weight);
      auto o4 = reinterpret_cast<float4*>(out) + p0 * per_row;
      if (idx_bits == 64)
        gather_rows_vec_kernel<64><<<grid, kThreads, 0, s>>>(w4, rows, (uint32_t)per_row, idx, offsets, fields,
                                                              items, p0, o4, status);
      else
        gather_rows_vec_kernel<32><<<grid, kThreads, 0, s>>>(w4, rows, (uint32_t)per_row, idx, offsets, fields,
                                                              items, p0, o4, status);
    } else {
      float* o = out + p0 * per_row;
      if (idx_bits == 64)
        gather_rows_scalar_kernel<64><<<grid, kThreads, 0, s>>>(weight, rows, embed, idx, offsets, fields, items,
                                                                 p0, o, status);
      else
        gather_rows_scalar_kernel<32><<<grid, kThreads, 0, s>>>(weight, rows, embed, idx, offsets, fields, items,
                                                                 p0, o, status);
    }
    int rc = check_launch("gather_rows");
    if (rc != TRS_OK) return rc;
  }
  return TRS_OK;
}

extern "C" int trs_embedding_gather_field_aware(const float* const* tables, int64_t rows, int embed,
                                                const void* idx, int idx_bits, const int64_t* offsets,
                                                int64_t batch, int fields, float* out, int32_t* status,
                                                void* stream) {
  TRS_REQUIRE(tables && idx && out && offsets, "trs_embedding_gather_field_aware: null pointer");
  TRS_REQUIRE(rows > 0 && embed > 0 && batch >= 0 && fields > 0, "trs_embedding_gather_field_aware: bad sizes");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_embedding_gather_field_aware: idx_bits must be 32 or 64");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // table pointers live in device memory, so their alignment is checked by the caller contract:
  // rows of `embed % 4 == 0` floats from a 16-byte aligned base (torch allocations are 256-byte aligned).
  const bool vec = (embed % 4 == 0) && aligned16(out);
  const int64_t per_sample = int64_t(fields) * fields * (vec ? embed / 4 : embed);
  int64_t b_step = ((int64_t(1) << 31) - 1) / per_sample;
  TRS_REQUIRE(b_step >= 1, "trs_embedding_gather_field_aware: fields*fields*embed too large");
  for (int64_t b0 = 0; b0 < batch; b0 += b_step) {
    const int64_t nb = (batch - b0 < b_step) ? batch - b0 : b_step;
    const uint32_t items = static_cast<uint32_t>(nb * per_sample);
    const int grid = grid_for((items + kUnroll - 1) / kUnroll, kThreads, 8);
    float* o = out + b0 * int64_t(fields) * fields * embed;
#define LAUNCH(BITS, VEC)                                                                                     \
  gather_field_aware_kernel<BITS, VEC><<<grid, kThreads, 0, s>>>(tables, rows, embed, idx, offsets, fields, \
                                                                  items, b0, o, status)
    if (idx_bits == 64) {
      if (vec) LAUNCH(64, true); else LAUNCH(64, false);
    } else {
      if (vec) LAUNCH(32, true); else LAUNCH(32, false);
    }
#undef LAUNCH
    int rc = check_launch("gather_field_aware");
    if (rc != TRS_OK) return rc;
  }
  return TRS_OK;
}

// ---- a4: the column concatenation of Inputs.forward (torecsys/inputs/inputs.py:76-81) -----------------------------------
// The DataLoader hands the model one (B,) or (B, w) index tensor per feature field; Inputs.forward concatenates the
// columns an embedding asked for into its (B, N) index matrix (39 tensors for the Criteo shape).  One kernel: a CTA
// transposes a tile of rows through shared memory, so every column is read and the matrix is written with full
// coalesced requests.
namespace trs {
namespace {

constexpr int kMaxConcatCols = 128;

struct ConcatArgs {
  const void* col[kMaxConcatCols];
  int width[kMaxConcatCols];
  int off[kMaxConcatCols];
  int ncols, total, tile_rows;
  int64_t batch;
  void* out;
};

template <typename T>
__global__ void __launch_bounds__(256) index_concat_kernel(ConcatArgs a) {
  extern __shared__ __align__(16) unsigned char concat_smem[];
  T* tile = reinterpret_cast<T*>(concat_smem);
  T* out = static_cast<T*>(a.out);
  for (int64_t b0 = (int64_t)blockIdx.x * a.tile_rows; b0 < a.batch; b0 += (int64_t)gridDim.x * a.tile_rows) {
    const int nb = static_cast<int>(a.batch - b0 < a.tile_rows ? a.batch - b0 : a.tile_rows);
    for (int c = 0; c < a.ncols; ++c) {
      const int w = a.width[c];
      const T* src = static_cast<const T*>(a.col[c]) + b0 * w;
      for (int t = threadIdx.x; t < nb * w; t += blockDim.x) {
        const int r = t / w, j = t - r * w;
        tile[r * a.total + a.off[c] + j] = src[t];
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nb * a.total; t += blockDim.x) out[b0 * a.total + t] = tile[t];
    __syncthreads();
  }
}

}  // namespace
}  // namespace trs

extern "C" int trs_index_concat(const void* const* columns, const int* widths, int ncols, int idx_bits, int64_t batch,
                                void* out, void* stream) {
  TRS_REQUIRE(columns && widths && out, "trs_index_concat: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_index_concat: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && ncols > 0, "trs_index_concat: bad sizes");
  TRS_UNSUPPORTED(ncols > kMaxConcatCols, "trs_index_concat: at most %d columns per call", kMaxConcatCols);
  ConcatArgs a{};
  int total = 0;
  for (int c = 0; c < ncols; ++c) {
    TRS_REQUIRE(columns[c] && widths[c] > 0, "trs_index_concat: bad column %d", c);
    a.col[c] = columns[c];
    a.width[c] = widths[c];
    a.off[c] = total;
    total += widths[c];
  }
  if (batch == 0) return TRS_OK;
  const int esz = idx_bits / 8;
  int tile_rows = 64;
  while (tile_rows > 1 && (size_t)tile_rows * total * esz > 48 * 1024) tile_rows >>= 1;
  TRS_UNSUPPORTED((size_t)tile_rows * total * esz > 48 * 1024, "trs_index_concat: %d index columns per row is too wide",
                  total);
  a.ncols = ncols; a.total = total; a.tile_rows = tile_rows; a.batch = batch; a.out = out;
  const int64_t tiles = (batch + tile_rows - 1) / tile_rows;
  const int grid = static_cast<int>(tiles < kNumSMs * 8 ? tiles : kNumSMs * 8);
  const size_t smem = (size_t)tile_rows * total * esz;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (idx_bits == 64) index_concat_kernel<long long><<<grid, 256, smem, s>>>(a);
  else index_concat_kernel<int><<<grid, 256, smem, s>>>(a);
  return check_launch("index_concat_kernel");
}
