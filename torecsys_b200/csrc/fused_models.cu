// a12: fused model forwards, indices -> logits, generic in (N, E, layer sizes).
//
// One kernel per model family gathers the embedding rows of a tile of samples straight into shared memory
// (128-bit no-allocate loads), computes the interaction there and emits one logit per sample; the (B,N,E)
// intermediate of the reference (and its (B,N,N,E)/(B,P,E) cousins) never exists in HBM.
//   fm_family_kernel : FactorizationMachineModel, DeepFactorizationMachineModel, and the non-CIN part of xDeepFM
//   dcn_kernel       : DeepAndCrossNetworkModel
//   ffm_model_kernel : FieldAwareFactorizationMachineModel (pair-parallel, rows used once: no staging at all)
// The Criteo-shaped DeepFM configuration (E = 16, hidden = 16) has a register-resident fast path in
// deepfm_fast.cu; trs_deepfm_forward dispatches to it when the shape matches.
#include "tile_ops.cuh"

namespace trs {

int cin_run(const float* x, const float* const* conv_w, const float* const* scale, const float* const* shift,
            const int* layer_sizes, int layers, int is_direct, int activation, const float* fc_w, const float* fc_b,
            int out_features, int64_t batch, int fields, int embed, float* out, int accumulate, void* workspace,
            int64_t workspace_bytes, cudaStream_t s);
int deepfm_fast_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation, int64_t rows);
int deepfm_fast_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                       const float* w_feat, const float* w_emb, int64_t rows, const float* const* mlp_w,
                       const float* const* mlp_b, int mlp_layers, float* logits, int32_t* status, cudaStream_t s,
                       float* x_out, const float* bias, int use_fm);

int mlp_chain_supported(const int* dims, int layers, int64_t rows, const void* x, const void* out, int accumulate);
int mlp_chain_run(const float* x, int64_t rows, const MlpParams& mp, float* out, int accumulate, cudaStream_t s,
                  const DenseFuse* gather);
int mlp_chain_gather_supported(const int* dims, int layers, int fields, int embed, int use_fm);
int dcn_tc_supported(int embed, int cross_layers, const int* mlp_dims, int mlp_layers, int activation);
int dcn_tc_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                  const float* w_emb, int64_t rows, int embed, const float* cross_w, const float* cross_b,
                  int cross_layers, const MlpParams& mp, const float* fc_w, const float* fc_b, float* logits,
                  int32_t* status, cudaStream_t s);

int dcn_tc5_supported(int embed, int fields, int cross_layers, const int* mlp_dims, int mlp_layers, int activation,
                      int64_t rows);
int dcn_tc5_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                   const float* w_emb, int64_t rows, int embed, const float* cross_w, const float* cross_b,
                   int cross_layers, const MlpParams& mp, const float* fc_w, const float* fc_b, float* logits,
                   int32_t* status, cudaStream_t s);

namespace {

struct FmFamilyArgs {
  const void* idx;
  const int64_t* offsets;
  const float* w_feat;  // may be null (no first-order term)
  const float* w_emb;
  const float* bias;    // may be null
  float* x_out;         // may be null; (B, N, E) copy of the gathered rows (xDeepFM hands it to CIN)
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, embed;
  int use_fm;
  int ts, pitch, hpitch;
  MlpParams mp;
};

template <int IdxBits>
__global__ void __launch_bounds__(256) fm_family_kernel(FmFamilyArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;
  float* buf0 = tile + (size_t)a.ts * a.pitch;
  float* buf1 = buf0 + (size_t)a.ts * a.hpitch;
  float* part = buf1 + (size_t)a.ts * a.hpitch;  // [ts] first-order + FM + bias
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int width = a.fields * a.embed;
  for (int64_t b0 = (int64_t)blockIdx.x * a.ts; b0 < a.batch; b0 += (int64_t)gridDim.x * a.ts) {
    const int valid = static_cast<int>(a.batch - b0 < a.ts ? a.batch - b0 : a.ts);
    __syncthreads();
    gather_tile<IdxBits>(a.w_emb, a.rows, a.embed, a.idx, a.offsets, b0, a.ts, valid, a.fields, tile, a.pitch,
                         a.embed, a.status);
    // first-order term (one warp per sample, lanes over fields) while the row loads are in flight
    for (int s = warp; s < a.ts; s += warps) {
      float first = 0.f;
      if (a.w_feat != nullptr && s < valid) {
        for (int n = lane; n < a.fields; n += 32) {
          const int64_t pos = (b0 + s) * a.fields + n;
          const int64_t r = load_index<IdxBits>(a.idx, pos) + __ldg(a.offsets + n);
          if (r >= 0 && r < a.rows) first += ldg_stream_f1(a.w_feat + r);
        }
        first = warp_sum(first);
      }
      if (lane == 0) part[s] = first + (a.bias ? __ldg(a.bias) : 0.f);
    }
    __syncthreads();
    if (a.use_fm) {
      for (int s = warp; s < valid; s += warps) {
        float fm = 0.f;
        for (int e = lane; e < a.embed; e += 32) {
          float sum = 0.f, sq = 0.f;
          const float* col = tile + s * a.pitch + e;
          for (int n = 0; n < a.fields; ++n) {
            const float v = col[n * a.embed];
            sum += v;
            sq = fmaf(v, v, sq);
          }
          fm += 0.5f * (sum * sum - sq);
        }
        fm = warp_sum(fm);
        if (lane == 0) part[s] += fm;
      }
    }
    if (a.x_out != nullptr) {
      for (int t = threadIdx.x; t < valid * width; t += blockDim.x) {
        const int s = t / width, c = t - s * width;
        a.x_out[(b0 + s) * width + c] = tile[s * a.pitch + c];
      }
    }
    const float* res = nullptr;
    if (a.mp.layers > 0) res = mlp_tile(a.mp, tile, a.pitch, a.ts, buf0, buf1, a.hpitch);
    else __syncthreads();
    for (int s = threadIdx.x; s < valid; s += blockDim.x)
      a.logits[b0 + s] = part[s] + (res ? res[s * a.hpitch] : 0.f);
  }
}

// -------------------------------------------------------------------------------------------------------- DCN
struct DcnArgs {
  const void* idx;
  const int64_t* offsets;
  const float* w_emb;
  const float* cross_w;
  const float* cross_b;
  const float* fc_w;
  const float* fc_b;
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, embed, cross_layers;
  int ts, pe, bp;  // samples per tile, row pitch of the x tile, row pitch of the two work buffers
  MlpParams mp;
};

template <int IdxBits>
__global__ void __launch_bounds__(256) dcn_kernel(DcnArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int nrows = a.ts * a.fields;  // (sample, field) rows in the tile
  float* xs = smem;                               // [nrows][pe]
  float* bufa = xs + (size_t)nrows * a.pe;        // [nrows][bp]
  float* bufb = bufa + (size_t)nrows * a.bp;      // [nrows][bp]
  float* part = bufb + (size_t)nrows * a.bp;      // [ts]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int od = a.mp.dims[a.mp.layers];
  const int cat = a.embed + od;
  for (int64_t b0 = (int64_t)blockIdx.x * a.ts; b0 < a.batch; b0 += (int64_t)gridDim.x * a.ts) {
    const int valid = static_cast<int>(a.batch - b0 < a.ts ? a.batch - b0 : a.ts);
    __syncthreads();
    gather_tile<IdxBits>(a.w_emb, a.rows, a.embed, a.idx, a.offsets, b0, a.ts, valid, a.fields, xs,
                         a.fields * a.pe, a.pe, a.status);
    __syncthreads();
    // cross network: h <- x * (W h + b) + x
    const float* cur = xs;
    int cur_pitch = a.pe;
    float* dst = bufa;
    for (int l = 0; l < a.cross_layers; ++l) {
      float* d = dst;
      const int bp = a.bp, pe = a.pe;
      dense_layer_tile(cur, cur_pitch, a.embed, a.cross_w + (int64_t)l * a.embed * a.embed,
                       a.cross_b + (int64_t)l * a.embed, a.embed, nrows, [=](int r, int o, float v) {
                         const float x0 = xs[r * pe + o];
                         d[r * bp + o] = fmaf(x0, v, x0);
                       });
      __syncthreads();
      cur = dst;
      cur_pitch = a.bp;
      dst = (dst == bufa) ? bufb : bufa;
    }
    // fc over the cross half: part[s] = fc_b + sum_{n,e} fc_w[n*cat + e] * cross[s,n,e]
    for (int s = warp; s < a.ts; s += warps) {
      float acc = 0.f;
      for (int t = lane; t < a.fields * a.embed; t += 32) {
        const int n = t / a.embed, e = t - n * a.embed;
        acc = fmaf(__ldg(a.fc_w + n * cat + e), cur[(s * a.fields + n) * cur_pitch + e], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) part[s] = acc + __ldg(a.fc_b);
    }
    __syncthreads();
    // per-field MLP on the same rows, then the deep half of fc
    const float* res = mlp_tile(a.mp, xs, a.pe, nrows, bufa, bufb, a.bp);
    for (int s = warp; s < valid; s += warps) {
      float acc = 0.f;
      for (int t = lane; t < a.fields * od; t += 32) {
        const int n = t / od, o = t - n * od;
        acc = fmaf(__ldg(a.fc_w + n * cat + a.embed + o), res[(s * a.fields + n) * a.bp + o], acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) a.logits[b0 + s] = part[s] + acc;
    }
  }
}

// -------------------------------------------------------------------------------------------------------- FFM model
// One warp per sample.  Work item = (pair, 16-byte chunk): two random 128-bit loads T_i[r_j], T_j[r_i], 4 FMAs.
// Each lane keeps 4 items = 8 loads in flight; 8 CTAs x 8 warps per SM.
constexpr int kFfmUnroll = 4;

template <int IdxBits, bool Vec>
__global__ void __launch_bounds__(256) ffm_model_kernel(const void* __restrict__ idx,
                                                        const int64_t* __restrict__ offsets, int64_t batch,
                                                        int fields, const float* __restrict__ w_feat,
                                                        const float* const* __restrict__ tables, int64_t rows,
                                                        int embed, const float* __restrict__ bias,
                                                        float* __restrict__ logits, int32_t* status,
                                                        const int* __restrict__ pair_list, int n_pairs,
                                                        int64_t first_begin, int64_t first_end) {
  // pair_list == null: all N(N-1)/2 pairs.  Otherwise only the listed pairs ((i << 16) | j) are summed -- the share of
  // one rank in the owner-side sharded scheme -- and the first-order term + bias are added only for the samples in
  // [first_begin, first_end) (each sample's first-order term must enter the cross-rank sum exactly once).
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int all_pairs = fields * (fields - 1) / 2;
  const int pairs = pair_list ? n_pairs : all_pairs;
  int* ptab = reinterpret_cast<int*>(smem_raw);                                   // [all_pairs]
  const float** tabs = reinterpret_cast<const float**>(ptab + ((all_pairs + 1) & ~1));  // [fields]
  int64_t* rid = reinterpret_cast<int64_t*>(tabs + fields) + (size_t)warp * fields;  // [warps][fields]
  for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
    if (pair_list) {
      ptab[p] = __ldg(pair_list + p);
    } else {
      int i, j;
      pair_from_index(p, fields, i, j);
      ptab[p] = (i << 16) | j;
    }
  }
  for (int t = threadIdx.x; t < fields; t += blockDim.x) tabs[t] = tables[t];
  __syncthreads();
  constexpr int W = Vec ? 4 : 1;
  const int chunks = embed / W;
  const int items = pairs * chunks;
  for (int64_t b = (int64_t)blockIdx.x * warps + warp; b < batch; b += (int64_t)gridDim.x * warps) {
    float first = 0.f;
    const bool with_first = b >= first_begin && b < first_end;
    for (int n = lane; n < fields; n += 32) {
      const int64_t pos = b * fields + n;
      int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + n);
      if (r < 0 || r >= rows) {
        if (with_first) report_oob(status, pos);   // reported once, by the rank that owns the sample
        r = -1;
      } else if (w_feat != nullptr && with_first) {
        first += ldg_stream_f1(w_feat + r);
      }
      rid[n] = r;
    }
    __syncwarp();
    float acc = 0.f;
    for (int base = lane; base < items; base += 32 * kFfmUnroll) {
      const float* pa[kFfmUnroll];
      const float* pb[kFfmUnroll];
#pragma unroll
      for (int u = 0; u < kFfmUnroll; ++u) {
        const int item = base + u * 32;
        pa[u] = nullptr;
        pb[u] = nullptr;
        if (item < items) {
          const int p = item / chunks, c = item - p * chunks;
          const int ij = ptab[p];
          const int i = ij >> 16, j = ij & 0xffff;
          const int64_t ri = rid[i], rj = rid[j];
          if (ri >= 0 && rj >= 0) {
            pa[u] = tabs[i] + rj * embed + c * W;  // field-aware row (table i, feature j)
            pb[u] = tabs[j] + ri * embed + c * W;  // (table j, feature i)
          }
        }
      }
      if (Vec) {
        float4 va[kFfmUnroll], vb[kFfmUnroll];
#pragma unroll
        for (int u = 0; u < kFfmUnroll; ++u) {
          va[u] = pa[u] ? ldg_stream_f4(reinterpret_cast<const float4*>(pa[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
          vb[u] = pb[u] ? ldg_stream_f4(reinterpret_cast<const float4*>(pb[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kFfmUnroll; ++u) {
          acc = fmaf(va[u].x, vb[u].x, acc);
          acc = fmaf(va[u].y, vb[u].y, acc);
          acc = fmaf(va[u].z, vb[u].z, acc);
          acc = fmaf(va[u].w, vb[u].w, acc);
        }
      } else {
        float va[kFfmUnroll], vb[kFfmUnroll];
#pragma unroll
        for (int u = 0; u < kFfmUnroll; ++u) {
          va[u] = pa[u] ? ldg_stream_f1(pa[u]) : 0.f;
          vb[u] = pb[u] ? ldg_stream_f1(pb[u]) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kFfmUnroll; ++u) acc = fmaf(va[u], vb[u], acc);
      }
    }
    const float tot = warp_sum(acc + first);
    if (lane == 0) logits[b] = tot + ((bias && with_first) ? __ldg(bias) : 0.f);
    __syncwarp();
  }
}

int launch_fm_family(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                     const float* w_feat, const float* w_emb, int64_t rows, int embed, int use_fm,
                     const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                     int activation, const float* bias, float* x_out, float* logits, int32_t* status,
                     cudaStream_t s, const char* who) {
  TRS_REQUIRE(idx && offsets && w_emb && logits, "%s: null pointer", who);
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "%s: idx_bits must be 32 or 64", who);
  TRS_REQUIRE(batch >= 0 && fields > 0 && embed > 0 && rows > 0, "%s: bad sizes", who);
  FmFamilyArgs a{};
  if (mlp_layers > 0) {
    TRS_REQUIRE(mlp_dims && mlp_w, "%s: null MLP description", who);
    TRS_REQUIRE(fill_mlp_params(a.mp, mlp_dims, mlp_layers, mlp_w, mlp_b, activation) == 0,
                "%s: bad MLP description (at most %d layers)", who, MlpParams::kMaxLayers);
    TRS_REQUIRE(mlp_dims[0] == fields * embed, "%s: MLP input size %d != fields*embed %d", who, mlp_dims[0],
                fields * embed);
    TRS_REQUIRE(mlp_dims[mlp_layers] == 1, "%s: MLP output size must be 1", who);
  } else {
    a.mp.layers = 0;
  }
  if (batch == 0) return TRS_OK;
  // Criteo-shaped deep branch (E = 16, hidden widths 16, ReLU): the register-resident mma.sync kernel of deepfm_fast.cu,
  // with or without the FM term, the rows written out for a consumer behind it (xDeepFM: 446 -> ~200 us per 65 536)
  if (mlp_layers > 0 && w_feat != nullptr && aligned16(w_emb) && aligned16(mlp_w[0]) && (x_out == nullptr || aligned16(x_out)) &&
      deepfm_fast_supported(fields, embed, mlp_dims, mlp_layers, activation, rows)) {
    bool all = true;
    for (int l = 0; l < mlp_layers; ++l) all = all && mlp_w[l] != nullptr && mlp_b != nullptr && mlp_b[l] != nullptr;
    if (all)
      return deepfm_fast_launch(idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, mlp_w, mlp_b, mlp_layers,
                                logits, status, s, x_out, bias, use_fm);
  }
  a.idx = idx; a.offsets = offsets; a.w_feat = w_feat; a.w_emb = w_emb; a.bias = bias; a.x_out = x_out;
  a.logits = logits; a.status = status; a.batch = batch; a.rows = rows; a.fields = fields; a.embed = embed;
  a.use_fm = use_fm;
  // wide deep branch: gather + first-order + FM in this kernel (rows written out once), then the MLP as a chain of
  // tensor-core layers (dense.cu / cin_tc.cu) accumulating into the logits
  const MlpParams chain = a.mp;
  float* x_scratch = nullptr;
  bool chained = false;
  if (mlp_layers > 0 && mlp_chain_supported(mlp_dims, mlp_layers, batch, logits, logits, 1)) {
    // Fully fused form: the first tensor-core layer gathers the rows itself and emits first-order + FM per
    // sample, the last hidden layer carries the logit Linear in its epilogue -- no gather kernel, no (B, N*E) matrix.
    if (a.x_out == nullptr && mlp_chain_gather_supported(mlp_dims, mlp_layers, fields, embed, use_fm)) {
      DenseFuse g;
      g.idx = idx; g.idx_bits = idx_bits; g.offsets = offsets; g.table = w_emb; g.w_feat = w_feat; g.bias = bias;
      g.status = status; g.table_rows = rows; g.fields = fields; g.embed = embed; g.use_fm = use_fm;
      return mlp_chain_run(nullptr, batch, chain, logits, 1, s, &g);
    }
    chained = true;
    a.mp.layers = 0;
    mlp_layers = 0;
    if (a.x_out == nullptr) {
      TRS_CUDA(scratch_alloc(reinterpret_cast<void**>(&x_scratch), (size_t)batch * fields * embed * sizeof(float), s));
      a.x_out = x_scratch;
    }
  }
  a.pitch = tile_pitch(fields * embed);
  a.hpitch = mlp_layers > 0 ? tile_pitch(mlp_max_hidden(mlp_dims, mlp_layers)) : 4;
  int ts = 64;
  size_t smem;
  for (;; ts >>= 1) {
    smem = ((size_t)ts * a.pitch + 2 * (size_t)ts * a.hpitch + ts) * sizeof(float);
    if (smem <= 100 * 1024 || ts == 1) break;  // two CTAs per SM when possible
  }
  if (smem > (size_t)kMaxDynSmem) {
    if (x_scratch) cudaFreeAsync(x_scratch, s);
    TRS_UNSUPPORTED(true, "%s: fields*embed / MLP widths do not fit shared memory", who);
  }
  a.ts = ts;
  const int64_t tiles = (batch + ts - 1) / ts;
  const int grid = static_cast<int>(tiles < kNumSMs * 2 ? tiles : kNumSMs * 2);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN(fm_family_kernel<64>);
    fm_family_kernel<64><<<grid, 256, smem, s>>>(a);
  } else {
    TRS_SMEM_OPT_IN(fm_family_kernel<32>);
    fm_family_kernel<32><<<grid, 256, smem, s>>>(a);
  }
  int rc = check_launch("fm_family_kernel");
  if (chained) {
    if (rc == TRS_OK) rc = mlp_chain_run(a.x_out, batch, chain, logits, 1, s, nullptr);
    if (x_scratch) cudaFreeAsync(x_scratch, s);
  }
  return rc;
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_fm_model_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                    const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                    const float* bias, float* logits, int32_t* status, void* stream) {
  TRS_REQUIRE(w_feat, "trs_fm_model_forward: null w_feat");
  return launch_fm_family(idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, embed, 1, nullptr, 0, nullptr,
                          nullptr, TRS_ACT_NONE, bias, nullptr, logits, status, static_cast<cudaStream_t>(stream),
                          "trs_fm_model_forward");
}

extern "C" int trs_deepfm_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                  const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                  const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                  const float* const* mlp_b, int activation, float* logits, int32_t* status,
                                  void* stream) {
  TRS_REQUIRE(w_feat, "trs_deepfm_forward: null w_feat");
  TRS_REQUIRE(mlp_layers >= 1, "trs_deepfm_forward: the deep part needs at least the output layer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (idx && offsets && w_emb && logits && mlp_dims && mlp_w && mlp_b && batch > 0 &&
      deepfm_fast_supported(fields, embed, mlp_dims, mlp_layers, activation, rows)) {
    return deepfm_fast_launch(idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, mlp_w, mlp_b, mlp_layers,
                              logits, status, s, nullptr, nullptr, 1);
  }
  return launch_fm_family(idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, embed, 1, mlp_dims, mlp_layers,
                          mlp_w, mlp_b, activation, nullptr, nullptr, logits, status, s, "trs_deepfm_forward");
}

extern "C" int trs_dcn_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                               const float* w_emb, int64_t rows, int embed, const float* cross_w,
                               const float* cross_b, int cross_layers, const int* mlp_dims, int mlp_layers,
                               const float* const* mlp_w, const float* const* mlp_b, int activation,
                               const float* fc_w, const float* fc_b, float* logits, int32_t* status, void* stream) {
  TRS_REQUIRE(idx && offsets && w_emb && fc_w && fc_b && logits && mlp_dims && mlp_w,
              "trs_dcn_forward: null pointer");
  TRS_REQUIRE(cross_layers == 0 || (cross_w && cross_b), "trs_dcn_forward: null cross parameters");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_dcn_forward: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && embed > 0 && rows > 0 && cross_layers >= 0 && mlp_layers >= 1,
              "trs_dcn_forward: bad sizes");
  DcnArgs a{};
  TRS_REQUIRE(fill_mlp_params(a.mp, mlp_dims, mlp_layers, mlp_w, mlp_b, activation) == 0,
              "trs_dcn_forward: bad MLP description (at most %d layers)", MlpParams::kMaxLayers);
  TRS_REQUIRE(mlp_dims[0] == embed, "trs_dcn_forward: the per-field MLP input size must be embed");
  if (batch == 0) return TRS_OK;
  // the per-row dense chains in tensor memory on tcgen05 (dcn_tc5.cu) where the shape allows and the batch fills the SMs,
  // else on 3xTF32 mma.sync (dcn_tc.cu)
  if (batch * fields >= 128 * 4 && dcn_tc5_supported(embed, fields, cross_layers, mlp_dims, mlp_layers, activation, rows))
    return dcn_tc5_launch(idx, idx_bits, offsets, batch, fields, w_emb, rows, embed, cross_w, cross_b, cross_layers,
                          a.mp, fc_w, fc_b, logits, status, static_cast<cudaStream_t>(stream));
  if (dcn_tc_supported(embed, cross_layers, mlp_dims, mlp_layers, activation) && (fields * 16) % 16 == 0)
    return dcn_tc_launch(idx, idx_bits, offsets, batch, fields, w_emb, rows, embed, cross_w, cross_b, cross_layers,
                         a.mp, fc_w, fc_b, logits, status, static_cast<cudaStream_t>(stream));
  a.idx = idx; a.offsets = offsets; a.w_emb = w_emb; a.cross_w = cross_w; a.cross_b = cross_b; a.fc_w = fc_w;
  a.fc_b = fc_b; a.logits = logits; a.status = status; a.batch = batch; a.rows = rows; a.fields = fields;
  a.embed = embed; a.cross_layers = cross_layers;
  a.pe = tile_pitch(embed);
  const int hmax = mlp_max_hidden(mlp_dims, mlp_layers);
  a.bp = tile_pitch(hmax > embed ? hmax : embed);
  int ts = 16;
  size_t smem;
  for (;; ts >>= 1) {
    smem = ((size_t)ts * fields * (a.pe + 2 * a.bp) + ts) * sizeof(float);
    if (smem <= (size_t)kMaxDynSmem - 2048 || ts == 1) break;
  }
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_dcn_forward: fields*embed tile does not fit shared memory");
  a.ts = ts;
  const int64_t tiles = (batch + ts - 1) / ts;
  const int grid = static_cast<int>(tiles < kNumSMs * 4 ? tiles : kNumSMs * 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN(dcn_kernel<64>);
    dcn_kernel<64><<<grid, 256, smem, s>>>(a);
  } else {
    TRS_SMEM_OPT_IN(dcn_kernel<32>);
    dcn_kernel<32><<<grid, 256, smem, s>>>(a);
  }
  return check_launch("dcn_kernel");
}

extern "C" int64_t trs_xdeepfm_workspace_bytes(int64_t batch, int fields, int embed, const int* cin_layer_sizes,
                                               int cin_layers, int cin_is_direct) {
  const int64_t cin = trs_cin_workspace_bytes(batch, fields, embed, cin_layer_sizes, cin_layers, cin_is_direct);
  if (cin < 0) return cin;
  const int64_t x_bytes = ((batch * fields * embed * (int64_t)sizeof(float) + 255) / 256) * 256;
  return x_bytes + cin;
}

extern "C" int trs_xdeepfm_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                   const float* w_feat, const float* w_emb, int64_t rows, int embed,
                                   const float* const* cin_w, const float* const* cin_scale,
                                   const float* const* cin_shift, const int* cin_layer_sizes, int cin_layers,
                                   int cin_is_direct, int cin_activation, const float* cin_fc_w,
                                   const float* cin_fc_b, const int* mlp_dims, int mlp_layers,
                                   const float* const* mlp_w, const float* const* mlp_b, int mlp_activation,
                                   const float* bias, float* logits, void* workspace, int64_t workspace_bytes,
                                   int32_t* status, void* stream) {
  TRS_REQUIRE(w_feat && workspace, "trs_xdeepfm_forward: null pointer");
  const int64_t need = trs_xdeepfm_workspace_bytes(batch, fields, embed, cin_layer_sizes, cin_layers, cin_is_direct);
  TRS_REQUIRE(need >= 0 && workspace_bytes >= need, "trs_xdeepfm_forward: workspace too small");
  TRS_REQUIRE(aligned16(workspace), "trs_xdeepfm_forward: workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* x = static_cast<float*>(workspace);
  const int64_t x_bytes = ((batch * fields * embed * (int64_t)sizeof(float) + 255) / 256) * 256;
  // stage 1: gather once; logits = sum_n w_feat + MLP(flatten(emb)) + bias; the gathered rows are kept for CIN
  int rc = launch_fm_family(idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, embed, 0, mlp_dims,
                            mlp_layers, mlp_w, mlp_b, mlp_activation, bias, x, logits, status, s,
                            "trs_xdeepfm_forward");
  if (rc != TRS_OK) return rc;
  // stage 2: logits += CIN(x)
  return cin_run(x, cin_w, cin_scale, cin_shift, cin_layer_sizes, cin_layers, cin_is_direct, cin_activation,
                 cin_fc_w, cin_fc_b, 1, batch, fields, embed, logits, 1,
                 static_cast<unsigned char*>(workspace) + x_bytes, workspace_bytes - x_bytes, s);
}

extern "C" int trs_ffm_model_forward(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                     int fields, const float* w_feat, const float* const* tables, int64_t rows,
                                     int embed, const float* bias, float* logits, int32_t* status, void* stream) {
  return trs_ffm_model_forward_pairs(idx, idx_bits, offsets, batch, fields, w_feat, tables, rows, embed, bias, nullptr,
                                     0, 0, batch, logits, status, stream);
}

extern "C" int trs_ffm_model_forward_pairs(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                           int fields, const float* w_feat, const float* const* tables, int64_t rows,
                                           int embed, const float* bias, const int* pair_list, int n_pairs,
                                           int64_t first_begin, int64_t first_end, float* logits, int32_t* status,
                                           void* stream) {
  TRS_REQUIRE(idx && offsets && tables && logits, "trs_ffm_model_forward: null pointer");
  TRS_REQUIRE(pair_list == nullptr || (n_pairs >= 0 && n_pairs <= fields * (fields - 1) / 2),
              "trs_ffm_model_forward_pairs: bad pair list length %d", n_pairs);
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_ffm_model_forward: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 1 && embed > 0 && rows > 0, "trs_ffm_model_forward: bad sizes");
  TRS_REQUIRE(fields <= 4096, "trs_ffm_model_forward: too many fields");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int pairs = fields * (fields - 1) / 2;
  const int warps = 8;
  const size_t smem = (size_t)((pairs + 1) & ~1) * sizeof(int) + (size_t)fields * sizeof(float*) +
                      (size_t)warps * fields * sizeof(int64_t);
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "trs_ffm_model_forward: too many fields for shared memory");
  const bool vec = embed % 4 == 0;
  const int grid = grid_for(batch * 32, warps * 32, 8);
#define LAUNCH(BITS, VEC)                                                                                        \
  do {                                                                                                           \
    TRS_SMEM_OPT_IN((ffm_model_kernel<BITS, VEC>));                                                              \
    ffm_model_kernel<BITS, VEC><<<grid, warps * 32, smem, s>>>(idx, offsets, batch, fields, w_feat, tables, rows, \
                                                               embed, bias, logits, status, pair_list, n_pairs,  \
                                                               first_begin, first_end);                          \
  } while (0)
  if (idx_bits == 64) {
    if (vec) LAUNCH(64, true); else LAUNCH(64, false);
  } else {
    if (vec) LAUNCH(32, true); else LAUNCH(32, false);
  }
#undef LAUNCH
  return check_launch("ffm_model_kernel");
}
