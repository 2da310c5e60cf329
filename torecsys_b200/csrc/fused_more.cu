// Fused indices -> logits forwards of three of the SURVEY.md 8f-3 models: one kernel gathers the embedding rows of a
// tile of samples into shared memory, builds the model's MLP input there and runs the MLP on the tile -- neither the
// (B,N,E) lookup nor the (B,P) inner products nor the concatenated MLP input exist in HBM.
//
//   trs_nfm_forward        NeuralFactorizationMachineModel.forward (models/ctr/neural_factorization_machine.py:66-96)
//                          logit = MLP( FM(emb) ) + sum_n feat (+ bias)
//   trs_fnn_forward        FactorizationMachineSupportedNeuralNetworkModel.forward
//                          (models/ctr/factorization_machine_supported_neural_network.py:61-101)
//                          logit = MLP( cat[ feat (B,N), FM(emb) (B,E) ] )
//   trs_pnn_inner_forward  ProductNeuralNetworkModel.forward, prod_method='inner'
//                          (models/ctr/product_neural_network.py:81-115)
//                          logit = MLP( cat[ IPN(emb) (B,NC2), feat (B,N), bias (B,1) ] )
// each behind Sequential.forward = Inputs.forward (inputs/inputs.py:56-89) + the model forward.
#include "tile_ops.cuh"

namespace trs {
namespace {

enum FeatMode { kNfm = 0, kFnn = 1, kPnnInner = 2 };

struct FeatArgs {
  const void* idx;
  const int64_t* offsets;
  const float* w_feat;
  const float* w_emb;
  const float* bias;   // may be null
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, embed, mode;
  int ts, pitch, row_pitch, fpitch, hpitch;
  MlpParams mp;
};

template <int IdxBits>
__global__ void __launch_bounds__(256) feature_mlp_kernel(FeatArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                                  // [ts][pitch]: field n of sample s at s*pitch + n*row_pitch
  float* feat = tile + (size_t)a.ts * a.pitch;         // [ts][fpitch]: the MLP input
  float* buf0 = feat + (size_t)a.ts * a.fpitch;
  float* buf1 = buf0 + (size_t)a.ts * a.hpitch;
  float* part = buf1 + (size_t)a.ts * a.hpitch;        // [ts]: terms added to the MLP output (NFM)
  int* ptab = reinterpret_cast<int*>(part + a.ts);     // [pairs] (PNN): (i << 16) | j
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int n_f = a.fields, e_dim = a.embed;
  const int pairs = n_f * (n_f - 1) / 2;
  if (a.mode == kPnnInner) {
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
      int i, j;
      pair_from_index(p, n_f, i, j);
      ptab[p] = (i << 16) | j;
    }
  }
  const int feat_off = a.mode == kFnn ? 0 : pairs;     // where the first-order values go in the MLP input
  const int fm_off = a.mode == kFnn ? n_f : 0;         // where the FM vector goes
  for (int64_t b0 = (int64_t)blockIdx.x * a.ts; b0 < a.batch; b0 += (int64_t)gridDim.x * a.ts) {
    const int valid = static_cast<int>(a.batch - b0 < a.ts ? a.batch - b0 : a.ts);
    __syncthreads();
    gather_tile<IdxBits>(a.w_emb, a.rows, e_dim, a.idx, a.offsets, b0, a.ts, valid, n_f, tile, a.pitch, a.row_pitch,
                         a.status);
    // first-order lookups (one warp per sample, lanes over fields) while the row loads are in flight
    for (int s = warp; s < a.ts; s += warps) {
      float first = 0.f;
      for (int n = lane; n < n_f; n += 32) {
        float f = 0.f;
        if (s < valid) {
          const int64_t pos = (b0 + s) * n_f + n;
          const int64_t r = load_index<IdxBits>(a.idx, pos) + __ldg(a.offsets + n);
          if (r >= 0 && r < a.rows) f = ldg_stream_f1(a.w_feat + r);
        }
        if (a.mode == kNfm) first += f;
        else feat[s * a.fpitch + feat_off + n] = f;
      }
      if (a.mode == kNfm) {
        first = warp_sum(first);
        if (lane == 0) part[s] = first + (a.bias ? __ldg(a.bias) : 0.f);
      } else if (lane == 0) {
        part[s] = 0.f;
        if (a.mode == kPnnInner) feat[s * a.fpitch + pairs + n_f] = a.bias ? __ldg(a.bias) : 0.f;
      }
    }
    __syncthreads();
    if (a.mode == kPnnInner) {
      // inner products: warp per sample, lane <-> pair (consecutive pairs share i, consecutive j: row_pitch = E + 4
      // spreads them over the banks)
      for (int s = warp; s < a.ts; s += warps) {
        const float* xs = tile + s * a.pitch;
        for (int p = lane; p < pairs; p += 32) {
          const int ij = ptab[p];
          const float* xi = xs + (ij >> 16) * a.row_pitch;
          const float* xj = xs + (ij & 0xffff) * a.row_pitch;
          float acc = 0.f;
#pragma unroll 4
          for (int e = 0; e < e_dim; ++e) acc = fmaf(xi[e], xj[e], acc);
          feat[s * a.fpitch + p] = acc;
        }
      }
    } else {
      // FM vector: warp per sample, lanes over embedding components
      for (int s = warp; s < a.ts; s += warps) {
        for (int e = lane; e < e_dim; e += 32) {
          float sum = 0.f, sq = 0.f;
          const float* col = tile + s * a.pitch + e;
          for (int n = 0; n < n_f; ++n) {
            const float v = col[n * a.row_pitch];
            sum += v;
            sq = fmaf(v, v, sq);
          }
          feat[s * a.fpitch + fm_off + e] = 0.5f * (sum * sum - sq);
        }
      }
    }
    __syncthreads();
    const float* res = mlp_tile(a.mp, feat, a.fpitch, a.ts, buf0, buf1, a.hpitch);
    for (int s = threadIdx.x; s < valid; s += blockDim.x) a.logits[b0 + s] = part[s] + res[s * a.hpitch];
  }
}

int launch_feature_mlp(int mode, const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                       const float* w_feat, const float* w_emb, int64_t rows, int embed, const int* mlp_dims,
                       int mlp_layers, const float* const* mlp_w, const float* const* mlp_b, int activation,
                       const float* bias, float* logits, int32_t* status, cudaStream_t s, const char* who) {
  TRS_REQUIRE(idx && offsets && w_feat && w_emb && logits && mlp_dims && mlp_w && mlp_b, "%s: null pointer", who);
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "%s: idx_bits must be 32 or 64", who);
  TRS_REQUIRE(batch >= 0 && fields > 0 && (mode != kPnnInner || fields > 1) && fields < 65536 && rows > 0 && embed > 0 &&
                  mlp_layers >= 1,
              "%s: bad sizes", who);
  TRS_REQUIRE(activation >= TRS_ACT_NONE && activation <= TRS_ACT_TANH, "%s: unknown activation", who);
  const int pairs = fields * (fields - 1) / 2;
  const int width = mode == kNfm ? embed : mode == kFnn ? fields + embed : pairs + fields + 1;
  TRS_REQUIRE(mlp_dims[0] == width && mlp_dims[mlp_layers] == 1,
              "%s: the MLP must map %d inputs to 1 output (got %d -> %d)", who, width, mlp_dims[0],
              mlp_dims[mlp_layers]);
  if (batch == 0) return TRS_OK;
  FeatArgs a{};
  TRS_REQUIRE(fill_mlp_params(a.mp, mlp_dims, mlp_layers, mlp_w, mlp_b, activation) == 0,
              "%s: bad MLP description (at most %d layers)", who, MlpParams::kMaxLayers);
  a.idx = idx; a.offsets = offsets; a.w_feat = w_feat; a.w_emb = w_emb; a.bias = bias; a.logits = logits;
  a.status = status; a.batch = batch; a.rows = rows; a.fields = fields; a.embed = embed; a.mode = mode;
  // the vector path of gather_tile needs 16-byte aligned rows in the tile: embed % 4 == 0 -> pitch E + 4 for PNN
  a.row_pitch = (mode == kPnnInner && (embed & 3) == 0) ? embed + 4 : (mode == kPnnInner ? embed | 1 : embed);
  a.pitch = tile_pitch(fields * a.row_pitch);
  a.fpitch = tile_pitch(width);
  a.hpitch = tile_pitch(mlp_max_hidden(mlp_dims, mlp_layers));
  int ts = 64;
  size_t smem;
  for (;; ts >>= 1) {
    smem = ((size_t)ts * (a.pitch + a.fpitch + 2 * a.hpitch + 1) + (mode == kPnnInner ? pairs : 0)) * sizeof(float);
    if (smem <= 100 * 1024 || ts == 1) break;   // two CTAs per SM when possible
  }
  TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "%s: fields*embed / MLP widths do not fit shared memory", who);
  a.ts = ts;
  const int64_t tiles = (batch + ts - 1) / ts;
  const int grid = static_cast<int>(tiles < kNumSMs * 2 ? tiles : kNumSMs * 2);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN(feature_mlp_kernel<64>);
    feature_mlp_kernel<64><<<grid, 256, smem, s>>>(a);
  } else {
    TRS_SMEM_OPT_IN(feature_mlp_kernel<32>);
    feature_mlp_kernel<32><<<grid, 256, smem, s>>>(a);
  }
  return check_launch("feature_mlp_kernel");
}

}  // namespace
}  // namespace trs

using namespace trs;

#define TRS_FEATURE_MODEL(NAME, MODE, BIAS)                                                                          \
  extern "C" int NAME(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,             \
                      const float* w_feat, const float* w_emb, int64_t rows, int embed, const int* mlp_dims,         \
                      int mlp_layers, const float* const* mlp_w, const float* const* mlp_b, int activation,         \
                      const float* bias, float* logits, int32_t* status, void* stream) {                            \
    return launch_feature_mlp(MODE, idx, idx_bits, offsets, batch, fields, w_feat, w_emb, rows, embed, mlp_dims,      \
                              mlp_layers, mlp_w, mlp_b, activation, BIAS, logits, status,                            \
                              static_cast<cudaStream_t>(stream), #NAME);                                             \
  }

TRS_FEATURE_MODEL(trs_nfm_forward, kNfm, bias)
TRS_FEATURE_MODEL(trs_fnn_forward, kFnn, nullptr)
TRS_FEATURE_MODEL(trs_pnn_inner_forward, kPnnInner, bias)
