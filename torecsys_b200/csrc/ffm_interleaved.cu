// FieldAwareFactorizationMachineModel forward on an INTERLEAVED shadow of the field-aware tables: the B200-first memory
// layout of configs[4] (torecsys/models/ctr/field_aware_factorization_machine.py:39-81 over
// torecsys/inputs/base/multi_indices_field_aware_emb.py:90-111).
//
// The reference keeps one (R, E) table per field t and a sample reads T_t[r_f] for every (t, f): N*(N-1) random
// 64-byte rows, and random rows cost a 128-byte DRAM transaction each at a bounded request rate.
// But the N rows T_0[r] .. T_{N-1}[r] of ONE row id r are always wanted together (field_emb[b, t*N+f] = T_t[r_f]
// for all t), so the shadow stores them next to each other, with the first-order weight of the same row id behind:
//     packed[r] = [ T_0[r][0..E) | T_1[r][0..E) | ... | T_{N-1}[r][0..E) | w_feat[r] | 0 .. ]      pitch = 128 B multiple
// A sample then needs N contiguous chunks of N*E*4 + 4 bytes (2.5 KB at N = 39, E = 16) instead of N*(N-1) scattered
// rows, and each chunk is ONE bulk copy of the TMA unit (cp.async.bulk global -> shared, completion on an mbarrier).
//
// Kernel: persistent CTA per SM, a ring of sample-sized stages in shared memory (2 at 39 x 16, up to 8 for narrower
// rows).  Warps 8-15 are producers: each resolves the row ids of ITS fields (f = warp, warp + 8, ..) two samples ahead,
// waits for the stage to be released and issues its bulk copies -- a warp-wide cp.async.bulk is issued lane by lane
// (~60 cycles each), so one producer warp bounds narrow rows at ~4 000 cycles per sample (measured: 4 columns per table
// 1.82 ms per 131 072 samples); producer 0 also writes the finished sample's logit (fixed-order sum of the eight
// consumer partials + bias).  Warps 0-7 consume: with A_f = the chunk of
// field f, logit = sum_{i<j} <A_j[i], A_i[j]> + sum_f A_f[N*E]; work item = (pair, 16-byte piece), both operands
// read from shared memory with 128-bit loads (chunk pitch = 64 mod 128 bytes: conflict-free across the pairs of a warp).
#include "tc5.cuh"

namespace trs {
namespace {

using tc5::bulk_g2s;
using tc5::mbar_arrive;
using tc5::mbar_expect_tx;
using tc5::mbar_init;
using tc5::mbar_wait;
using tc5::smem_u32;

constexpr int kConsumers = 8;
constexpr int kProducers = 8;
constexpr int kMaxStages = 8;
constexpr int kThreads = (kConsumers + kProducers) * 32;

struct Layout {
  int pitch_floats;   // row-id pitch of the packed table (multiple of 32 floats)
  int copy_bytes;     // bytes one bulk copy moves: the N rows + the first-order weight, rounded up to 16
  int stage_pitch;    // floats between two chunks in shared memory (bytes = 64 mod 128)
  int stages;         // samples in flight
  size_t smem_bytes;
};

inline Layout layout_for(int fields, int embed) {
  Layout l;
  const int payload = fields * embed + 1;
  l.pitch_floats = (payload + 31) / 32 * 32;
  l.copy_bytes = (payload * 4 + 15) / 16 * 16;
  int sp = (l.copy_bytes + 127) / 128 * 128 + 64;
  if (sp - 128 >= l.copy_bytes) sp -= 128;
  l.stage_pitch = sp / 4;
  const int pairs = fields * (fields - 1) / 2;
  const size_t fixed = (size_t)((pairs + 1) / 2 * 2) * sizeof(uint16_t) * 2 + 8 + 2 * kMaxStages * 8 +
                       kMaxStages * kConsumers * 4 + 64;
  const size_t stage = (size_t)fields * sp;
  int st = static_cast<int>(((size_t)kMaxDynSmem - fixed) / stage);
  l.stages = st > kMaxStages ? kMaxStages : st;
  l.smem_bytes = (size_t)(l.stages > 0 ? l.stages : 2) * stage + fixed;
  return l;
}

__global__ void __launch_bounds__(256) pack_kernel(const float* const* __restrict__ tables,
                                                   const float* __restrict__ w_feat, int64_t rows, int fields,
                                                   int embed, int pitch_floats, float* __restrict__ packed) {
  // work item = one float of the packed table; a warp writes 128 contiguous bytes and reads (mostly) one table row
  const int64_t items = rows * pitch_floats;
  const int payload = fields * embed;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = item / pitch_floats;
    const int c = static_cast<int>(item - r * pitch_floats);
    float v = 0.f;
    if (c < payload) {
      const int t = c / embed;
      v = __ldg(tables[t] + r * embed + (c - t * embed));
    } else if (c == payload && w_feat != nullptr) {
      v = __ldg(w_feat + r);
    }
    packed[item] = v;
  }
}

template <int IdxBits>
__global__ void __launch_bounds__(kThreads, 1) ffm_interleaved_kernel(const void* __restrict__ idx,
                                                                      const int64_t* __restrict__ offsets,
                                                                      int64_t batch, int fields, int embed,
                                                                      const float* __restrict__ packed, int64_t rows,
                                                                      int pitch_floats, int copy_bytes, int stage_pitch,
                                                                      int stages, const float* __restrict__ bias,
                                                                      float* __restrict__ logits, int32_t* status) {
  extern __shared__ __align__(128) unsigned char fi_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = fields * (fields - 1) / 2;
  float* stage0 = reinterpret_cast<float*>(fi_smem);
  const size_t stage_floats = (size_t)fields * stage_pitch;
  unsigned char* tail = fi_smem + (size_t)stages * stage_floats * sizeof(float);
  uint16_t* pair_i = reinterpret_cast<uint16_t*>(tail);
  uint16_t* pair_j = pair_i + (pairs + 1) / 2 * 2;
  unsigned char* ctl = reinterpret_cast<unsigned char*>(pair_j + (pairs + 1) / 2 * 2);
  ctl = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ctl) + 7) & ~uintptr_t(7));
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctl);            // full[kMaxStages], empty[kMaxStages]
  float* part = reinterpret_cast<float*>(bars + 2 * kMaxStages);   // [kMaxStages][kConsumers]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);

  for (int p = threadIdx.x; p < pairs; p += kThreads) {
    int i, j;
    pair_from_index(p, fields, i, j);
    pair_i[p] = static_cast<uint16_t>(i);
    pair_j[p] = static_cast<uint16_t>(j);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full0 + 8 * s, kProducers);
      mbar_init(empty0 + 8 * s, kConsumers);
    }
    tc5::fence_barrier_init();
  }
  __syncthreads();

  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t mine = first < batch ? (batch - first + step - 1) / step : 0;   // samples of this CTA

  if (warp >= kConsumers) {
    // ------------------------------------------------------------------ producers: warp j takes the fields j, j + P, ..
    const int pw = warp - kConsumers;
    const int f = pw + lane * kProducers;                    // this lane's field (lanes beyond the fields idle)
    const int my_fields = pw < fields ? (fields - pw + kProducers - 1) / kProducers : 0;
    const float bias_v = bias != nullptr ? __ldg(bias) : 0.f;
    auto resolve = [&](int64_t k) -> int64_t {
      if (k >= mine || f >= fields) return 0;
      const int64_t pos = (first + k * step) * fields + f;
      int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + f);
      if (r < 0 || r >= rows) {
        report_oob(status, pos);
        r = 0;
      }
      return r;
    };
    auto finish = [&](int64_t k, int s, uint32_t wrap) {   // sample k's partial sums are complete (stage s)
      mbar_wait(empty0 + 8 * s, wrap & 1);
      if (pw == 0 && lane == 0) {                          // one logit, fixed summation order
        float v = bias_v;
#pragma unroll
        for (int w = 0; w < kConsumers; ++w) v += part[s * kConsumers + w];
        logits[first + k * step] = v;
      }
      __syncwarp();
    };
    int64_t r = resolve(0), r_next = resolve(1);   // row ids two samples ahead: a narrow shard's sample lasts less
    int s = 0;                                     // than an index load
    uint32_t wrap = 0;
    for (int64_t k = 0; k < mine; ++k) {
      if (wrap > 0) finish(k - stages, s, wrap - 1);
      const uint32_t bar = full0 + 8 * s;
      if (lane == 0) mbar_expect_tx(bar, static_cast<uint32_t>(my_fields) * copy_bytes);
      __syncwarp();
      if (f < fields)
        bulk_g2s(smem_u32(stage0 + s * stage_floats) + f * stage_pitch * 4, packed + r * pitch_floats, copy_bytes, bar);
      r = r_next;
      r_next = resolve(k + 2);
      if (++s == stages) { s = 0; ++wrap; }
    }
    for (int64_t k = mine > stages ? mine - stages : 0; k < mine; ++k)
      finish(k, static_cast<int>(k % stages), static_cast<uint32_t>(k / stages));
  } else {
    // ------------------------------------------------------------------ consumers
    const int lpp = embed >> 2;                 // lanes per pair: one 16-byte piece each (power of two)
    const int lpp_shift = 31 - __clz(lpp);
    const int items = pairs << lpp_shift;
    const int first_at = fields * embed;
    int s = 0;
    uint32_t wrap = 0;
    for (int64_t k = 0; k < mine; ++k) {
      mbar_wait(full0 + 8 * s, wrap & 1);
      const float* A = stage0 + s * stage_floats;
      float acc = 0.f;
      for (int item = warp * 32 + lane; item < items; item += kConsumers * 32) {
        const int p = item >> lpp_shift, c = (item & (lpp - 1)) << 2;
        const int i = pair_i[p], j = pair_j[p];
        const float4 a = *reinterpret_cast<const float4*>(A + j * stage_pitch + i * embed + c);
        const float4 b = *reinterpret_cast<const float4*>(A + i * stage_pitch + j * embed + c);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
      if (warp == 0)
        for (int f = lane; f < fields; f += 32) acc += A[f * stage_pitch + first_at];
      acc = warp_sum(acc);
      if (lane == 0) part[s * kConsumers + warp] = acc;
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);
      if (++s == stages) { s = 0; ++wrap; }
    }
  }
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int64_t trs_ffm_interleaved_pitch(int fields, int embed) {
  if (fields <= 0 || embed <= 0) return 0;
  return layout_for(fields, embed).pitch_floats;
}

extern "C" int trs_ffm_pack_tables(const float* const* tables, const float* w_feat, int64_t rows, int fields, int embed,
                                   float* packed, void* stream) {
  TRS_REQUIRE(tables && packed, "trs_ffm_pack_tables: null pointer");
  TRS_REQUIRE(rows >= 0 && fields > 0 && embed > 0, "trs_ffm_pack_tables: bad sizes");
  if (rows == 0) return TRS_OK;
  const Layout l = layout_for(fields, embed);
  pack_kernel<<<grid_for(rows * l.pitch_floats, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tables, w_feat, rows, fields, embed, l.pitch_floats, packed);
  return check_launch("ffm pack_kernel");
}

extern "C" int trs_ffm_model_forward_interleaved(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                                 int fields, const float* packed, int64_t rows, int embed,
                                                 const float* bias, float* logits, int32_t* status, void* stream) {
  TRS_REQUIRE(idx && offsets && packed && logits && status, "trs_ffm_model_forward_interleaved: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_ffm_model_forward_interleaved: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 1 && rows > 0 && embed > 0, "trs_ffm_model_forward_interleaved: bad sizes");
  TRS_UNSUPPORTED(fields > 64, "trs_ffm_model_forward_interleaved: at most 64 fields (got %d)", fields);
  TRS_UNSUPPORTED(embed < 4 || embed > 128 || (embed & (embed - 1)) != 0,
                  "trs_ffm_model_forward_interleaved: embed must be a power of two in [4, 128] (got %d)", embed);
  TRS_UNSUPPORTED(!aligned16(packed), "trs_ffm_model_forward_interleaved: packed table must be 16-byte aligned");
  const Layout l = layout_for(fields, embed);
  TRS_UNSUPPORTED(l.stages < 2,
                  "trs_ffm_model_forward_interleaved: two samples of %d x %d do not fit shared memory", fields, embed);
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = static_cast<int>(batch < kNumSMs ? batch : kNumSMs);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN(ffm_interleaved_kernel<64>);
    ffm_interleaved_kernel<64><<<grid, kThreads, l.smem_bytes, s>>>(idx, offsets, batch, fields, embed, packed, rows,
                                                                   l.pitch_floats, l.copy_bytes, l.stage_pitch, l.stages,
                                                                   bias, logits, status);
  } else {
    TRS_SMEM_OPT_IN(ffm_interleaved_kernel<32>);
    ffm_interleaved_kernel<32><<<grid, kThreads, l.smem_bytes, s>>>(idx, offsets, batch, fields, embed, packed, rows,
                                                                   l.pitch_floats, l.copy_bytes, l.stage_pitch, l.stages,
                                                                   bias, logits, status);
  }
  return check_launch("ffm_interleaved_kernel");
}
