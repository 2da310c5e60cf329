// DeepFM forward (indices -> logits) on the packed [v|w] table with layer 1 on the 5th-generation tensor cores.
//
// Why a second kernel for the headline path: the round-1 kernel (deepfm_packed.cu) spends ~450 warp instructions per
// sample -- mma.sync layer 1 with the TF32 hi/lo split of every activation done by the warps -- and is latency bound at
// one CTA of 8 warps per SM (65.9 us even when every row is an L2 hit, profiles/r01_deepfm_notes.md).  The memory
// system itself saturates at 35.7 G random 128-byte lines/s whatever the request shape or depth
// (tools/r2_probe.cu, profiles/r02_gather_ceiling.md): 71.6 us for the 2.56 M rows of a 65 536-sample batch.  To
// sit on that ceiling the SM has to do almost nothing per row, so here
//   * layer 1 (the 624 x 16 GEMM per sample) is tcgen05.mma kind::tf32 with a 128-sample tile on M, the accumulators in
//     tensor memory; the gathered rows are staged by cp.async STRAIGHT INTO the UMMA K-major operand layout
//     (16-byte chunk planes), so the "hi" operand is the raw fp32 row (the tensor core truncates fp32 to TF32 --
//     measured, tools/r2_probe.cu `trunc`), W1 is pre-split once per model into [hi | lo] rows (N = 32), and only the
//     residual lo = v - trunc(v) is produced by threads, written to tensor memory (tcgen05.st) and multiplied from
//     there (A operand in TMEM):  h1 = (trunc(v) + lo) (W_hi + W_lo), fp32 accumulate = the 3xTF32 scheme plus the
//     lo x lo term (a width-32 MMA costs what a width-16 one does);
//   * warp roles (18 warps in the one-CTA-per-SM shape): 4 row producers (cp.async, 8 lanes per row = ONE 80-byte
//     request [v|w]; each keeps its OWN ring of raw indices, requested eight stages ahead with 8-/4-byte cp.async --
//     the loaded DRAM latency is ~5 us, anything shallower starves), 8 consumers (one pass over the staged rows: FM
//     sums in fp32, first-order sum, lo residual), 4 epilogue warps (tcgen05.ld -> bias + ReLU -> the 16x16 hidden
//     layers in FP32 FFMA -> logit), 1 MMA issuer, 1 loader of the W1 slabs (cp.async.bulk); everything is ordered
//     by mbarriers, no CTA barrier in the loop;
//   * a stage = 128 samples x G fields; the second shape (two CTAs of 12 warps per SM, G = 2) lets consecutive
//     launches overlap on every SM under programmatic dependent launch: no dead time between batches.
// Per sample ~150 warp instructions instead of ~450, none of them on the critical path of the gather.
#include <stdlib.h>

#include "tc5.cuh"

namespace trs {
namespace {

using namespace tc5;

constexpr int kTileM = 128;                 // samples per tile = UMMA M
constexpr int kPlane = 2048 + 64;           // bytes between 16-byte chunk planes (+64: cp.async bank spreading)
constexpr int kPlanesPerField = 5;          // 4 chunks of v + the chunk that holds w
constexpr int kSlabRows = 32;               // W1 rows per chunk plane: 16 hi + 16 lo
constexpr int kRidPitch = kTileM + 1;       // row-id table pitch (ints): conflict-free scattered stores
constexpr int kMaxHidden = 4;
constexpr int kAccCols = 64;                // per tile: [0,32) = trunc(v) x [W_hi | W_lo], [32,64) = lo x [W_hi | W_lo]
constexpr int kLoCol0 = 128;                // lo operand ring in tensor memory: STAGES x (G x 16) columns

// Shape of the pipeline.  G fields per stage (a stage = 128 samples x G fields), STAGES stages (STAGES - 1 in flight),
// PW producer warps, CH = 2: eight consumer warps (each thread half a row) / 1: four (a whole row), CTAS = CTAs per SM
// the launch bounds are written for, IDXD = how many stages ahead the raw indices are requested (ring slots).
template <int G_, int STAGES_, int PW_, int CH_, int CTAS_, int IDXD_>
struct Cfg {
  static constexpr int G = G_, STAGES = STAGES_, PW = PW_, CH = CH_, CTAS = CTAS_, IDXD = IDXD_;
  static constexpr int SPW = kTileM / PW;                       // samples per producer warp
  static constexpr int kIdxPerWarp = SPW * G;                   // indices per stage and producer warp
  static constexpr int kStageBytes = G * kPlanesPerField * kPlane;
  static constexpr int kSlabBytes = G * 4 * kSlabRows * 16;
  static constexpr int kConsumerWarps = 4 * CH, kEpilogueWarps = 4;
  static constexpr int kWarpConsumer0 = PW, kWarpEpilogue0 = PW + kConsumerWarps;
  static constexpr int kWarpMma = kWarpEpilogue0 + kEpilogueWarps, kWarpSlab = kWarpMma + 1;
  static constexpr int kThreads = (kWarpSlab + 1) * 32;
  static constexpr int kTmemCols = kLoCol0 + STAGES * G * 16 <= 256 ? 256 : 512;
  static_assert(SPW == 32 || SPW == 64, "a producer warp owns 32 or 64 samples of the tile");
  static_assert(kIdxPerWarp % 32 == 0 && IDXD >= STAGES && IDXD <= 8, "bad index ring");
};
using CfgBig = Cfg<3, 4, 4, 2, 1, 8>;     // one CTA per SM: 18 warps, 4 stages of 384 rows
using CfgDuo = Cfg<2, 3, 2, 1, 2, 8>;     // two CTAs per SM (consecutive launches overlap on every SM): 12 warps

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// 8- / 4-byte copies of raw indices (src_bytes = 0: nothing is read, the destination is zero-filled)
__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// predicated form: no branch around the copy (pred == 0: the instruction is not issued at all)
// Row copies allocate in L1 (.ca): with skewed (Zipf / tiny-field) indices the same few rows are requested by most
// samples of a stage and by every SM at the same moment -- through L2 only (.cg) those lines hot-spot a handful of L2
// slices (measured: Criteo-shaped batches 92.7 -> 67.5 us with .ca, uniform batches unchanged).
__device__ __forceinline__ void cp_async16_zfill_if(uint32_t dst, const void* src, int src_bytes, int pred) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %3, 0;\n\t"
      "@p cp.async.ca.shared.global [%0], [%1], 16, %2;\n\t"
      "}" ::"r"(dst), "l"(src), "r"(src_bytes), "r"(pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// the mbarrier receives one arrival from this thread when all of its earlier cp.async copies have landed (no thread
// has to come back for them); .noinc: the arrival is part of the barrier's initial count
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct Tc5Args {
  const void* idx;
  const int64_t* offsets;
  const float* packed;          // (rows, 32): v[16], w, pad
  // row-sharded table (world > 1): row g lives on rank g % world at local row g / world of shard[g % world] --
  // local HBM for this rank's own shard, peer-mapped memory (NVLink) for the others
  const float* shard[8];
  int world;
  const float* w1p;             // prepared W1: [group][12 chunk planes][32 rows: hi 0..15, lo 16..31][4]
  const float* b1;
  const float* wh[kMaxHidden];  // (16, 16)
  const float* bh[kMaxHidden];
  const float* w_out;           // (1, 16)
  const float* b_out;
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, groups, hidden_layers;
  int samples_per_cta;          // multiple of 4
  long long* trace;             // debug: per-role event clocks of CTA 0 (trs_debug_tc5_trace), else null
};
constexpr int kTraceSlots = 512, kTraceEvents = 4;
__device__ __forceinline__ void trace_ev(const Tc5Args& a, int role, int q, int k) {
  if (a.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && q < kTraceSlots)
    a.trace[(role * kTraceSlots + q) * kTraceEvents + k] = clock64();
}

struct Layout {   // shared-memory carve-up (bytes from the base)
  int stages, slabs, ring, hid, bias, side, off, bars, tmem_slot, total;
};
template <class C>
__host__ __device__ inline Layout make_layout(int fields, int idx_bytes) {
  Layout l;
  int p = 0;
  l.stages = p; p += C::STAGES * C::kStageBytes;
  l.slabs = p;  p += C::STAGES * C::kSlabBytes;
  l.ring = p;   p += C::PW * C::IDXD * C::kIdxPerWarp * idx_bytes;
  p = (p + 15) & ~15;
  l.hid = p;    p += kMaxHidden * 256 * 4;
  l.bias = p;   p += ((1 + kMaxHidden) * 16 + 16 + 4) * 4;
  l.side = p;   p += 2 * 2 * kTileM * 4;
  l.off = p;    p += ((fields * 8 + 15) & ~15) + 64;   // + the 8 shard base addresses
  l.bars = p;   p += (4 * C::STAGES + 8) * 8;
  l.tmem_slot = p; p += 16;
  l.total = p;
  return l;
}

template <int IdxBits, class C, bool kSharded>
__global__ void __launch_bounds__(C::kThreads, C::CTAS) deepfm_tc5_kernel(Tc5Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int kIdxBytes = IdxBits / 8;
  constexpr int G = C::G, S = C::STAGES;
  const Layout L = make_layout<C>(a.fields, kIdxBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_fields = a.fields, n_groups = a.groups;

  float* hid_s = reinterpret_cast<float*>(smem + L.hid);
  float* bias_s = reinterpret_cast<float*>(smem + L.bias);
  float* side_s = reinterpret_cast<float*>(smem + L.side);
  long long* off_s = reinterpret_cast<long long*>(smem + L.off);
  unsigned long long* shard_s = reinterpret_cast<unsigned long long*>(smem + L.off + ((a.fields * 8 + 15) & ~15));
  const uint32_t bar0 = smem_u32(smem + L.bars);
  auto v_full = [&](int s) { return bar0 + 8u * s; };
  auto slot_free = [&](int s) { return bar0 + 8u * (S + s); };
  auto w_full = [&](int s) { return bar0 + 8u * (2 * S + s); };
  auto lo_full = [&](int s) { return bar0 + 8u * (3 * S + s); };
  const uint32_t bar1 = bar0 + 8u * 4 * S;
  auto acc_full = [&](int b) { return bar1 + 8u * b; };
  auto acc_empty = [&](int b) { return bar1 + 8u * (2 + b); };
  auto side_full = [&](int b) { return bar1 + 8u * (4 + b); };
  auto side_empty = [&](int b) { return bar1 + 8u * (6 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);

  pdl_launch_dependents();
  if (threadIdx.x == 0) trace_ev(a, 6, 0, 0);

  // this CTA's samples [cta_begin, cta_begin + n_cta): tiles of 128, the last one partial
  const int64_t cta_begin = (int64_t)blockIdx.x * a.samples_per_cta;
  const int n_cta = static_cast<int>(a.batch - cta_begin < a.samples_per_cta ? a.batch - cta_begin : a.samples_per_cta);
  const int n_tiles = (n_cta + kTileM - 1) / kTileM;
  const int n_stages_total = n_tiles * n_groups;

  // ---- one-time set-up ------------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(v_full(s), C::PW * 32);   // one asynchronous arrival per producer thread (cp.async completion)
      mbar_init(slot_free(s), 1);
      mbar_init(w_full(s), 1);
      mbar_init(lo_full(s), C::kConsumerWarps);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), C::kEpilogueWarps);
      mbar_init(side_full(b), C::kConsumerWarps);
      mbar_init(side_empty(b), C::kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp >= C::PW) {   // (the producers start requesting indices at once; they read the offsets after the barrier)
    for (int i = threadIdx.x - C::PW * 32; i < n_fields; i += blockDim.x - C::PW * 32) off_s[i] = __ldg(a.offsets + i);
    if (kSharded && threadIdx.x - C::PW * 32 < 8)
      shard_s[threadIdx.x - C::PW * 32] = reinterpret_cast<unsigned long long>(a.shard[threadIdx.x - C::PW * 32]);
    for (int i = threadIdx.x - C::PW * 32; i < a.hidden_layers * 256; i += blockDim.x - C::PW * 32)
      hid_s[i] = __ldg(a.wh[i >> 8] + (i & 255));
    if (warp == C::kWarpEpilogue0 && lane < 16) {
      bias_s[lane] = __ldg(a.b1 + lane);
      for (int l = 0; l < a.hidden_layers; ++l) bias_s[(1 + l) * 16 + lane] = __ldg(a.bh[l] + lane);
      bias_s[(1 + kMaxHidden) * 16 + lane] = __ldg(a.w_out + lane);
      if (lane == 0) bias_s[(1 + kMaxHidden) * 16 + 16] = __ldg(a.b_out);
    }
  }
  // producers: the raw indices of the first IDXD stages are requested before anything else
  constexpr int SPW = C::SPW, E = C::kIdxPerWarp, K = E / 32, D = C::IDXD;
  unsigned char* ring = smem + L.ring + (size_t)(warp < C::PW ? warp : 0) * D * E * kIdxBytes;
  const uint32_t ring_s = smem_u32(ring);
  int tf = 0, gf = 0;   // (tile, group) of the next index request
  auto fetch_idx = [&](int q) {
    const int valid_s = n_cta - tf * kTileM - SPW * warp;   // samples of this warp that exist in tile tf (may be <= 0)
    const unsigned char* src0 = static_cast<const unsigned char*>(a.idx) +
                                ((size_t)(cta_begin + (int64_t)tf * kTileM + SPW * warp) * n_fields + gf * G) * kIdxBytes;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int e = lane + 32 * k, sm = e / G, fl = e - sm * G;        // fetch order: sample-major (coalesced)
      const bool live = q < n_stages_total && sm < valid_s && gf * G + fl < n_fields;
      const unsigned char* src = live ? src0 + ((size_t)sm * n_fields + fl) * kIdxBytes : static_cast<const unsigned char*>(a.idx);
      const uint32_t dst = ring_s + ((q % D) * E + e) * kIdxBytes;
      if (IdxBits == 64) cp_async8_zfill(dst, src, live ? 8 : 0);
      else cp_async4_zfill(dst, src, live ? 4 : 0);
    }
    if (++gf == n_groups) { gf = 0; ++tf; }
  };
  if (warp < C::PW) {
    for (int q = 0; q < D; ++q) {
      fetch_idx(q);
      cp_async_commit();
    }
  }
  if (warp == C::kWarpMma) tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) trace_ev(a, 6, 0, 1);

  if (warp < C::PW) {
    // =========================== row producers =============================================================================
    // Warp pw owns samples [SPW pw, SPW (pw + 1)) of every tile.  Per stage: (1) its raw indices (requested IDXD stages
    // ago into the warp's own ring, coalesced: a sample's G fields are contiguous) -> row ids: offset add + range check,
    // lane <-> (field, sample); (2) the row copies, one instruction = 4 rows x 8 lanes (5 active: chunks 0..3 = v,
    // chunk 4 = the 16 bytes that hold w): ONE 80-byte request per row, the row id comes from its owner lane by shuffle;
    // (3) the request for the indices of stage q + IDXD.
    const int sub = lane & 7, rsel = lane >> 3;
    const bool lane_on = sub < 5;
    const unsigned long long src_lane = reinterpret_cast<unsigned long long>(a.packed) + 16ull * sub;
    const uint32_t stage0 = smem_u32(smem + L.stages) + sub * kPlane + (SPW * warp + rsel) * 16;
    bool waited = false;
    int q = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int tile_valid = n_cta - t * kTileM < kTileM ? n_cta - t * kTileM : kTileM;
      const int valid_s = tile_valid - SPW * warp;
      for (int g = 0; g < n_groups; ++g, ++q) {
        cp_async_wait<D - 1>();   // the indices of stage q (group q of this thread) have landed ...
        __syncwarp();             // ... and so have the other lanes'
        // resolve: element (fl, s) = (e2 / SPW, e2 % SPW), e2 = lane + 32 k -- the order the copy loop wants
        int rid[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int e2 = lane + 32 * k, fl = e2 / SPW, sm = e2 - fl * SPW, f = g * G + fl;
          rid[k] = -1;
          if (sm < valid_s && f < n_fields) {
            const unsigned char* slot_p = ring + (size_t)((q % D) * E + sm * G + fl) * kIdxBytes;
            const int64_t ix = IdxBits == 64 ? *reinterpret_cast<const long long*>(slot_p)
                                             : static_cast<int64_t>(*reinterpret_cast<const int*>(slot_p));
            const int64_t row = ix + off_s[f];
            if (row >= 0 && row < a.rows) {
              if (kSharded) {   // (owner rank, row within its shard) in one word: owner above bit 27
                const int g = static_cast<int>(row);
                rid[k] = ((g % a.world) << 28) | (g / a.world);
              } else {
                rid[k] = static_cast<int>(row);
              }
            } else {
              if (!waited) { pdl_wait(); waited = true; }   // status may still be written by the work this launch overlaps
              report_oob(a.status, (cta_begin + (int64_t)t * kTileM + SPW * warp + sm) * n_fields + f);
            }
          }
        }
        const int slot = q % S;
        mbar_wait(slot_free(slot), ((q / S) & 1) ^ 1);
        if (warp == 0) trace_ev(a, 0, q, 0);
        const uint32_t dst0 = stage0 + slot * C::kStageBytes;
#pragma unroll
        for (int fl = 0; fl < G; ++fl) {
#pragma unroll
          for (int i = 0; i < SPW / 4; ++i) {
            const int r = __shfl_sync(0xffffffffu, rid[(fl * SPW + 4 * i) / 32], (4 * i + rsel) & 31);
            unsigned long long src;
            if (kSharded) {
              const unsigned long long base = shard_s[max(r, 0) >> 28] + 16ull * sub;
              asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(static_cast<unsigned>(max(r, 0)) & 0x0fffffffu), "l"(base));
            } else {
              asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(static_cast<unsigned>(max(r, 0))), "l"(src_lane));
            }
            // row groups beyond a partial tile and the three idle lanes of every row issue nothing
            cp_async16_zfill_if(dst0 + fl * kPlanesPerField * kPlane + 4 * i * 16, reinterpret_cast<const void*>(src),
                                r >= 0 ? 16 : 0, lane_on && 4 * i < valid_s);
          }
        }
        // the stage is "full" when every producer thread's copies have landed: the hardware arrives for the thread,
        // which goes straight on to the next stage (S - 1 stages of rows in flight)
        cp_async_arrive_noinc(v_full(slot));
        if (warp == 0) trace_ev(a, 0, q, 1);
        fetch_idx(q + D);
        cp_async_commit();
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < C::kWarpEpilogue0) {
    // =========================== consumers: FM sums, first-order, lo residual -> tensor memory ================================
    constexpr int HALVES = 3 - C::CH;                  // column halves per thread: CH = 2 -> 1, CH = 1 -> 2
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const int h0 = C::CH == 2 ? (warp - C::kWarpConsumer0) >> 2 : 0;
    const int r = 32 * quad + lane;
    const unsigned char* row0 = smem + L.stages + r * 16;
    const uint32_t t_lo = tmem_base + (static_cast<uint32_t>(32 * quad) << 16) + kLoCol0;
    float s[8 * HALVES], sq[8 * HALVES], wacc = 0.f;   // per embedding component: sum and sum of squares over the fields
#pragma unroll
    for (int j = 0; j < 8 * HALVES; ++j) s[j] = sq[j] = 0.f;
    int q = 0;
    for (int t = 0; t < n_tiles; ++t) {
      for (int g = 0; g < n_groups; ++g, ++q) {
        const int slot = q % S;
        mbar_wait(v_full(slot), (q / S) & 1);
        if (warp == C::kWarpConsumer0) trace_ev(a, 1, q, 0);
        const unsigned char* base = row0 + slot * C::kStageBytes;
#pragma unroll
        for (int fl = 0; fl < G; ++fl) {
#pragma unroll
          for (int hh = 0; hh < HALVES; ++hh) {
            const int h = h0 + hh;
            const unsigned char* pl = base + (fl * kPlanesPerField + 2 * h) * kPlane;
            const float4 va = *reinterpret_cast<const float4*>(pl);
            const float4 vb = *reinterpret_cast<const float4*>(pl + kPlane);
            const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
            uint32_t lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s[8 * hh + j] += v[j];
              sq[8 * hh + j] = fmaf(v[j], v[j], sq[8 * hh + j]);
              const float hi = __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);   // what the tensor core reads
              lo[j] = (__float_as_uint(v[j] - hi) + 0x1000u) & 0xffffe000u;            // exact residual -> TF32 (rn)
            }
            tmem_st8(t_lo + slot * (G * 16) + fl * 16 + 8 * h, lo);
            if (h == 1)   // the first-order weight sits in the fifth chunk plane of the field
              wacc += *reinterpret_cast<const float*>(base + (fl * kPlanesPerField + 4) * kPlane);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(lo_full(slot));
        if (warp == C::kWarpConsumer0) trace_ev(a, 1, q, 1);
      }
      // tile done: this thread's share of  sum_e 0.5 (s_e^2 - q_e) + sum_n w
      // per component (sum x)^2 - sum x^2 like the reference (two rounded terms: exactly zero for a single field)
      float fm = 0.f;
#pragma unroll
      for (int j = 0; j < 8 * HALVES; ++j) fm += __fsub_rn(__fmul_rn(s[j], s[j]), sq[j]);
      const float part = fmaf(0.5f, fm, wacc);
      mbar_wait(side_empty(t & 1), ((t >> 1) & 1) ^ 1);
      side_s[(t & 1) * 2 * kTileM + h0 * kTileM + r] = part;
      if (C::CH == 1) side_s[(t & 1) * 2 * kTileM + kTileM + r] = 0.f;
      __syncwarp();
      if (lane == 0) mbar_arrive(side_full(t & 1));
#pragma unroll
      for (int j = 0; j < 8 * HALVES; ++j) s[j] = sq[j] = 0.f;
      wacc = 0.f;
    }
  } else if (warp < C::kWarpMma) {
    // =========================== epilogue: accumulators -> hidden layers -> logit =============================================
    const int quad = warp & 3;
    const int r = 32 * quad + lane;
    bool first_store = true;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int tile_valid = n_cta - t * kTileM < kTileM ? n_cta - t * kTileM : kTileM;
      mbar_wait(acc_full(buf), (t >> 1) & 1);
      if (warp == C::kWarpEpilogue0) trace_ev(a, 5, t, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * quad) << 16) + buf * kAccCols;
      float hcur[16];
      {   // h1 = ReLU(b1 + ((lo W_lo + lo W_hi) + trunc(v) W_lo) + trunc(v) W_hi): the small terms first
        uint32_t d[16];
        tmem_ld16(taddr + 48, d);
#pragma unroll
        for (int o = 0; o < 16; ++o) hcur[o] = __uint_as_float(d[o]);
        tmem_ld16(taddr + 32, d);
#pragma unroll
        for (int o = 0; o < 16; ++o) hcur[o] += __uint_as_float(d[o]);
        tmem_ld16(taddr + 16, d);
#pragma unroll
        for (int o = 0; o < 16; ++o) hcur[o] += __uint_as_float(d[o]);
        tmem_ld16(taddr, d);
#pragma unroll
        for (int o = 0; o < 16; ++o) hcur[o] = fmaxf(bias_s[o] + (hcur[o] + __uint_as_float(d[o])), 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));     // tensor memory of this tile is free again
      for (int layer = 0; layer < a.hidden_layers; ++layer) {
        float nxt[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const float4* wr = reinterpret_cast<const float4*>(hid_s + layer * 256 + o * 16);   // broadcast reads
          float acc = bias_s[(1 + layer) * 16 + o];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 w4 = wr[k4];
            acc = fmaf(w4.x, hcur[4 * k4], acc);
            acc = fmaf(w4.y, hcur[4 * k4 + 1], acc);
            acc = fmaf(w4.z, hcur[4 * k4 + 2], acc);
            acc = fmaf(w4.w, hcur[4 * k4 + 3], acc);
          }
          nxt[o] = fmaxf(acc, 0.f);
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) hcur[o] = nxt[o];
      }
      float out = bias_s[(1 + kMaxHidden) * 16 + 16];
#pragma unroll
      for (int o = 0; o < 16; ++o) out = fmaf(bias_s[(1 + kMaxHidden) * 16 + o], hcur[o], out);
      mbar_wait(side_full(buf), (t >> 1) & 1);
      out += side_s[buf * 2 * kTileM + r] + side_s[buf * 2 * kTileM + kTileM + r];
      __syncwarp();
      if (lane == 0) mbar_arrive(side_empty(buf));
      if (first_store) {
        pdl_wait();   // first global write: an overlapped previous grid must be complete
        first_store = false;
      }
      if (r < tile_valid) a.logits[cta_begin + (int64_t)t * kTileM + r] = out;
      if (warp == C::kWarpEpilogue0) trace_ev(a, 5, t, 1);
    }
  } else if (warp == C::kWarpMma) {
    // =========================== MMA issuer ================================================================================
    const uint32_t idesc = umma_idesc_tf32(32);
    const uint64_t a_desc0 = umma_desc(smem_u32(smem + L.stages), kPlane, 128);
    const uint64_t b_desc0 = umma_desc(smem_u32(smem + L.slabs), kSlabRows * 16, 128);
    // iteration q: the lo part of stage q - 1 (the consumers have written the residual to tensor memory) and the commit
    // that frees that stage, THEN the hi part of stage q (rows and W1 slab landed).  lo first: a slot is released as
    // early as possible and never waits for younger rows -- the tensor pipe has slack, the row slots do not.
    int t_hi = 0, g_hi = 0, t_lo = 0, g_lo = 0;
    for (int q = 0; q <= n_stages_total; ++q) {
      if (q > 0) {
        const int pq = q - 1, slot = pq % S;
        mbar_wait(lo_full(slot), (pq / S) & 1);
        trace_ev(a, 2, pq, 0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (t_lo & 1) * kAccCols + 32;
          const uint64_t bd = b_desc0 + ((slot * C::kSlabBytes) >> 4);
#pragma unroll
          for (int fl = 0; fl < G; ++fl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_tf32_ts(d, tmem_base + kLoCol0 + slot * (G * 16) + fl * 16 + 8 * ks,
                           bd + (((fl * 4 + 2 * ks) * kSlabRows * 16) >> 4), idesc,
                           (g_lo > 0 || fl > 0 || ks > 0) ? 1u : 0u);
          umma_commit(slot_free(slot));                      // rows, slab and lo columns of this stage are reusable
          if (g_lo == n_groups - 1) umma_commit(acc_full(t_lo & 1));
        }
        __syncwarp();
        if (++g_lo == n_groups) { g_lo = 0; ++t_lo; }
      }
      if (q < n_stages_total) {
        const int slot = q % S;
        if (g_hi == 0) mbar_wait(acc_empty(t_hi & 1), ((t_hi >> 1) & 1) ^ 1);   // epilogue of tile t - 2 has drained it
        mbar_wait(v_full(slot), (q / S) & 1);
        mbar_wait(w_full(slot), (q / S) & 1);
        trace_ev(a, 2, q, 1);
        fence_proxy_async();   // the rows were written by cp.async (generic proxy); the tensor core reads via the async proxy
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + (t_hi & 1) * kAccCols;
          const uint64_t ad = a_desc0 + ((slot * C::kStageBytes) >> 4), bd = b_desc0 + ((slot * C::kSlabBytes) >> 4);
#pragma unroll
          for (int fl = 0; fl < G; ++fl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              umma_tf32(d, ad + (((fl * kPlanesPerField + 2 * ks) * kPlane) >> 4),
                        bd + (((fl * 4 + 2 * ks) * kSlabRows * 16) >> 4), idesc,
                        (g_hi > 0 || fl > 0 || ks > 0) ? 1u : 0u);
        }
        __syncwarp();
        trace_ev(a, 2, q, 2);
        if (++g_hi == n_groups) { g_hi = 0; ++t_hi; }
      }
    }
  } else if (warp == C::kWarpSlab) {
    // =========================== W1 slab loader (L2 -> smem, one bulk copy per stage) ===========================================
    int g = 0;
    for (int q = 0; q < n_stages_total; ++q) {
      const int slot = q % S;
      mbar_wait(slot_free(slot), ((q / S) & 1) ^ 1);
      trace_ev(a, 3, q, 0);
      if (lane == 0) {
        mbar_expect_tx(w_full(slot), C::kSlabBytes);
        bulk_g2s(smem_u32(smem + L.slabs + slot * C::kSlabBytes),
                 reinterpret_cast<const unsigned char*>(a.w1p) + (size_t)g * C::kSlabBytes, C::kSlabBytes, w_full(slot));
      }
      __syncwarp();
      if (++g == n_groups) g = 0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_ev(a, 6, 0, 2);
  if (warp == C::kWarpMma) tmem_dealloc(tmem_base, C::kTmemCols);
}

// W1 (16, 16 N) -> w1p[g][pl = fl*4 + c][n][j]:  n < 16: hi(W1[n][(G g + fl)*16 + 4c + j]),  n >= 16: lo of output n - 16
__global__ void __launch_bounds__(256) deepfm_tc5_prep_kernel(const float* __restrict__ w1, int fields, int groups,
                                                              int fields_per_group, float* __restrict__ w1p) {
  const int planes = fields_per_group * 4;
  const int items = groups * planes * kSlabRows * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int j = i & 3, n = (i >> 2) & 31, pl = (i >> 7) % planes, g = (i >> 7) / planes;
    const int f = g * fields_per_group + pl / 4, c = pl & 3, o = n & 15;
    float v = 0.f;
    if (f < fields) v = __ldg(w1 + (size_t)o * 16 * fields + f * 16 + 4 * c + j);
    const uint32_t hi = tf32_rna(v);
    const uint32_t lo = tf32_rna(v - __uint_as_float(hi));
    w1p[i] = __uint_as_float(n < 16 ? hi : lo);
  }
}

inline int groups_of(int fields, int g) { return (fields + g - 1) / g; }
long long* g_trace = nullptr;

// variant 0 = CfgBig (one CTA per SM), 1 = CfgDuo (two CTAs per SM)
int variant_g(int variant) { return variant == 1 ? CfgDuo::G : CfgBig::G; }

template <class C>
int launch_tc5(Tc5Args a, int idx_bits, unsigned flags, cudaStream_t s) {
  a.groups = groups_of(a.fields, C::G);
  a.trace = g_trace;
  // Equal contiguous sample ranges, one CTA per SM slot.  The two-CTAs-per-SM shape launched with
  // TRS_LAUNCH_OVERLAP_PREVIOUS takes ONE slot per SM: the other one belongs to the neighbouring launch of the train
  // (batch k + 1 gathers while batch k drains -- no dead time between batches); launched alone it fills both slots.
  const int slots = (C::CTAS == 2 && !(flags & TRS_LAUNCH_OVERLAP_PREVIOUS)) ? 2 * kNumSMs : kNumSMs;
  int64_t want = (a.batch + 31) / 32;
  const int grid = static_cast<int>(want < slots ? want : slots);
  a.samples_per_cta = static_cast<int>((((a.batch + grid - 1) / grid) + 3) / 4 * 4);
  const int grid_used = static_cast<int>((a.batch + a.samples_per_cta - 1) / a.samples_per_cta);
  const size_t smem = make_layout<C>(a.fields, idx_bits / 8).total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_used);
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (flags & TRS_LAUNCH_OVERLAP_PREVIOUS) ? 1 : 0;
  cudaError_t e;
#define TRS_TC5_LAUNCH(BITS, SH)                                         \
  do {                                                                   \
    TRS_SMEM_OPT_IN((deepfm_tc5_kernel<BITS, C, SH>));                   \
    e = cudaLaunchKernelEx(&cfg, deepfm_tc5_kernel<BITS, C, SH>, a);     \
  } while (0)
  if (a.world > 1) {
    if (idx_bits == 64) TRS_TC5_LAUNCH(64, true);
    else TRS_TC5_LAUNCH(32, true);
  } else {
    if (idx_bits == 64) TRS_TC5_LAUNCH(64, false);
    else TRS_TC5_LAUNCH(32, false);
  }
#undef TRS_TC5_LAUNCH
  if (e != cudaSuccess) {
    set_error("launch of deepfm_tc5_kernel failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return TRS_ERR_CUDA;
  }
  return TRS_OK;
}

}  // namespace

int deepfm_tc5_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation, int64_t rows,
                         int variant) {
  static const bool disabled = getenv("TRS_DISABLE_DEEPFM_TC5") != nullptr;
  if (disabled || variant < 0 || variant > 1) return 0;
  if (embed != 16 || activation != TRS_ACT_RELU || rows >= (int64_t(1) << 31) || fields < 1) return 0;
  if (mlp_layers < 2 || mlp_layers - 2 > kMaxHidden) return 0;
  if (mlp_dims[0] != fields * 16 || mlp_dims[mlp_layers] != 1) return 0;
  for (int l = 1; l < mlp_layers; ++l)
    if (mlp_dims[l] != 16) return 0;
  if (variant == 1)   // two CTAs per SM: half of the SM's shared memory each (1 KB per CTA is reserved by the system)
    return make_layout<CfgDuo>(fields, 8).total <= (kMaxDynSmem + 1024) / 2 - 1024;
  return make_layout<CfgBig>(fields, 8).total <= kMaxDynSmem;
}

}  // namespace trs

using namespace trs;

// debug hook: device buffer of 7 x 512 x 4 int64 that receives the event clocks of CTA 0's roles (null: off)
extern "C" int trs_debug_tc5_trace(long long* device_buf) {
  g_trace = device_buf;
  return TRS_OK;
}

extern "C" int64_t trs_deepfm_tc_workspace_bytes(int fields, int variant) {
  if (fields < 1 || variant < 0 || variant > 1) return 0;
  return variant == 1 ? (int64_t)groups_of(fields, CfgDuo::G) * CfgDuo::kSlabBytes
                      : (int64_t)groups_of(fields, CfgBig::G) * CfgBig::kSlabBytes;
}

extern "C" int trs_deepfm_tc_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation,
                                       int64_t rows, int variant) {
  if (!mlp_dims || mlp_layers < 1) return 0;
  return deepfm_tc5_supported(fields, embed, mlp_dims, mlp_layers, activation, rows, variant);
}

extern "C" int trs_deepfm_tc_prepare(int fields, const float* w1, int variant, float* workspace, void* stream) {
  TRS_REQUIRE(w1 && workspace && fields >= 1, "trs_deepfm_tc_prepare: bad arguments");
  TRS_REQUIRE(variant == 0 || variant == 1, "trs_deepfm_tc_prepare: variant must be 0 or 1");
  TRS_REQUIRE(aligned16(workspace), "trs_deepfm_tc_prepare: workspace must be 16-byte aligned");
  const int g = variant_g(variant);
  const int groups = groups_of(fields, g);
  const int items = groups * g * 4 * kSlabRows * 4;
  deepfm_tc5_prep_kernel<<<(items + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w1, fields, groups, g,
                                                                                           workspace);
  return check_launch("deepfm_tc5_prep_kernel");
}

static int forward_tc_common(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                             const float* packed, const float* const* shards, int world, int64_t rows,
                             const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                             int activation, const float* workspace, int variant, float* logits, int32_t* status,
                             unsigned flags, void* stream);

extern "C" int trs_deepfm_forward_tc_sharded(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                             int fields, const float* const* shards, int world, int64_t rows,
                                             const int* mlp_dims, int mlp_layers, const float* const* mlp_w,
                                             const float* const* mlp_b, int activation, const float* workspace,
                                             int variant, float* logits, int32_t* status, unsigned flags,
                                             void* stream) {
  TRS_REQUIRE(shards && world >= 1 && world <= 8, "trs_deepfm_forward_tc_sharded: 1..8 shards expected");
  for (int r = 0; r < world; ++r)
    TRS_REQUIRE(shards[r] && (reinterpret_cast<uintptr_t>(shards[r]) & 127u) == 0,
                "trs_deepfm_forward_tc_sharded: shard %d is null or not 128-byte aligned", r);
  TRS_UNSUPPORTED(world > 1 && (rows + world - 1) / world >= (int64_t(1) << 28),
                  "trs_deepfm_forward_tc_sharded: at most 2^28 rows per shard");
  return forward_tc_common(idx, idx_bits, offsets, batch, fields, shards[0], shards, world, rows, mlp_dims, mlp_layers,
                           mlp_w, mlp_b, activation, workspace, variant, logits, status, flags, stream);
}

extern "C" int trs_deepfm_forward_tc(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                     const float* packed, int64_t rows, const int* mlp_dims, int mlp_layers,
                                     const float* const* mlp_w, const float* const* mlp_b, int activation,
                                     const float* workspace, int variant, float* logits, int32_t* status,
                                     unsigned flags, void* stream) {
  return forward_tc_common(idx, idx_bits, offsets, batch, fields, packed, nullptr, 1, rows, mlp_dims, mlp_layers, mlp_w,
                           mlp_b, activation, workspace, variant, logits, status, flags, stream);
}

static int forward_tc_common(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                             const float* packed, const float* const* shards, int world, int64_t rows,
                             const int* mlp_dims, int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                             int activation, const float* workspace, int variant, float* logits, int32_t* status,
                             unsigned flags, void* stream) {
  TRS_REQUIRE((flags & ~TRS_LAUNCH_OVERLAP_PREVIOUS) == 0, "trs_deepfm_forward_tc: unknown flags 0x%x", flags);
  TRS_REQUIRE(idx && offsets && packed && logits && mlp_dims && mlp_w && mlp_b && workspace,
              "trs_deepfm_forward_tc: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_deepfm_forward_tc: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0 && mlp_layers >= 1, "trs_deepfm_forward_tc: bad sizes");
  TRS_UNSUPPORTED(!deepfm_tc5_supported(fields, 16, mlp_dims, mlp_layers, activation, rows, variant),
                  "trs_deepfm_forward_tc: needs embed 16, hidden widths 16, ReLU, rows < 2^31, variant 0/1 and a field "
                  "count whose staging fits shared memory");
  TRS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127u) == 0 && aligned16(idx) && aligned16(workspace),
              "trs_deepfm_forward_tc: packed table must be 128-byte aligned, idx and workspace 16-byte aligned");
  if (batch == 0) return TRS_OK;
  Tc5Args a{};
  a.idx = idx; a.offsets = offsets; a.packed = packed; a.w1p = workspace; a.logits = logits; a.status = status;
  a.batch = batch; a.rows = rows; a.fields = fields;
  a.world = world;
  for (int r = 0; r < 8; ++r) a.shard[r] = (shards != nullptr && r < world) ? shards[r] : packed;
  a.hidden_layers = mlp_layers - 2;
  for (int l = 0; l < mlp_layers; ++l) TRS_REQUIRE(mlp_w[l] && mlp_b[l], "trs_deepfm_forward_tc: null MLP parameter");
  a.b1 = mlp_b[0];
  for (int l = 0; l < a.hidden_layers; ++l) {
    a.wh[l] = mlp_w[1 + l];
    a.bh[l] = mlp_b[1 + l];
  }
  a.w_out = mlp_w[mlp_layers - 1];
  a.b_out = mlp_b[mlp_layers - 1];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return variant == 1 ? launch_tc5<CfgDuo>(a, idx_bits, flags, s) : launch_tc5<CfgBig>(a, idx_bits, flags, s);
}
