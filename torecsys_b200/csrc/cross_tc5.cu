// a7 CrossNetworkLayer on the 5th-generation tensor cores: the whole L-layer chain of a 128-row tile stays in TENSOR
// MEMORY -- h_{l+1} = x0 * (h_l W_l^T + b_l) + x0 per (sample, field) row (torecsys/layers/ctr/cross_network.py:52-87).
//
//   * a CTA keeps kSlots independent 128-row tiles in flight.  Per tile ("slot") TMEM holds the accumulator D (E
//     columns) and the next layer's A operand as TF32 hi / lo planes (E columns each): lane = row, column = k.
//   * layer l of a slot = 3 * E/8 tcgen05.mma (kind::tf32, M = 128, N = E, K = 8; 3xTF32 split: A_lo*B_hi + A_hi*B_lo
//     + A_hi*B_hi, fp32 accumulate) with A read from TMEM and B = W_l (pre-split, K-major) from shared memory, issued
//     by one thread of the MMA warp, then tcgen05.commit -> mbarrier d_full[slot];
//   * the slot's four epilogue warps (thread = row, x0 of the row in registers) read D with tcgen05.ld, apply
//     x0 * (D + b_l) + x0 in FP32, split the result and write it straight back as the next A operand with
//     tcgen05.st (no shared-memory round trip, no proxy fence), then arrive on a_ready[slot]; the last layer's rows go
//     to global memory as full 128-byte lines;
//   * while one slot is in its epilogue the MMA warp serves the others, so the tensor pipe, the TMEM load/store path
//     and the FP32 pipe overlap.  The input is read once and the output written once: 2 * E * 4 bytes per row.
// Shapes: E in {16, 32}, up to kMaxLayers layers.
//
// MEASURED VERDICT (B200, 1.28 M rows, E = 32, 6 layers; tools/tc5_slots.sh, profiles/r01c_cross_tc5_notes.md): correct
// to 1e-5, but 325 us against 242 us for the register-resident mma.sync chain of dcn_tc.cu.  A tcgen05.mma of
// N = 32, K = 8 costs ~85-100 cycles whatever its size (the same as an N = 128 one), the 12 MMAs of a layer cannot be
// merged (K = 8 per instruction for TF32), and the time does not change with 1..4 tiles in flight per CTA or with
// 1..3 CTAs per SM: the tensor pipe's per-instruction floor bounds the chain, not the FP32 pipe (issue slots 26 %
// used) and not memory.  The kernel is therefore NOT on the default path of trs_cross_forward; it is kept, tested,
// behind trs_cross_forward_tc5 as the record of that experiment.  Narrow dense chains belong on mma.sync.
#include <stdlib.h>

#include "tc5.cuh"

namespace trs {
namespace {

using namespace tc5;

constexpr int kMaxSlots = 5;   // 5 x 3 x 32 = 480 of the 512 TMEM columns
constexpr int kMaxLayers = 12;

struct CrossTc5Args {
  const float* x;
  const float* w;   // (L, E, E)
  const float* b;   // (L, E)
  float* out;
  int64_t rows;
  int layers;
};

__host__ __device__ inline size_t stage_offset(int layers, int e, int slots) {
  const size_t fixed = ((size_t)layers * 2 * e * e + (size_t)layers * e) * sizeof(float) + 2 * slots * 8 + 16;
  return (fixed + 127) / 128 * 128;
}

template <int E>
__device__ __forceinline__ void write_operand(uint32_t taddr_hi, const float (&h)[E]) {
  // h -> TF32 hi plane [E columns] | lo plane [E columns] of this thread's TMEM lane
#pragma unroll
  for (int c = 0; c < E; c += 16) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // tcgen05.mma kind::tf32 TRUNCATES its fp32 operands (profiles/r02_gather_ceiling.md): the raw value is a valid
      // hi operand, and lo = h - trunc(h) is exact (its own truncation costs 2^-21 relative, one-sided)
      hi[j] = __float_as_uint(h[c + j]);
      lo[j] = __float_as_uint(h[c + j] - __uint_as_float(hi[j] & 0xffffe000u));
    }
    tmem_st16(taddr_hi + c, hi);
    tmem_st16(taddr_hi + E + c, lo);
  }
  tmem_st_wait();
}

template <int E, int kSlots>
__global__ void __launch_bounds__(kSlots * 128 + 32, 1) cross_tc5_kernel(CrossTc5Args a) {
  constexpr int kGroups = kSlots;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // weights: [layer][hi|lo][kc = E/4][n = E][4 floats]  (K-major core matrices: 8 rows x 16 bytes contiguous)
  constexpr int kPlaneFloats = E * E;
  float* w_s = reinterpret_cast<float*>(smem_raw);
  float* b_s = w_s + (size_t)a.layers * 2 * kPlaneFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + a.layers * E);   // a_ready[kSlots], d_full[kSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots);
  constexpr int kRowPitch = E + 4;                      // floats; rows 16-byte aligned, LDS.128 conflict-free
  constexpr int kStageFloats = 32 * kRowPitch;          // one buffer of one warp
  float* stage_s = reinterpret_cast<float*>(smem_raw + stage_offset(a.layers, E, kSlots));
  const uint32_t bar0 = smem_u32(bars);
  auto a_ready = [&](int s) { return bar0 + 8u * s; };
  auto d_full = [&](int s) { return bar0 + 8u * (kSlots + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < a.layers * kPlaneFloats; i += blockDim.x) {
    const int l = i / kPlaneFloats, rem = i - l * kPlaneFloats;
    const int n = rem / E, k = rem - n * E;          // W_l[n][k]
    const float v = __ldg(a.w + i);
    const uint32_t hi = tf32_rna(v);
    const uint32_t lo = tf32_rna(v - __uint_as_float(hi));
    const int pos = ((k >> 2) * E + n) * 4 + (k & 3);
    w_s[(size_t)(2 * l) * kPlaneFloats + pos] = __uint_as_float(hi);
    w_s[(size_t)(2 * l + 1) * kPlaneFloats + pos] = __uint_as_float(lo);
  }
  for (int i = threadIdx.x; i < a.layers * E; i += blockDim.x) b_s[i] = __ldg(a.b + i);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(a_ready(s), 4);   // one arrive per epilogue warp of the slot
      mbar_init(d_full(s), 1);    // tcgen05.commit
    }
    fence_barrier_init();
  }
  // TMEM columns of this CTA: a power of two >= kSlots * 3E, so that several CTAs can share the SM's 512 columns
  constexpr uint32_t kTmemCols = kSlots * 3 * E <= 32 ? 32 : kSlots * 3 * E <= 64 ? 64 : kSlots * 3 * E <= 128 ? 128
                                 : kSlots * 3 * E <= 256 ? 256 : 512;
  if (warp == kSlots * 4) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
  fence_proxy_async();   // the generic-proxy weight stores above are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t tiles = (a.rows + 127) / 128;
  const int64_t stride = (int64_t)gridDim.x * kSlots;

  if (warp < kSlots * 4) {
    // =========================== epilogue warps: slot = warp / 4, TMEM lane quarter = warp % 4 ====================
    // Rows travel through a per-warp staging buffer so that every global access is a full coalesced 512-byte warp
    // request: the 32 rows of a warp are 32 * E * 4 contiguous bytes; the next tile's rows are prefetched with cp.async
    // while the current tile runs its layers, and the last layer's rows leave through the same buffer.
    const int slot = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + slot * (3 * E);
    float* stage_w = stage_s + (size_t)warp * 2 * kStageFloats;          // [2][32 rows][kRowPitch]
    const uint32_t stage_w_s = smem_u32(stage_w);
    constexpr int kChunksPerRow = E / 4;                                  // 16-byte chunks per row
    constexpr int kCopies = 32 * kChunksPerRow / 32;                      // cp.async per lane per tile
    auto prefetch = [&](int64_t tile, int buf) {
      const int64_t row0 = tile * 128 + (warp & 3) * 32;
#pragma unroll
      for (int k = 0; k < kCopies; ++k) {
        const int c = k * 32 + lane;
        const int rr = c / kChunksPerRow, ch = c - rr * kChunksPerRow;
        const bool live = tile < tiles && row0 + rr < a.rows;
        const float* src = a.x + (live ? (row0 + rr) * E + 4 * ch : 0);
        const uint32_t dst = stage_w_s + ((buf * 32 + rr) * kRowPitch + 4 * ch) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(live ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t n_full = 0;
    int cur = 0;
    prefetch((int64_t)blockIdx.x * kSlots + slot, 0);
    for (int64_t tile = (int64_t)blockIdx.x * kSlots + slot; tile < tiles; tile += stride, cur ^= 1) {
      const int64_t m = tile * 128 + r;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      float x0[E];
      const float* mine = stage_w + (cur * 32 + lane) * kRowPitch;
#pragma unroll
      for (int c = 0; c < E; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(mine + c);
        x0[c] = v.x; x0[c + 1] = v.y; x0[c + 2] = v.z; x0[c + 3] = v.w;
      }
      prefetch(tile + stride, cur ^ 1);
      write_operand<E>(t_lane + E, x0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(slot));
      for (int l = 0; l < a.layers; ++l) {
        mbar_wait(d_full(slot), n_full & 1);
        ++n_full;
        tc_fence_after();
        float h[E];
#pragma unroll
        for (int c = 0; c < E; c += 16) {
          uint32_t raw[16];
          tmem_ld16(t_lane + c, raw);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            h[c + j] = fmaf(x0[c + j], __uint_as_float(raw[j]) + b_s[l * E + c + j], x0[c + j]);
        }
        if (l + 1 < a.layers) {
          write_operand<E>(t_lane + E, h);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready(slot));
        } else {
          float* row = stage_w + (cur * 32 + lane) * kRowPitch;   // every lane has read its x0 out of this buffer
#pragma unroll
          for (int c = 0; c < E; c += 4)
            *reinterpret_cast<float4*>(row + c) = make_float4(h[c], h[c + 1], h[c + 2], h[c + 3]);
          __syncwarp();
          const int64_t row0 = tile * 128 + (warp & 3) * 32;
#pragma unroll
          for (int k = 0; k < kCopies; ++k) {
            const int c = k * 32 + lane;
            const int rr = c / kChunksPerRow, ch = c - rr * kChunksPerRow;
            if (row0 + rr < a.rows)
              stg_stream_f4(reinterpret_cast<float4*>(a.out + (row0 + rr) * E + 4 * ch),
                            *reinterpret_cast<const float4*>(stage_w + (cur * 32 + rr) * kRowPitch + 4 * ch));
          }
          __syncwarp();   // the buffer is the prefetch target of the next iteration
        }
      }
      (void)m;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // =========================== MMA issuer ===============================================================================
    // The slots form kGroups groups that are served alternately, one layer per visit: while the MMAs of one group
    // run, the other groups are in their epilogues.  Inside a group the MMAs of the different slots are issued
    // interleaved: consecutive MMAs into the SAME accumulator serialise on it, different accumulators pipeline.
    constexpr int kPer = kSlots / kGroups;
    const uint32_t idesc = umma_idesc_tf32(E);
    constexpr uint32_t kLbo = E * 16;                                  // bytes between 16-byte K chunks
    const uint64_t w_desc0 = umma_desc(smem_u32(w_s), kLbo, 128);
    constexpr uint32_t kPlaneU = (kPlaneFloats * 4) >> 4;              // 16-byte units per (layer, plane)
    constexpr uint32_t kStepU = (2 * kLbo) >> 4;                       // one k-step = 8 tf32 = two chunks
    int64_t tile[kSlots];
    uint32_t n_ready[kSlots];
    int layer[kGroups];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
      tile[s] = (int64_t)blockIdx.x * kSlots + s;
      n_ready[s] = 0;
    }
#pragma unroll
    for (int g = 0; g < kGroups; ++g) layer[g] = 0;
    for (;;) {
      bool any = false;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        bool act[kPer];
        bool group_any = false;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          act[j] = tile[g * kPer + j] < tiles;
          group_any = group_any || act[j];
        }
        if (!group_any) continue;
        any = true;
        const int l = layer[g];
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
          if (!act[j]) continue;
          const int s = g * kPer + j;
          mbar_wait(a_ready(s), n_ready[s] & 1);
          ++n_ready[s];
        }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t b_hi0 = w_desc0 + (uint64_t)(2 * l) * kPlaneU, b_lo0 = b_hi0 + kPlaneU;
#pragma unroll
          for (int ks = 0; ks < E / 8; ++ks) {
            const uint64_t b_hi = b_hi0 + ks * kStepU, b_lo = b_lo0 + ks * kStepU;
#pragma unroll
            for (int term = 0; term < 3; ++term) {   // 0: A_lo*B_hi, 1: A_hi*B_lo, 2: A_hi*B_hi
#pragma unroll
              for (int j = 0; j < kPer; ++j) {
                if (!act[j]) continue;
                const uint32_t d = tmem_base + (g * kPer + j) * (3 * E);
                const uint32_t a_op = d + (term == 0 ? 2 * E : E) + 8 * ks;
                umma_tf32_ts(d, a_op, term == 1 ? b_lo : b_hi, idesc, (ks > 0 || term > 0) ? 1u : 0u);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < kPer; ++j)
            if (act[j]) umma_commit(d_full(g * kPer + j));
        }
        __syncwarp();
        if (++layer[g] == a.layers) {
          layer[g] = 0;
#pragma unroll
          for (int j = 0; j < kPer; ++j) tile[g * kPer + j] += stride;
        }
      }
      if (!any) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSlots * 4) tmem_dealloc(tmem_base, kTmemCols);
}

template <int E, int kSlots>
int cross_tc5_dispatch(const CrossTc5Args& a, cudaStream_t s) {
  const size_t smem = stage_offset(a.layers, E, kSlots) + (size_t)kSlots * 4 * 2 * 32 * (E + 4) * sizeof(float);
  if (smem > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;   // many layers: the caller falls through to mma.sync
  const int64_t tiles = (a.rows + 127) / 128;
  const int64_t want = (tiles + kSlots - 1) / kSlots;
  const int cols = kSlots * 3 * E <= 32 ? 32 : kSlots * 3 * E <= 64 ? 64 : kSlots * 3 * E <= 128 ? 128
                   : kSlots * 3 * E <= 256 ? 256 : 512;
  int per_sm = 512 / cols;                                   // CTAs per SM by tensor memory ...
  while (per_sm > 1 && per_sm * (smem + 1024) > 227 * 1024) --per_sm;   // ... and by shared memory
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  const int grid = static_cast<int>(want < cap ? want : cap);
  TRS_SMEM_OPT_IN((cross_tc5_kernel<E, kSlots>));
  cross_tc5_kernel<E, kSlots><<<grid, kSlots * 128 + 32, smem, s>>>(a);
  return check_launch("cross_tc5_kernel");
}

template <int E>
int cross_tc5_pick(const CrossTc5Args& a, cudaStream_t s) {
  static const int slots = getenv("TRS_TC5_SLOTS") ? atoi(getenv("TRS_TC5_SLOTS")) : 3;   // 1: 148 us, 2: 151, 3: 141, 4: 161
  switch (slots) {
    case 1: return cross_tc5_dispatch<E, 1>(a, s);
    case 2: return cross_tc5_dispatch<E, 2>(a, s);
    case 3: return cross_tc5_dispatch<E, 3>(a, s);
    case 5: return cross_tc5_dispatch<E, 5>(a, s);
    default: return cross_tc5_dispatch<E, 4>(a, s);
  }
}

}  // namespace

// CrossNetworkLayer on tcgen05; TRS_ERR_UNSUPPORTED when the shape is not covered
int cross_tc5_launch(const float* x, const float* w, const float* b, int layers, int64_t rows, int embed, float* out,
                     cudaStream_t s) {
  if (layers < 1 || layers > kMaxLayers || rows < 128 * kMaxSlots || !aligned16(x) || !aligned16(out))
    return TRS_ERR_UNSUPPORTED;
  CrossTc5Args a{x, w, b, out, rows, layers};
  switch (embed) {
    case 16: return cross_tc5_pick<16>(a, s);
    case 32: return cross_tc5_pick<32>(a, s);
  }
  return TRS_ERR_UNSUPPORTED;
}

}  // namespace trs

using namespace trs;

extern "C" int trs_cross_forward_tc5(const float* x, const float* weights, const float* biases, int layers, int64_t rows,
                                     int embed, float* out, void* stream) {
  TRS_REQUIRE(x && out && weights && biases, "trs_cross_forward_tc5: null pointer");
  TRS_REQUIRE(rows >= 0 && embed > 0 && layers >= 1, "trs_cross_forward_tc5: bad sizes");
  if (rows == 0) return TRS_OK;
  const int rc = cross_tc5_launch(x, weights, biases, layers, rows, embed, out, static_cast<cudaStream_t>(stream));
  TRS_UNSUPPORTED(rc == TRS_ERR_UNSUPPORTED,
                  "trs_cross_forward_tc5: needs embed 16 or 32, >= 640 rows, 16-byte aligned x/out, weights of all "
                  "layers in shared memory");
  return rc;
}
