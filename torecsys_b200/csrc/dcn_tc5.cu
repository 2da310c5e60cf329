// DeepAndCrossNetworkModel forward, indices -> logits, with the per-row dense chains on the 5th-generation tensor cores
// and the activations of a tile living in TENSOR MEMORY from the gathered row to the logit
// (torecsys/models/ctr/deep_and_cross_network.py:76-98: fc(flatten(cat[Cross(x), MLP_per_field(x)])) over
//  torecsys/layers/ctr/cross_network.py:65-79 and torecsys/layers/ctr/multilayer_perceptron.py:53-84).
//
// Work per (sample, field) row x (E floats): cross chain h <- x * (W_l h + b_l) + x (L layers), per-field MLP on x, and the
// row's share of the final Linear: <fc_w[f, :E], cross_out> + <fc_w[f, E:], mlp_out>.  BASELINE configs[2]: E = 32, L = 6,
// MLP 32-32-16-8-4, 39 fields, batch 131 072 = 5.1 M rows per launch.
//
// Mapping (what the round-2 probes showed, profiles/r02_gather_ceiling.md + tools/r2_probe.cu):
//   * tcgen05.mma kind::tf32 with the A operand in tensor memory costs N/2 cycles (N = 16: 9, N = 32: 16): narrow chains
//     are NOT penalised per instruction -- as long as the MMAs are issued under elect.sync (tc5.cuh elect_one) and A does
//     not come from shared memory (SS form: ~40 cycles whatever N).  So: rows on M (128 per tile = floor(128 / fields)
//     whole samples), every layer = 3 * K/8 MMAs (3xTF32: A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate) of N in
//     {16, 32}, A from TMEM, W pre-split in shared memory;
//   * the bias rides in the GEMM: every layer has one more k-step whose A columns are the constant [1, 0, .., 0] (written
//     once per kernel into the slot's TMEM) and whose W rows hold [b_hi | b_lo]: two extra N/2-cycle MMAs replace an LDS +
//     FADD per element in the epilogue, which is the issue-bound side;
//   * a layer's epilogue (thread = row): tcgen05.ld D -> FP32 update -> raw value as the hi operand (the tensor core
//     truncates to TF32 itself), lo = v - trunc(v) -> tcgen05.st as the next layer's A.  No shared-memory round trip;
//   * kSlots tiles are in flight per CTA, each with its OWN MMA-issuing warp blocked on that slot's operand barrier, so
//     the slots drift apart and the tensor pipe, the TMEM path and the FP32 pipe overlap (one issuer serving the slots
//     round-robin with blocking waits locks them in phase: 33 % tensor-pipe activity in cross_tc5.cu; one issuer
//     polling them burns a quarter of the SM's issue slots: measured here, 1.06 ms);
//   * rows are gathered by each epilogue warp for ITS 32 rows of the slot's next tile: indices two tiles ahead in a
//     register, the 128-byte rows one tile ahead by cp.async into a staging buffer (8 lanes per row: one coalesced
//     request per row); x0 then lives in registers for the whole chain;
//   * a tile holds whole samples, so the per-sample sum of the row partials is a fixed-order sum inside the CTA
//     (shared memory + a named barrier per slot): deterministic, no atomics.
// Shapes outside (E in {16, 32}, ReLU, MLP widths <= 32, fields <= 128) stay on the mma.sync kernel (dcn_tc.cu).
#include <stdlib.h>

#include "tc5.cuh"
#include "tile_ops.cuh"

namespace trs {
namespace {

using namespace tc5;

constexpr int kSlots = 4;
constexpr int kMaxSteps = 16;
constexpr int kSlotCols = 104;
constexpr int kThreads = kSlots * 128 + kSlots * 32;   // 4 epilogue warps + 1 MMA-issuing warp per slot

enum StepKind { kCross = 0, kCrossLast = 1, kDeepHidden = 2, kDeepOut = 3 };

struct Step {
  int w_off;      // float offset of the step's hi plane in the weight store ([hi|lo][(K + 8)/4][npad][4])
  int b_off;      // (unused: the bias is row K of the weights)
  int k, npad;    // K of the activations (multiple of 8; the weights have K + 8 rows) and N (16 or 32) of the MMAs
  int kind;
  int n_valid;    // outputs that exist (kDeepOut: the MLP's output width)
};

struct DcnTc5Args {
  const void* idx;
  const int64_t* offsets;
  const float* w_emb;
  const float* cross_w;   // (L, E, E)
  const float* cross_b;   // (L, E)
  const float* fc_w;      // (1, N * (E + Od))
  const float* fc_b;
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, cross_layers, od;
  int n_steps, w_floats, b_floats;
  Step steps[kMaxSteps];
  MlpParams mp;
};

__device__ __forceinline__ void named_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// v[0..16) -> 16 columns of this thread's TMEM lane: raw values as the hi plane (the MMA truncates them), exact
// remainders as lo.  The caller issues tcgen05.wait::st once per layer.
__device__ __forceinline__ void write_operand16(uint32_t t_hi, uint32_t t_lo, const float (&v)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = __float_as_uint(v[j]);
    lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j] & 0xffffe000u));
  }
  tmem_st16(t_hi, hi);
  tmem_st16(t_lo, lo);
}

template <int E, int IdxBits>
__global__ void __launch_bounds__(kThreads, 1) dcn_tc5_kernel(const DcnTc5Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kRowPitch = E + 4;   // floats: rows 16-byte aligned, a thread's LDS.128 of its own row conflict-free
  const int n_fields = a.fields, cat = E + a.od;
  float* w_s = reinterpret_cast<float*>(smem_raw);
  float* b_s = w_s + a.w_floats;
  float* fcw_s = b_s + a.b_floats;                               // [fields][E + od]
  float* part_s = fcw_s + ((n_fields * cat + 3) & ~3);           // [kSlots][128] row partials
  long long* off_s = reinterpret_cast<long long*>(part_s + kSlots * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(off_s + ((n_fields + 1) & ~1));   // a_ready[kSlots], d_full[kSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kSlots);
  float* stage_s = reinterpret_cast<float*>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));      // [kSlots * 4 warps][32][kRowPitch]
  const uint32_t bar0 = smem_u32(bars);
  auto a_ready = [&](int s) { return bar0 + 8u * s; };
  auto d_full = [&](int s) { return bar0 + 8u * (kSlots + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- weights -> pre-split, K-major core-matrix layout; biases; fc_w; offsets --------------------------------------
  for (int t = 0; t < a.n_steps; ++t) {
    const Step st = a.steps[t];
    const bool cross = st.kind == kCross || st.kind == kCrossLast;
    const int layer = cross ? t : t - a.cross_layers;
    const float* w = cross ? a.cross_w + (size_t)layer * E * E : a.mp.w[layer];
    const float* b = cross ? a.cross_b + (size_t)layer * E : a.mp.b[layer];
    const int k_in = cross ? E : a.mp.dims[layer], n_out = cross ? E : a.mp.dims[layer + 1];
    const int kk = st.k + 8;                 // + the bias k-step: row K = bias, rows K+1 .. K+7 = 0
    const int plane = kk * st.npad;
    for (int i = threadIdx.x; i < plane; i += blockDim.x) {
      const int n = i / kk, k = i - n * kk;
      float v = 0.f;
      if (n < n_out) {
        if (k < k_in) v = __ldg(w + (size_t)n * k_in + k);
        else if (k == st.k && b != nullptr) v = __ldg(b + n);
      }
      const uint32_t hi = tf32_rna(v);
      const uint32_t lo = tf32_rna(v - __uint_as_float(hi));
      const int pos = ((k >> 2) * st.npad + n) * 4 + (k & 3);
      w_s[st.w_off + pos] = __uint_as_float(hi);
      w_s[st.w_off + plane + pos] = __uint_as_float(lo);
    }
  }
  for (int i = threadIdx.x; i < n_fields * cat; i += blockDim.x) fcw_s[i] = __ldg(a.fc_w + i);
  for (int i = threadIdx.x; i < n_fields; i += blockDim.x) off_s[i] = __ldg(a.offsets + i);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(a_ready(s), 4);   // one arrive per epilogue warp of the slot
      mbar_init(d_full(s), 1);    // tcgen05.commit
    }
    fence_barrier_init();
  }
  if (warp == kSlots * 4) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async();   // the generic-proxy weight stores are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int spt = 128 / n_fields;            // whole samples per tile
  const int tile_rows = spt * n_fields;
  const int64_t tiles = (a.batch + spt - 1) / spt;
  const int64_t total_rows = a.batch * n_fields;
  const int64_t stride = (int64_t)gridDim.x * kSlots;

  if (warp < kSlots * 4) {
    // =========================== epilogue warps: slot = warp / 4, TMEM lane quarter = warp % 4 ====================
    const int slot = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;                                  // row of the tile = TMEM lane
    // kSlotCols columns per slot whatever E is: D [0, 32) | A hi [32, 64) | constant [64, 72) | A lo [72, 104)
    const uint32_t t_d = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + slot * kSlotCols;
    const uint32_t t_hi = t_d + 32, t_lo = t_d + 72;
    float* stage_w = stage_s + (size_t)warp * 32 * kRowPitch;
    const uint32_t stage_w_s = smem_u32(stage_w);
    constexpr int kChunks = E / 4;                                // 16-byte chunks per row
    // resolved table row of this thread's row in tile `t` (-1: no such row / out of range -> zeros)
    auto resolve = [&](int64_t t) -> int {
      if (t >= tiles || r >= tile_rows) return -1;
      const int64_t m = t * tile_rows + r;
      if (m >= total_rows) return -1;
      const int64_t row = load_index<IdxBits>(a.idx, m) + off_s[r % n_fields];
      if (row < 0 || row >= a.rows) {
        report_oob(a.status, m);
        return -1;
      }
      return static_cast<int>(row);
    };
    auto gather = [&](int rid) {   // the warp's 32 rows -> staging, 8 (E = 32) lanes per row
#pragma unroll
      for (int k = 0; k < kChunks; ++k) {
        const int c = k * 32 + lane;
        const int rr = c / kChunks, ch = c - rr * kChunks;
        const int id = __shfl_sync(0xffffffffu, rid, rr);
        const float* src = a.w_emb + (id >= 0 ? static_cast<int64_t>(id) * E + 4 * ch : 0);
        const uint32_t dst = stage_w_s + (rr * kRowPitch + 4 * ch) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(id >= 0 ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const float* fcw = fcw_s + (r < tile_rows ? r % n_fields : 0) * cat;   // a tile starts at a sample: field = r % fields
    {   // the constant k-step of every layer: A columns [64, 72) = [1, 0, .., 0]  (the 8 columns behind are A lo, rewritten per layer)
      uint32_t one[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) one[j] = j == 0 ? __float_as_uint(1.0f) : 0u;
      tmem_st16(t_d + 64, one);
      tmem_st_wait();
    }
    const int64_t tile0 = (int64_t)blockIdx.x * kSlots + slot;
    gather(resolve(tile0));
    int rid_next = resolve(tile0 + stride);
    uint32_t n_full = 0;
    for (int64_t tile = tile0; tile < tiles; tile += stride) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      float x0[E];
      {
        const float* mine = stage_w + lane * kRowPitch;
#pragma unroll
        for (int c = 0; c < E; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(mine + c);
          x0[c] = v.x; x0[c + 1] = v.y; x0[c + 2] = v.z; x0[c + 3] = v.w;
        }
      }
      __syncwarp();                       // every lane has its row: the buffer takes the next tile's rows
      gather(rid_next);
      rid_next = resolve(tile + 2 * stride);
      auto write_x0 = [&]() {
#pragma unroll
        for (int c = 0; c < E; c += 16) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = x0[c + j];
          write_operand16(t_hi + c, t_lo + c, v);
        }
      };
      write_x0();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(slot));
      float partial = 0.f;
      for (int t = 0; t < a.n_steps; ++t) {
        const Step st = a.steps[t];
        mbar_wait(d_full(slot), n_full & 1);
        ++n_full;
        tc_fence_after();
        if (st.kind == kCross || st.kind == kCrossLast) {
          // 16 columns at a time (register budget: x0 stays resident for the whole chain)
#pragma unroll
          for (int c = 0; c < E; c += 16) {
            uint32_t raw[16];
            tmem_ld16(t_d + c, raw);
            float h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = fmaf(x0[c + j], __uint_as_float(raw[j]), x0[c + j]);
            if (st.kind == kCross) {
              write_operand16(t_hi + c, t_lo + c, h);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) partial = fmaf(h[j], fcw[c + j], partial);
            }
          }
          if (st.kind == kCrossLast) write_x0();   // the per-field MLP starts from the row again
          tmem_st_wait();
        } else {
          // deep layers: 16 or 32 (padded) outputs; hidden -> ReLU -> next A operand, output -> its share of the logit
#pragma unroll
          for (int c = 0; c < 32; c += 16) {
            if (c < st.npad) {
              uint32_t raw[16];
              tmem_ld16(t_d + c, raw);
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
              if (st.kind == kDeepHidden) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                write_operand16(t_hi + c, t_lo + c, v);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c + j < st.n_valid) partial = fmaf(v[j], fcw[E + c + j], partial);
              }
            }
          }
          if (st.kind == kDeepHidden) tmem_st_wait();
        }
        if (t + 1 < a.n_steps) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_ready(slot));
        }
      }
      // ---- per-sample sum of the row partials, fixed order ---------------------------------------------------------
      part_s[slot * 128 + r] = partial;
      named_barrier(1 + slot, 128);
      if (r < spt) {
        const int64_t b = tile * spt + r;
        if (b < a.batch) {
          float sum = __ldg(a.fc_b);
          const float* p = part_s + slot * 128 + r * n_fields;
          for (int i = 0; i < n_fields; ++i) sum += p[i];
          a.logits[b] = sum;
        }
      }
      // (part_s of this slot is rewritten a whole tile later: every warp of the slot passes n_steps MMA round trips,
      //  each of which needs this warp's arrival, before it gets there)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // =========================== MMA issuers: one warp per slot, blocked on ITS slot's operand ========================
    // (one warp polling all slots burned a quarter of the SM's issue slots on mbarrier tests -- and starved the
    //  epilogue warps of its scheduler, which every slot waits for; a round-robin of blocking waits locks the slots in
    //  phase.  Per-slot issuers sleep in mbarrier.try_wait and the slots drift apart freely.)
    const int s = warp - kSlots * 4;
    const uint32_t w_base = smem_u32(w_s);
    const uint32_t d = tmem_base + s * kSlotCols;
    uint32_t n_ready = 0;
    for (int64_t tile = (int64_t)blockIdx.x * kSlots + s; tile < tiles; tile += stride) {
      for (int t = 0; t < a.n_steps; ++t) {
        // everything the MMAs need is computed BEFORE the wait: after the wake-up only the issue remains
        const Step st = a.steps[t];
        const uint32_t idesc = umma_idesc_tf32(st.npad);
        const uint32_t lbo = st.npad * 16;                                   // bytes between 16-byte K chunks
        const uint64_t b_hi0 = umma_desc(w_base + st.w_off * 4, lbo, 128);
        const uint64_t b_lo0 = b_hi0 + (((st.k + 8) * st.npad * 4) >> 4);
        const uint32_t step_u = (2 * lbo) >> 4;                              // one k-step = two chunks
        mbar_wait(a_ready(s), n_ready & 1);
        ++n_ready;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks * 8 < st.k) {
#pragma unroll
              for (int term = 0; term < 3; ++term) {   // 0: A_lo*B_hi, 1: A_hi*B_lo, 2: A_hi*B_hi
                const uint32_t a_op = d + (term == 0 ? 72 : 32) + 8 * ks;
                umma_tf32_ts(d, a_op, (term == 1 ? b_lo0 : b_hi0) + ks * step_u, idesc, (ks > 0 || term > 0) ? 1u : 0u);
              }
            }
          }
          const uint32_t kb = (st.k >> 3) * step_u;    // the bias k-step: A = [1, 0, ..] (its lo plane is zero)
          umma_tf32_ts(d, d + 64, b_lo0 + kb, idesc, 1u);
          umma_tf32_ts(d, d + 64, b_hi0 + kb, idesc, 1u);
          umma_commit(d_full(s));
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSlots * 4) tmem_dealloc(tmem_base, 512);
}

struct Plan {
  DcnTc5Args a;
  size_t smem;
};

bool make_plan(int embed, int fields, int cross_layers, const MlpParams& mp, Plan& p) {
  DcnTc5Args& a = p.a;
  if (cross_layers < 1 || cross_layers + mp.layers > kMaxSteps || mp.layers < 1) return false;
  int w = 0, b = 0, n = 0;
  for (int l = 0; l < cross_layers; ++l) {
    a.steps[n++] = Step{w, b, embed, embed, l + 1 < cross_layers ? kCross : kCrossLast, embed};
    w += 2 * (embed + 8) * embed;
    b += embed;
  }
  int k = embed;   // K of the next deep layer = padded width of the previous one
  for (int l = 0; l < mp.layers; ++l) {
    const int out = mp.dims[l + 1], npad = out <= 16 ? 16 : 32;
    if (out < 1 || out > 32 || mp.dims[l] > k) return false;
    a.steps[n++] = Step{w, b, k, npad, l + 1 < mp.layers ? kDeepHidden : kDeepOut, out};
    w += 2 * (k + 8) * npad;
    b += npad;
    k = npad;
  }
  a.n_steps = n;
  a.w_floats = w;
  a.b_floats = (b + 3) & ~3;
  const int cat = embed + mp.dims[mp.layers];
  size_t bytes = ((size_t)a.w_floats + a.b_floats + ((fields * cat + 3) & ~3) + kSlots * 128) * sizeof(float);
  bytes += ((fields + 1) & ~1) * sizeof(long long) + 2 * kSlots * 8 + 16 + 128;
  bytes += (size_t)kSlots * 4 * 32 * (embed + 4) * sizeof(float);
  p.smem = bytes;
  return bytes <= (size_t)kMaxDynSmem;
}

template <int E>
int dispatch(Plan& p, int idx_bits, cudaStream_t s) {
  const int spt = 128 / p.a.fields;
  const int64_t tiles = (p.a.batch + spt - 1) / spt;
  const int64_t want = (tiles + kSlots - 1) / kSlots;
  const int grid = static_cast<int>(want < kNumSMs ? want : kNumSMs);
  if (idx_bits == 64) {
    TRS_SMEM_OPT_IN((dcn_tc5_kernel<E, 64>));
    dcn_tc5_kernel<E, 64><<<grid, kThreads, p.smem, s>>>(p.a);
  } else {
    TRS_SMEM_OPT_IN((dcn_tc5_kernel<E, 32>));
    dcn_tc5_kernel<E, 32><<<grid, kThreads, p.smem, s>>>(p.a);
  }
  return check_launch("dcn_tc5_kernel");
}

}  // namespace

int dcn_tc5_supported(int embed, int fields, int cross_layers, const int* mlp_dims, int mlp_layers, int activation,
                      int64_t rows) {
  static const bool disabled = getenv("TRS_DISABLE_TC5") != nullptr || getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled) return 0;
  if (!(embed == 16 || embed == 32) || fields < 1 || fields > 128 || activation != TRS_ACT_RELU) return 0;
  if (rows >= (int64_t(1) << 31) || mlp_layers < 1 || mlp_layers > MlpParams::kMaxLayers || mlp_dims[0] != embed) return 0;
  MlpParams mp{};
  mp.layers = mlp_layers;
  for (int l = 0; l <= mlp_layers; ++l) mp.dims[l] = mlp_dims[l];
  Plan p{};
  return make_plan(embed, fields, cross_layers, mp, p) ? 1 : 0;
}

int dcn_tc5_launch(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                   const float* w_emb, int64_t rows, int embed, const float* cross_w, const float* cross_b,
                   int cross_layers, const MlpParams& mp, const float* fc_w, const float* fc_b, float* logits,
                   int32_t* status, cudaStream_t s) {
  Plan p{};
  if (!make_plan(embed, fields, cross_layers, mp, p)) {
    set_error("trs_dcn_forward: shape not covered by the tcgen05 kernel");
    return TRS_ERR_UNSUPPORTED;
  }
  DcnTc5Args& a = p.a;
  a.idx = idx; a.offsets = offsets; a.w_emb = w_emb; a.cross_w = cross_w; a.cross_b = cross_b; a.fc_w = fc_w;
  a.fc_b = fc_b; a.logits = logits; a.status = status; a.batch = batch; a.rows = rows; a.fields = fields;
  a.cross_layers = cross_layers; a.od = mp.dims[mp.layers]; a.mp = mp;
  TRS_REQUIRE(aligned16(w_emb), "trs_dcn_forward: w_emb must be 16-byte aligned");
  return embed == 32 ? dispatch<32>(p, idx_bits, s) : dispatch<16>(p, idx_bits, s);
}

}  // namespace trs
