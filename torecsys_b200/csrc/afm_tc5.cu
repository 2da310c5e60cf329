// a11 on the 5th-generation tensor cores: AttentionalFactorizationMachineLayer forward
// (torecsys/layers/ctr/attentional_factorization_machine.py:99-120, eval mode) for embed in {16, 32}, attn <= 32.
//
//   prod_p = x_i * x_j (P = N(N-1)/2 pairs x E);  raw_p = w2 . relu(W1 prod_p + b1) + b2;  scores = softmax_p(raw);
//   out = sum_p scores_p prod_p
// The P x E x A contraction is 95 % of the work.  Round 1 ran it as register-resident mma.sync tiles (afm_tc.cu: 52 M
// samples/s at 39 x 16 x 16, issue-bound).  Here it is the same scheme as dcn_tc5.cu: the (sample, pair) rows of the
// whole batch are cut into 128-row tiles; thread = row builds prod_p from the two field rows, writes it to TENSOR MEMORY
// as the A operand (raw value = hi, the tensor core truncates; v - trunc(v) = lo); the layer is 3 * E/8 + 2
// tcgen05.mma kind::tf32 of N = 16 / 32 with A in TMEM (N/2 cycles each), W1 pre-split in shared memory, b1 riding as one
// more k-step against a constant [1, 0, ..] block; the epilogue reads the A accumulators, applies ReLU and the w2 dot and
// stores ONE raw score per row.  Every slot (4 epilogue warps + 1 MMA-issuing warp) keeps TWO tiles in flight: the
// operand of tile k + 1 is built before the wait for tile k's accumulator.
// Second kernel (one warp per sample): softmax over the sample's P raw scores (in place) and the weighted sum, which
// re-forms the products from x (cheaper than storing them: 47 KB per sample).
#include <stdlib.h>

#include "tc5.cuh"

namespace trs {
namespace {

using namespace tc5;

constexpr int kMaxSlots = 5;
constexpr int kMaxFields = 64;

struct AfmTc5Args {
  const float* x;       // (B, N, E)
  const float* w1;      // (A, E)
  const float* b1;      // (A)
  const float* w2;      // (A)
  const float* b2;      // (1)
  float* scores;        // (B, P): raw scores out
  int64_t rows;         // B * P
  int fields, pairs, attn, npad, slots, buf_cols;
  int smax;             // samples a 128-row tile can touch: 127 / pairs + 2
};

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

template <int E>
__global__ void __launch_bounds__(kMaxSlots * 160, 1) afm_tc5_kernel(const AfmTc5Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int npad = a.npad, kk = E + 8;
  float* w_s = reinterpret_cast<float*>(smem_raw);            // [hi|lo][kk/4][npad][4]
  float* w2_s = w_s + 2 * kk * npad;                          // [npad] (zero padded), then b2
  int* ptab = reinterpret_cast<int*>(w2_s + npad + 4);        // [pairs]: i << 16 | j
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(ptab + a.pairs) + 7) & ~uintptr_t(7));   // a_ready[slots][2], d_full[slots][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * kMaxSlots);
  constexpr int kPitch = E + 4;   // floats per staged field row: rows i, i+1, .. fall into different bank groups
  float* xs_all = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
  const int xbuf_floats = a.smax * a.fields * kPitch;            // one tile's samples; [slot][2] of them
  const uint32_t bar0 = smem_u32(bars);
  auto a_ready = [&](int s, int b) { return bar0 + 8u * (2 * s + b); };
  auto d_full = [&](int s, int b) { return bar0 + 8u * (2 * kMaxSlots + 2 * s + b); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slots = a.slots;
  const bool is_mma = warp >= slots * 4;

  // ---- W1 (+ b1 as row E) -> pre-split, K-major core-matrix layout; w2, b2; pair table -----------------------------
  for (int i = threadIdx.x; i < kk * npad; i += blockDim.x) {
    const int n = i / kk, k = i - n * kk;
    float v = 0.f;
    if (n < a.attn) {
      if (k < E) v = __ldg(a.w1 + n * E + k);
      else if (k == E) v = __ldg(a.b1 + n);
    }
    const uint32_t hi = tf32_rna(v);
    const uint32_t lo = tf32_rna(v - __uint_as_float(hi));
    const int pos = ((k >> 2) * npad + n) * 4 + (k & 3);
    w_s[pos] = __uint_as_float(hi);
    w_s[kk * npad + pos] = __uint_as_float(lo);
  }
  for (int i = threadIdx.x; i < npad; i += blockDim.x) w2_s[i] = i < a.attn ? __ldg(a.w2 + i) : 0.f;
  if (threadIdx.x == 0) w2_s[npad] = __ldg(a.b2);
  for (int p = threadIdx.x; p < a.pairs; p += blockDim.x) {
    int i, j;
    pair_from_index(p, a.fields, i, j);
    ptab[p] = (i << 16) | j;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s)
      for (int b = 0; b < 2; ++b) {
        mbar_init(a_ready(s, b), 4);   // one arrive per epilogue warp of the slot
        mbar_init(d_full(s, b), 1);    // tcgen05.commit
      }
    fence_barrier_init();
  }
  if (warp == slots * 4) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int bc = a.buf_cols;                       // columns of one buffer: D [0, npad) | A hi [npad, +E) | A lo [.., +E)
  const uint32_t const_col = 2 * slots * bc;       // 8 columns [1, 0, .., 0] shared by all slots

  const int64_t tiles = (a.rows + 127) / 128;
  const int64_t stride = (int64_t)gridDim.x * slots;

  if (warp >= slots * 5) return;                   // (the block is sized for kMaxSlots)
  if (!is_mma) {
    // =========================== epilogue warps: slot = warp / 4, TMEM lane quarter = warp % 4 ====================
    const int slot = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(32 * q) << 16);
    {
      uint32_t one[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) one[j] = j == 0 ? __float_as_uint(1.0f) : 0u;
      tmem_st16(lane_base + const_col, one);
      tmem_st_wait();
    }
    // The x rows of the samples a tile touches are staged in shared memory (coalesced cp.async by the slot's 128
    // threads, one tile ahead): gathering x_i / x_j per thread straight from global memory costs a tag lookup per lane
    // and instruction -- the L1 pipe was the limit (613 cycles per tile).
    float* xs_slot = xs_all + (size_t)slot * 2 * xbuf_floats;
    const int slot_tid = r;                                   // 0 .. 127 within the slot
    const int chunks_per_sample = a.fields * (E / 4);
    // (all row arithmetic in 32 bits: rows < 2^31 is checked on the host; 64-bit divisions by `pairs` were a quarter of
    //  the kernel's instructions)
    const uint32_t n_rows = static_cast<uint32_t>(a.rows), n_pairs = static_cast<uint32_t>(a.pairs);
    auto prefetch_x = [&](int64_t tile, int buf) {
      if (tile < tiles) {
        const uint32_t first = static_cast<uint32_t>(tile) * 128u;
        const uint32_t last = first + 127u < n_rows - 1u ? first + 127u : n_rows - 1u;
        const uint32_t b_lo = first / n_pairs;
        const int count = static_cast<int>(last / n_pairs - b_lo) + 1;
        const uint32_t dst0 = smem_u32(xs_slot + buf * xbuf_floats);
        const float* src0 = a.x + (size_t)b_lo * a.fields * E;
        for (int i = slot_tid; i < count * chunks_per_sample; i += 128) {
          const int f = i / (E / 4), c = i - f * (E / 4);         // f = sample-local field row index (s * fields + n)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (f * kPitch + 4 * c) * 4),
                       "l"(src0 + (size_t)i * 4) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto build = [&](int64_t tile, int buf) {      // prod of this thread's (sample, pair) row -> A operand of `buf`
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");   // the tile's x rows are in; the other buffer is free
      prefetch_x(tile + stride, buf ^ 1);
      const uint32_t first = static_cast<uint32_t>(tile) * 128u;
      const uint32_t row = first + r;
      const uint32_t t_hi = lane_base + (2 * slot + buf) * bc + npad, t_lo = t_hi + E;
      const bool live = row < n_rows;
      const uint32_t row32 = live ? row : first;
      const uint32_t b_lo = first / n_pairs;
      const uint32_t b = row32 / n_pairs;
      const int p = static_cast<int>(row32 - b * n_pairs);
      const int ij = ptab[p];
      const float* xb = xs_slot + buf * xbuf_floats + static_cast<int>(b - b_lo) * a.fields * kPitch;
      const float* xi = xb + (ij >> 16) * kPitch;
      const float* xj = xb + (ij & 0xffff) * kPitch;
#pragma unroll
      for (int c = 0; c < E; c += 16) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 u = *reinterpret_cast<const float4*>(xi + c + 4 * j), w = *reinterpret_cast<const float4*>(xj + c + 4 * j);
          v[4 * j] = u.x * w.x;            // (rows beyond the batch repeat the tile's first row: their score is
          v[4 * j + 1] = u.y * w.y;        //  never stored)
          v[4 * j + 2] = u.z * w.z;
          v[4 * j + 3] = u.w * w.w;
        }
        write_operand16(t_hi + c, t_lo + c, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready(slot, buf));
    };
    const int64_t tile0 = (int64_t)blockIdx.x * slots + slot;
    prefetch_x(tile0, 0);
    if (tile0 < tiles) build(tile0, 0);
    int64_t k = 0;
    for (int64_t tile = tile0; tile < tiles; tile += stride, ++k) {
      const int buf = static_cast<int>(k & 1);
      if (tile + stride < tiles) build(tile + stride, buf ^ 1);   // the next tile's operand before this tile's wait
      mbar_wait(d_full(slot, buf), static_cast<uint32_t>((k >> 1) & 1));
      tc_fence_after();
      const uint32_t t_d = lane_base + (2 * slot + buf) * bc;
      float score = w2_s[npad];
#pragma unroll
      for (int c = 0; c < 32; c += 16) {
        if (c < npad) {
          uint32_t raw[16];
          tmem_ld16(t_d + c, raw);
#pragma unroll
          for (int j = 0; j < 16; ++j) score = fmaf(fmaxf(__uint_as_float(raw[j]), 0.f), w2_s[c + j], score);
        }
      }
      tc_fence_before();
      const int64_t row = tile * 128 + r;
      if (row < a.rows) a.scores[row] = score;
    }
  } else {
    // =========================== MMA issuers: one warp per slot ===========================================================
    const int s = warp - slots * 4;
    const uint32_t idesc = umma_idesc_tf32(npad);
    const uint32_t lbo = npad * 16;
    const uint64_t b_hi0 = umma_desc(smem_u32(w_s), lbo, 128);
    const uint64_t b_lo0 = b_hi0 + ((kk * npad * 4) >> 4);
    const uint32_t step_u = (2 * lbo) >> 4;
    const uint32_t a_const = tmem_base + const_col;
    int64_t k = 0;
    for (int64_t tile = (int64_t)blockIdx.x * slots + s; tile < tiles; tile += stride, ++k) {
      const int buf = static_cast<int>(k & 1);
      const uint32_t d = tmem_base + (2 * s + buf) * bc;
      mbar_wait(a_ready(s, buf), static_cast<uint32_t>((k >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < E / 8; ++ks) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {   // 0: A_lo*B_hi, 1: A_hi*B_lo, 2: A_hi*B_hi
            const uint32_t a_op = d + npad + (term == 0 ? E : 0) + 8 * ks;
            umma_tf32_ts(d, a_op, (term == 1 ? b_lo0 : b_hi0) + ks * step_u, idesc, (ks > 0 || term > 0) ? 1u : 0u);
          }
        }
        const uint32_t kb = (E / 8) * step_u;      // the bias k-step: A = [1, 0, ..]
        umma_tf32_ts(d, a_const, b_lo0 + kb, idesc, 1u);
        umma_tf32_ts(d, a_const, b_hi0 + kb, idesc, 1u);
        umma_commit(d_full(s, buf));
      }
      __syncwarp();
    }
  }
  // (no CTA-wide barrier below: surplus warps have left; TMEM is released by the allocating warp after a named barrier)
  tc_fence_before();
  asm volatile("bar.sync 1, %0;" ::"r"(slots * 5 * 32) : "memory");
  if (warp == slots * 4) tmem_dealloc(tmem_base, 512);
}

// softmax over the P raw scores of a sample (in place) and out = sum_p scores_p * x_i * x_j.  One warp per sample; the
// weighted sum is taken as out[e] = sum_i x_i[e] * y_i[e] with y_i = sum_{j>i} s_ij x_j (the upper-triangular score matrix
// times X): lane = (row i of a group of 32 / (E/4) rows, 16-byte chunk of e), so per j the warp reads ONE row of X
// (broadcast) and one score per row -- half the shared-memory traffic of forming every pair product per lane.
template <int E>
__global__ void __launch_bounds__(256) afm_finish_kernel(const float* __restrict__ x, float* __restrict__ scores,
                                                         int64_t batch, int fields, int pairs, float* __restrict__ out) {
  extern __shared__ __align__(16) float fin_smem[];     // [8 warps]: X [fields][E + 4] | scores [pairs rounded up to 4]
  constexpr int kPitch = E + 4, kChunks = E / 4, kGroup = 32 / kChunks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sp = (pairs + 3) & ~3;
  float* mine = fin_smem + (size_t)warp * (fields * kPitch + sp);
  float* ss = mine + fields * kPitch;
  const int ri = lane / kChunks, ec = lane - ri * kChunks;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < batch; b += (int64_t)gridDim.x * 8) {
    __syncwarp();
    for (int i = lane; i < fields * kChunks; i += 32) {
      const int f = i / kChunks, c = i - f * kChunks;
      *reinterpret_cast<float4*>(mine + f * kPitch + 4 * c) = __ldg(reinterpret_cast<const float4*>(x + (b * fields + f) * E) + c);
    }
    float* sc = scores + b * pairs;
    float mx = -INFINITY;
    for (int p = lane; p < pairs; p += 32) {
      const float v = sc[p];
      ss[p] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int p = lane; p < pairs; p += 32) {
      const float e = expf(ss[p] - mx);
      ss[p] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int p = lane; p < pairs; p += 32) {
      const float v = ss[p] * inv;
      ss[p] = v;
      sc[p] = v;
    }
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g0 = 0; g0 < fields - 1; g0 += kGroup) {
      const int i = g0 + ri;
      const bool row_on = i < fields - 1;
      const float* srow = ss + (row_on ? i * (2 * fields - i - 1) / 2 - i - 1 : 0);   // srow[j] = s_ij for j > i
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = g0 + 1; j < fields; ++j) {
        const float4 xj = *reinterpret_cast<const float4*>(mine + j * kPitch + 4 * ec);
        const float sij = (row_on && j > i) ? srow[j] : 0.f;
        y.x = fmaf(sij, xj.x, y.x); y.y = fmaf(sij, xj.y, y.y); y.z = fmaf(sij, xj.z, y.z); y.w = fmaf(sij, xj.w, y.w);
      }
      if (row_on) {
        const float4 xi = *reinterpret_cast<const float4*>(mine + i * kPitch + 4 * ec);
        acc.x = fmaf(xi.x, y.x, acc.x); acc.y = fmaf(xi.y, y.y, acc.y);
        acc.z = fmaf(xi.z, y.z, acc.z); acc.w = fmaf(xi.w, y.w, acc.w);
      }
    }
#pragma unroll
    for (int o = kChunks; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (ri == 0) *reinterpret_cast<float4*>(out + b * E + 4 * ec) = acc;
  }
}

template <int E>
int launch(AfmTc5Args a, int64_t batch, float* out, cudaStream_t s) {
  const int npad = a.npad;
  a.buf_cols = npad + 2 * E;
  int slots = (512 - 8) / (2 * a.buf_cols);
  if (slots > kMaxSlots) slots = kMaxSlots;
  a.slots = slots;
  a.smax = 127 / a.pairs + 2;
  size_t smem = ((size_t)2 * (E + 8) * npad + npad + 4 + a.pairs) * 4 + 8 + (4 * kMaxSlots) * 8 + 16 + 128 + 16;
  while (slots > 1 && smem + (size_t)slots * 2 * a.smax * a.fields * (E + 4) * 4 > (size_t)kMaxDynSmem) --slots;
  if (smem + (size_t)slots * 2 * a.smax * a.fields * (E + 4) * 4 > (size_t)kMaxDynSmem) return TRS_ERR_UNSUPPORTED;
  a.slots = slots;
  smem += (size_t)slots * 2 * a.smax * a.fields * (E + 4) * 4;
  const int64_t tiles = (a.rows + 127) / 128;
  const int64_t want = (tiles + slots - 1) / slots;
  const int grid = static_cast<int>(want < kNumSMs ? want : kNumSMs);
  TRS_SMEM_OPT_IN((afm_tc5_kernel<E>));
  afm_tc5_kernel<E><<<grid, slots * 160, smem, s>>>(a);
  int rc = check_launch("afm_tc5_kernel");
  if (rc != TRS_OK) return rc;
  const int64_t want2 = (batch + 7) / 8;
  const size_t smem2 = (size_t)8 * (a.fields * (E + 4) + ((a.pairs + 3) & ~3)) * 4;
  TRS_SMEM_OPT_IN((afm_finish_kernel<E>));
  afm_finish_kernel<E><<<static_cast<int>(want2 < kNumSMs * 4 ? want2 : kNumSMs * 4), 256, smem2, s>>>(
      a.x, a.scores, batch, a.fields, a.pairs, out);
  return check_launch("afm_finish_kernel");
}

}  // namespace

// AttentionalFactorizationMachineLayer on tcgen05; TRS_ERR_UNSUPPORTED when the shape is not covered
int afm_tc5_launch(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, int64_t batch,
                   int fields, int embed, int attn, float* out, float* scores, cudaStream_t s) {
  static const bool disabled = getenv("TRS_DISABLE_TC5") != nullptr || getenv("TRS_DISABLE_TC") != nullptr;
  const int64_t pairs = (int64_t)fields * (fields - 1) / 2;
  if (disabled || !(embed == 16 || embed == 32) || attn < 1 || attn > 32 || fields < 2 || fields > kMaxFields) return TRS_ERR_UNSUPPORTED;
  if (batch * pairs < 128 * 148 || batch * pairs >= (int64_t(1) << 31) || !aligned16(x) || !aligned16(out)) return TRS_ERR_UNSUPPORTED;
  AfmTc5Args a{};
  a.x = x; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.scores = scores;
  a.rows = batch * pairs; a.fields = fields; a.pairs = static_cast<int>(pairs); a.attn = attn;
  a.npad = attn <= 16 ? 16 : 32;
  return embed == 16 ? launch<16>(a, batch, out, s) : launch<32>(a, batch, out, s);
}

}  // namespace trs
