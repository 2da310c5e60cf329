// a8 on the 5th-generation tensor cores: Compress Interaction Network with tcgen05.mma (kind::tf32), accumulators in
// TMEM, FP32-accurate through the 3xTF32 split (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate).
//
// GEMM view per layer: D[(b,e), c] = sum_{x,y} x0[(b,e), x] * h[(b,e), y] * W[c, x*H + y].
//   * rows (b,e) are independent, so activations are kept ROW-major between layers: x0T [(b,e)][Hp0] (one transpose of
//     the (B,N,E) input) and h_l [(b,e)][Hp_l] -- which is exactly the accumulator layout (TMEM lane = row, column =
//     channel), so the epilogue writes full 128-byte lines and the next layer reads contiguous rows;
//   * the K index is re-ordered y-chunk-major: k' = (yc*N + x)*16 + (y%16).  A producer thread owns one row, keeps
//     the 16 h values of the current y-chunk in registers across the N chunks that share it, and per chunk emits
//     z = x0[x] * h[y] split into TF32 hi/lo straight into the UMMA canonical (K-major, no swizzle) smem layout;
//     the A operand never exists in HBM.  The weights are pre-split and pre-permuted once per call into the same
//     chunked layout (cin_tc_prep_weights) and stream from L2 with one cp.async.bulk per chunk;
//   * CTA tile = 256 rows x Npad channels (two M=128 accumulators = up to 512 TMEM columns), BK = 16 per stage;
//     per chunk 2 halves x 2 k-steps x 3 split terms = 12 tcgen05.mma; tcgen05.commit frees the stage / signals
//     the epilogue; roles: warps 0-7 = A producers then epilogue (tcgen05.ld -> folded Conv-bias + eval-BN + act ->
//     hidden half stored row-major, direct half summed over e with shuffles -> pooled), warp 8 = MMA issuer,
//     warp 9 = weight loader;
//   * the dead hidden half of the last layer (computed and discarded by the reference) is not computed.
// Shapes outside (E in {8,16,32}, channels <= 256 per layer) use the FFMA path in cin.cu.
#include <stdlib.h>

#include "tc5.cuh"

namespace trs {
namespace {

constexpr int kTileM = 256;          // rows per CTA tile (2 x UMMA_M 128)
constexpr int kBK = 16;              // K per stage = 4 sixteen-byte chunks = 2 UMMA k-steps
constexpr int kProducerThreads = 256;
constexpr int kThreads = kProducerThreads + 64;   // + MMA warp + weight-loader warp
constexpr int kAStages = 2;          // operand-A ring in smem (32 KB per stage; generation is cheap, two stages suffice)
constexpr int kMaxAStages = 4;       // (four 64-column stages when the A ring lives in tensor memory)
constexpr int kMaxBStages = 8;       // weight ring: deep, the L2 -> smem stream is latency-bound

using namespace tc5;

// ---- preparation kernels ------------------------------------------------------------------------------------------------
// x (B, N, E) -> xt [(b,e)][hp0] row-major with zero padding (layer-0 "h" and the x0 operand of every layer)
__global__ void __launch_bounds__(256) cin_tc_transpose_kernel(const float* __restrict__ x, int64_t batch, int fields,
                                                               int embed, int hp0, float* __restrict__ xt) {
  const int64_t items = batch * embed * hp0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / hp0;
    const int n = static_cast<int>(i - m * hp0);
    const int64_t b = m / embed;
    const int e = static_cast<int>(m - b * embed);
    xt[i] = n < fields ? __ldg(x + (b * fields + n) * embed + e) : 0.f;
  }
}

// W (C, N*H) -> wp[q = yc*N + x][hi|lo][kc 0..3][n 0..npad-1][4 floats]  with k' = q*16 + kc*4 + j,  y = yc*16 + kc*4 + j
// fold != 0 (layer 0, where h = x0 and z[x, y] = z[y, x]): the pair (x, y) is kept only for y >= x, with the weight
// W[c, x, y] + W[c, y, x] (W[c, x, x] on the diagonal); the kernel then skips the chunks that lie below the diagonal.
__global__ void __launch_bounds__(256) cin_tc_prep_weights_kernel(const float* __restrict__ w, int c_begin, int c_eff, int fields,
                                                                  int h_prev, int hp, int npad, int fold,
                                                                  float* __restrict__ wp) {
  const int chunks = (hp / 16) * fields;
  const int64_t items = (int64_t)chunks * 4 * npad * 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = static_cast<int>(i & 3);
    const int n = static_cast<int>((i >> 2) % npad);
    const int kc = static_cast<int>(((i >> 2) / npad) & 3);
    const int q = static_cast<int>((i >> 2) / npad / 4);
    const int yc = q / fields, xf = q - yc * fields;
    const int y = yc * 16 + kc * 4 + j;
    float v = 0.f;
    if (n < c_eff && y < h_prev) {
      const float* wc = w + (int64_t)(c_begin + n) * fields * h_prev;
      if (!fold) v = __ldg(wc + xf * h_prev + y);
      else if (y > xf) v = __ldg(wc + xf * h_prev + y) + __ldg(wc + y * h_prev + xf);
      else if (y == xf) v = __ldg(wc + xf * h_prev + y);
    }
    const uint32_t hi = tf32_rna(v);
    const uint32_t lo = tf32_rna(v - __uint_as_float(hi));
    const int64_t base = (int64_t)q * (2 * 4 * npad * 4);
    wp[base + ((0 * 4 + kc) * npad + n) * 4 + j] = __uint_as_float(hi);
    wp[base + ((1 * 4 + kc) * npad + n) * 4 + j] = __uint_as_float(lo);
  }
}

struct CinTcArgs {
  const float* xt;       // [(b,e)][hp0]  (x0 operand; also layer-0 h)
  const float* h;        // [(b,e)][hp]   input activations of this layer
  const float* wp;       // prepared weights
  const float* scale;    // per channel (indexed c_begin + c), null = 1
  const float* shift;    // per channel, null = 0
  float* h_next;         // [(b,e)][hp_next] or null
  float* pooled;         // (B, pooled_width)
  int64_t m_rows;        // B * E
  int fields, embed, hp0, hp, npad, c_eff;   // c_eff = channels computed by this pass
  int c_begin;                                // first channel of this pass (channel blocks of a wide layer)
  int n_direct, hid_begin, hid_count, hp_next;
  int pool_off, pooled_width, act, b_stages;
  int h_pitch, k_valid;    // row pitch of h in floats and its valid columns (loads beyond read as zero); CIN: both = hp
  int fold;                // layer 0 of a CIN (h == x0): only the chunks with some y >= x are walked (see the weight prep)
  int c_total;             // gridDim.y > 1: channel blocks of `c_block` over c_total channels, one block per blockIdx.y
  int c_block;
  int64_t wp_pass_stride;  // floats between the prepared weights of consecutive channel blocks
};

// kATmem = true (npad <= 128): the generated A operand goes to TENSOR MEMORY (columns 256..511, four 64-column stages:
//   per half hi[16] | lo[16]) with tcgen05.st, the accumulators use columns 0..255, and only the weights stream from
//   shared memory -- the SS form with both operands in shared memory is smem-port bound for N = 128
//   (profiles/r01_cin_tcgen05_notes.md).  kATmem = false (npad = 256): accumulators need all 512 columns, A in smem.
template <bool kATmem>
__global__ void __launch_bounds__(kThreads, 1) cin_tc_layer_kernel(CinTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (gridDim.y > 1) {   // wide dense layer: this CTA column owns one channel block
    a.c_begin = blockIdx.y * a.c_block;
    a.c_eff = a.c_total - a.c_begin < a.c_block ? a.c_total - a.c_begin : a.c_block;
    a.npad = (a.c_eff + 31) & ~31;
    a.wp += blockIdx.y * a.wp_pass_stride;
  }
  constexpr int kAS = kATmem ? 4 : kAStages;                   // A ring depth
  const int npad = a.npad;
  const int dstride = kATmem ? 128 : npad;                     // TMEM columns between the two accumulator halves
  const uint32_t a_stage_bytes = kATmem ? 0 : 2 * 4 * kTileM * 16;   // hi/lo x 4 chunks x 256 rows x 16 B = 32 KB
  const uint32_t b_stage_bytes = 2 * 4 * npad * 16;            // hi/lo x 4 chunks x npad rows x 16 B
  unsigned char* a_smem = smem_raw;
  unsigned char* b_smem = a_smem + (size_t)kAS * a_stage_bytes;
  float* x0_s = reinterpret_cast<float*>(b_smem + (size_t)a.b_stages * b_stage_bytes);   // [fields][256]
  float* ss_s = x0_s + (size_t)a.fields * kTileM;                                        // scale[npad], shift[npad]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ss_s + 2 * npad);
  // barriers: full_a[kMaxAStages], empty_a[kMaxAStages], full_b[kMaxBStages], empty_b[kMaxBStages], acc_full, acc_empty
  constexpr int kNumBars = 2 * kMaxAStages + 2 * kMaxBStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  const uint32_t bar0 = smem_u32(bars);
  auto full_a = [&](int s) { return bar0 + 8u * s; };
  auto empty_a = [&](int s) { return bar0 + 8u * (kMaxAStages + s); };
  auto full_b = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + s); };
  auto empty_b = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + kMaxBStages + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kMaxAStages + 2 * kMaxBStages), acc_empty = acc_full + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kAS; ++s) {
      mbar_init(full_a(s), kProducerThreads / 32);   // one arrive per producer warp
      mbar_init(empty_a(s), 1);
    }
    for (int s = 0; s < a.b_stages; ++s) {
      mbar_init(full_b(s), 1);
      mbar_init(empty_b(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kProducerThreads / 32);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    ss_s[i] = i < a.c_eff ? (a.scale ? __ldg(a.scale + a.c_begin + i) : 1.f) : 0.f;
    ss_s[npad + i] = (i < a.c_eff && a.shift) ? __ldg(a.shift + a.c_begin + i) : 0.f;
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t tiles = (a.m_rows + kTileM - 1) / kTileM;
  const int ychunks = a.hp / 16;
  // fields walked with y-chunk yc: all of them, or (folded layer 0) those with x <= the chunk's last y
  auto x_count = [&](int yc) { return a.fold ? (16 * yc + 16 < a.fields ? 16 * yc + 16 : a.fields) : a.fields; };
  int chunks = 0;                          // K' / 16
  for (int yc = 0; yc < ychunks; ++yc) chunks += x_count(yc);
  // Every CTA walks the K loop from its own starting point (the sum is order independent): otherwise all 148 CTAs
  // stream the SAME weight chunk from L2 at the same moment and hot-spot a few L2 slices.
  const int yc_rot = blockIdx.x % ychunks, x_rot = (blockIdx.x * 5) % a.fields;

  if (warp < 8) {
    // =========================== A producers, then epilogue =====================================================
    const int r = threadIdx.x;   // row within the tile
    int sa = 0;                  // A ring position and phase
    uint32_t pa = 0;
    uint32_t tile_n = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
      const int64_t m = tile * kTileM + r;
      const bool row_ok = m < a.m_rows;
      for (int xf = 0; xf < a.fields; ++xf) x0_s[xf * kTileM + r] = row_ok ? (a.xt ? __ldg(a.xt + m * a.hp0 + xf) : 1.f) : 0.f;   // xt == null: x0 = 1 (plain dense layer)
      // kPF y-chunks of this row's h are in flight in registers: with few fields (one, for a plain dense layer) a chunk
      // lasts less than a global-load latency, so the loads run that many chunks ahead of their use
      constexpr int kPF = kATmem ? 3 : 4;   // (register budget: the TMEM form also holds hi[16] / lo[16])
      float hbuf[kPF][16];
      auto load_h = [&](float (&dst)[16], int yci) {
        const int yc = yci + yc_rot < ychunks ? yci + yc_rot : yci + yc_rot - ychunks;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && yci < ychunks && yc * 16 + 4 * v4 < a.k_valid)
            v = __ldg(reinterpret_cast<const float4*>(a.h + m * a.h_pitch + yc * 16) + v4);
          dst[4 * v4 + 0] = v.x; dst[4 * v4 + 1] = v.y; dst[4 * v4 + 2] = v.z; dst[4 * v4 + 3] = v.w;
        }
      };
#pragma unroll
      for (int p = 0; p < kPF; ++p) load_h(hbuf[p], p);
      for (int y0 = 0; y0 < ychunks; y0 += kPF) {
#pragma unroll
      for (int p = 0; p < kPF; ++p) {
        if (y0 + p >= ychunks) break;
        float hreg[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) hreg[j] = hbuf[p][j];
        load_h(hbuf[p], y0 + p + kPF);
        const int yc_now = y0 + p + yc_rot < ychunks ? y0 + p + yc_rot : y0 + p + yc_rot - ychunks;
        const int xn = x_count(yc_now), xr = x_rot % xn;
        for (int xi = 0; xi < xn; ++xi) {
          const int xf = xi + xr < xn ? xi + xr : xi + xr - xn;
          mbar_wait(empty_a(sa), pa ^ 1);   // stage free (first pass: passes immediately)
          const float xv = x0_s[xf * kTileM + r];
          if (kATmem) {
            // 16 z values of this row -> TMEM columns [hi 0..15 | lo 16..31] of this half's slot in stage sa
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float z = xv * hreg[j];
              hi[j] = (__float_as_uint(z) + 0x1000u) & 0xffffe000u;
              lo[j] = (__float_as_uint(z - __uint_as_float(hi[j])) + 0x1000u) & 0xffffe000u;
            }
            const uint32_t ta = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + 256 + sa * 64 +
                                (warp >> 2) * 32;
            tmem_st16(ta, hi);
            tmem_st16(ta + 16, lo);
            tmem_st_wait();
            tc_fence_before();
          } else {
          unsigned char* dst = a_smem + (size_t)sa * a_stage_bytes + r * 16;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float z = xv * hreg[4 * kc + j];
              hi[j] = (__float_as_uint(z) + 0x1000u) & 0xffffe000u;   // round to TF32 (full-rate ALU)
              // exact remainder, itself rounded to TF32 (the MMA would otherwise truncate it: 2^-21 -> 2^-22 |z|)
              lo[j] = (__float_as_uint(z - __uint_as_float(hi[j])) + 0x1000u) & 0xffffe000u;
            }
            *reinterpret_cast<uint4*>(dst + (0 * 4 + kc) * (kTileM * 16)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(dst + (1 * 4 + kc) * (kTileM * 16)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
          fence_proxy_async();   // make the generic-proxy stores visible to the tensor core (async proxy)
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(full_a(sa));
          if (++sa == kAS) { sa = 0; pa ^= 1; }
        }
      }
      }
      // ---- epilogue of this tile: warps 0-3 -> accumulator half 0 (rows 0..127), warps 4-7 -> half 1 ------------
      mbar_wait(acc_full, tile_n & 1);
      tc_fence_after();
      const int half = warp >> 2;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + half * dstride;
      const int64_t b = row_ok ? m / a.embed : 0;
      const int e = row_ok ? static_cast<int>(m - b * a.embed) : 0;
      for (int c0 = 0; c0 < npad; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(taddr + c0, raw);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = apply_act(fmaf(__uint_as_float(raw[j]), ss_s[c0 + j], ss_s[npad + c0 + j]), a.act);
        // hidden half -> next layer's activations, row-major
        if (a.h_next != nullptr) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int c = a.c_begin + c0 + 4 * j4;   // global channel
            if (row_ok && c >= a.hid_begin && c + 3 < a.hid_begin + a.hid_count && c0 + 4 * j4 + 3 < a.c_eff &&
                (a.hp_next & 3) == 0)
              *reinterpret_cast<float4*>(a.h_next + m * a.hp_next + (c - a.hid_begin)) =
                  make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
            else if (row_ok) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j >= a.hid_begin && c + j < a.hid_begin + a.hid_count && c0 + 4 * j4 + j < a.c_eff)
                  a.h_next[m * a.hp_next + (c + j - a.hid_begin)] = v[4 * j4 + j];
            }
          }
        }
        // direct half -> sum over the `embed` rows of each sample (consecutive lanes), one writer per (b, c)
        if (a.c_begin + c0 < a.n_direct) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float sum = row_ok ? v[j] : 0.f;
            for (int o = a.embed >> 1; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (row_ok && e == 0 && a.c_begin + c0 + j < a.n_direct && c0 + j < a.c_eff)
              a.pooled[b * a.pooled_width + a.pool_off + a.c_begin + c0 + j] = sum;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  } else if (warp == 8) {
    // =========================== MMA issuer ===============================================================================
    const uint32_t idesc = umma_idesc_tf32(npad);
    const uint32_t a_lbo = kTileM * 16, b_lbo = npad * 16;
    // descriptors differ only in their 14-bit start-address field (16-byte units, shared memory < 256 KB): build the
    // two base descriptors once and add offsets, so that issuing an MMA costs a few integer adds in the one issuing
    // thread (with N = 128 an MMA lasts only ~64 cycles; descriptor arithmetic was the limiter)
    const uint64_t a_desc0 = umma_desc(smem_u32(a_smem), a_lbo, 128);
    const uint64_t b_desc0 = umma_desc(smem_u32(b_smem), b_lbo, 128);
    const uint32_t a_stage_u = a_stage_bytes >> 4, b_stage_u = b_stage_bytes >> 4;
    const uint32_t a_ks_u = (2 * a_lbo) >> 4, b_ks_u = (2 * b_lbo) >> 4;      // next k-step = two 16-byte chunks on
    const uint32_t a_lo_u = (4 * a_lbo) >> 4, b_lo_u = (4 * b_lbo) >> 4;      // hi -> lo plane
    const uint32_t a_half_u = (128 * 16) >> 4;                                // rows 128..255
    uint32_t tile_n = 0, pa = 0, pb = 0;
    int sa = 0, sb = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
      mbar_wait(acc_empty, (tile_n & 1) ^ 1);   // epilogue of the previous tile has drained TMEM
      tc_fence_after();
      for (int q = 0; q < chunks; ++q) {
        mbar_wait(full_a(sa), pa);
        mbar_wait(full_b(sb), pb);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = a_desc0 + sa * a_stage_u;
          const uint64_t bd = b_desc0 + sb * b_stage_u;
          const uint32_t acc0 = q > 0 ? 1u : 0u;
          // issue order alternates between the two accumulator halves: consecutive MMAs into the SAME TMEM tile
          // serialise on the accumulator (measured ~50 cycles per MMA), the halves are independent
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t b_hi = bd + ks * b_ks_u, b_lo = b_hi + b_lo_u;
#pragma unroll
            for (int term = 0; term < 3; ++term) {   // 0: A_lo*B_hi, 1: A_hi*B_lo, 2: A_hi*B_hi
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint32_t d = tmem_base + half * dstride;
                const uint32_t accf = (ks > 0 || term > 0) ? 1u : acc0;
                const uint64_t bsel = term == 1 ? b_lo : b_hi;
                if (kATmem) {
                  const uint32_t ta_hi = tmem_base + 256 + sa * 64 + half * 32 + 8 * ks;
                  umma_tf32_ts(d, term == 0 ? ta_hi + 16 : ta_hi, bsel, idesc, accf);
                } else {
                  const uint64_t a_hi = ad + ks * a_ks_u + half * a_half_u;
                  umma_tf32(d, term == 0 ? a_hi + a_lo_u : a_hi, bsel, idesc, accf);
                }
              }
            }
          }
          umma_commit(empty_a(sa));                     // stages reusable once these MMAs have read them
          umma_commit(empty_b(sb));
          if (q == chunks - 1) umma_commit(acc_full);   // accumulators complete -> epilogue
        }
        __syncwarp();
        if (++sa == kAS) { sa = 0; pa ^= 1; }
        if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
      }
    }
  } else {
    // =========================== weight loader (bulk copies L2 -> smem) ======================================================
    int sb = 0;
    uint32_t pb = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int yci = 0; yci < ychunks; ++yci) {               // same walk as the producers
        const int yc = yci + yc_rot < ychunks ? yci + yc_rot : yci + yc_rot - ychunks;
        const int xn = x_count(yc), xr = x_rot % xn;
        for (int xi = 0; xi < xn; ++xi) {
        const int xf = xi + xr < xn ? xi + xr : xi + xr - xn;
        const int q = yc * a.fields + xf;
        mbar_wait(empty_b(sb), pb ^ 1);
        if (lane == 0) {
          mbar_expect_tx(full_b(sb), b_stage_bytes);
          bulk_g2s(smem_u32(b_smem + (size_t)sb * b_stage_bytes),
                   reinterpret_cast<const unsigned char*>(a.wp) + (size_t)q * b_stage_bytes, b_stage_bytes,
                   full_b(sb));
        }
        __syncwarp();
        if (++sb == a.b_stages) { sb = 0; pb ^= 1; }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct CinTcPlan {
  int hp0, hp_max, pooled_width;
  int64_t xt_floats, h_floats, pooled_floats, w_floats;
};

static CinTcPlan cin_tc_plan(int64_t batch, int fields, int embed, const int* sizes, int layers, int is_direct) {
  CinTcPlan p{};
  p.hp0 = round_up(fields, 16);
  int hp = p.hp0;
  for (int l = 0; l < layers; ++l) {
    const bool last = l == layers - 1;
    const int c_eff = (is_direct || last) ? sizes[l] : 2 * sizes[l];
    const int npad = round_up(c_eff, 32);
    p.w_floats += (int64_t)(hp / 16) * fields * 2 * 4 * npad * 4;
    p.pooled_width += sizes[l];
    hp = round_up(sizes[l], 16);
    if (!last && hp > p.hp_max) p.hp_max = hp;
  }
  const int64_t m = batch * embed;
  p.xt_floats = (m * p.hp0 + 63) / 64 * 64;
  p.h_floats = (m * p.hp_max + 63) / 64 * 64;
  p.pooled_floats = (batch * p.pooled_width + 63) / 64 * 64;
  p.w_floats = (p.w_floats + 63) / 64 * 64;
  return p;
}

int cin_tc_supported(int fields, int embed, const int* sizes, int layers, int is_direct) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  if (disabled) return 0;
  if (!(embed == 8 || embed == 16 || embed == 32) || fields < 1 || fields > 64 || layers < 1) return 0;
  for (int l = 0; l < layers; ++l) {
    const int c_eff = (is_direct || l == layers - 1) ? sizes[l] : 2 * sizes[l];
    if (sizes[l] < 1 || c_eff > 256) return 0;
    if (!is_direct && l < layers - 1 && sizes[l] % 4 != 0) return 0;   // hidden half must start 16-byte aligned
  }
  return 1;
}

int64_t cin_tc_workspace_bytes(int64_t batch, int fields, int embed, const int* sizes, int layers, int is_direct) {
  const CinTcPlan p = cin_tc_plan(batch, fields, embed, sizes, layers, is_direct);
  return (p.xt_floats + 2 * p.h_floats + p.pooled_floats + p.w_floats) * (int64_t)sizeof(float) + 1024;
}

// out (+)= fc( pooled );  implemented in cin.cu
int cin_fc_launch(const float* pooled, int pooled_width, const float* fc_w, const float* fc_b, int out_features,
                  int64_t batch, float* out, int accumulate, cudaStream_t s);

int cin_tc_run(const float* x, const float* const* conv_w, const float* const* scale, const float* const* shift,
               const int* sizes, int layers, int is_direct, int activation, const float* fc_w, const float* fc_b,
               int out_features, int64_t batch, int fields, int embed, float* out, int accumulate, void* workspace,
               int64_t workspace_bytes, cudaStream_t s) {
  const CinTcPlan p = cin_tc_plan(batch, fields, embed, sizes, layers, is_direct);
  TRS_REQUIRE(workspace_bytes >= cin_tc_workspace_bytes(batch, fields, embed, sizes, layers, is_direct),
              "cin: workspace too small for the tensor-core path");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  ws += (128 - (reinterpret_cast<uintptr_t>(ws) & 127)) & 127;
  float* xt = reinterpret_cast<float*>(ws);
  float* hbuf[2] = {xt + p.xt_floats, xt + p.xt_floats + p.h_floats};
  float* pooled = hbuf[1] + p.h_floats;
  float* wprep = pooled + p.pooled_floats;
  const int64_t m_rows = batch * embed;

  cin_tc_transpose_kernel<<<grid_for(m_rows * p.hp0, 256, 8), 256, 0, s>>>(x, batch, fields, embed, p.hp0, xt);
  int rc = check_launch("cin_tc_transpose_kernel");
  if (rc != TRS_OK) return rc;
  TRS_SMEM_OPT_IN(cin_tc_layer_kernel<true>);
  TRS_SMEM_OPT_IN(cin_tc_layer_kernel<false>);

  const float* h = xt;
  int hp = p.hp0, h_prev = fields, pool_off = 0;
  float* wp = wprep;
  for (int l = 0; l < layers; ++l) {
    const bool last = l == layers - 1;
    const int hl = sizes[l];
    const int c_eff = (is_direct || last) ? hl : 2 * hl;
    const int npad = round_up(c_eff, 32);
    const int64_t w_items = (int64_t)(hp / 16) * fields * 2 * 4 * npad * 4;
    // layer 0 contracts x0 with itself: z[x, y] = z[y, x], so only y >= x is walked with the two weights added
    // (K 39 x 48 -> 87 of 117 chunks at 39 fields).  TRS_CIN_NO_FOLD=1 switches it off (A/B measurements).
    static const bool no_fold = getenv("TRS_CIN_NO_FOLD") != nullptr;
    const int fold = (l == 0 && !no_fold) ? 1 : 0;
    cin_tc_prep_weights_kernel<<<grid_for(w_items / 2, 256, 8), 256, 0, s>>>(conv_w[l], 0, c_eff, fields, h_prev, hp,
                                                                             npad, fold, wp);
    rc = check_launch("cin_tc_prep_weights_kernel");
    if (rc != TRS_OK) return rc;
    CinTcArgs a{};
    a.xt = xt; a.h = h; a.wp = wp; a.scale = scale[l]; a.shift = shift[l];
    a.h_next = last ? nullptr : hbuf[l & 1];
    a.pooled = pooled;
    a.m_rows = m_rows; a.fields = fields; a.embed = embed; a.hp0 = p.hp0; a.hp = hp; a.npad = npad; a.c_eff = c_eff;
    a.n_direct = hl;
    a.hid_begin = is_direct ? 0 : hl;
    a.hid_count = last ? 0 : hl;
    a.hp_next = round_up(hl, 16);
    a.pool_off = pool_off; a.pooled_width = p.pooled_width; a.act = activation;
    a.h_pitch = hp; a.k_valid = hp; a.fold = fold;
    if (a.h_next != nullptr && a.hp_next != hl)   // zero the padding columns the next layer will read
      TRS_CUDA(cudaMemsetAsync(a.h_next, 0, (size_t)m_rows * a.hp_next * sizeof(float), s));
    const int64_t tiles = (m_rows + kTileM - 1) / kTileM;
    const int grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
    // A layer wider than 128 channels either runs as one pass with the A operand in shared memory (SS form: the shared-
    // memory port carries A reads + A stores + B reads + the weight stream, ~138 of 128 B/cycle wanted) or -- TRS_CIN_SPLIT=1
    // -- as two passes of <= 128 channels each with the A operand in tensor memory (the operand is generated twice).
    // Measured (CIN [128,128], B = 65 536): one pass 9.09 ms, two passes 9.42 ms -- the split stays an experiment.
    static const bool split_wide = getenv("TRS_CIN_SPLIT") != nullptr;
    const int passes = (npad > 128 && split_wide) ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
      CinTcArgs ap = a;
      if (passes == 2) {
        ap.c_begin = pass * 128;
        ap.c_eff = c_eff - ap.c_begin < 128 ? c_eff - ap.c_begin : 128;
        ap.npad = round_up(ap.c_eff, 32);
        const int64_t items = (int64_t)(hp / 16) * fields * 2 * 4 * ap.npad * 4;
        float* wpp = wp + (pass == 0 ? 0 : (int64_t)(hp / 16) * fields * 2 * 4 * 128 * 4);
        cin_tc_prep_weights_kernel<<<grid_for(items / 2, 256, 8), 256, 0, s>>>(conv_w[l], ap.c_begin, ap.c_eff, fields,
                                                                               h_prev, hp, ap.npad, fold, wpp);
        rc = check_launch("cin_tc_prep_weights_kernel");
        if (rc != TRS_OK) return rc;
        ap.wp = wpp;
      }
      const int np = ap.npad;
      const bool a_tmem = np <= 128;   // accumulators leave 256 TMEM columns free: A operand goes to tensor memory
      const size_t a_stage = 2 * 4 * kTileM * 16, b_stage = (size_t)2 * 4 * np * 16;
      const size_t fixed = (a_tmem ? 0 : kAStages * a_stage) + (size_t)fields * kTileM * 4 + 2 * np * 4 +
                           (2 * kMaxAStages + 2 * kMaxBStages + 2) * 8 + 16 + 128;
      int b_stages = kMaxBStages;
      while (b_stages > 2 && b_stages * b_stage + fixed > (size_t)kMaxDynSmem) --b_stages;
      static const int cap_b = getenv("TRS_CIN_B_STAGES") ? atoi(getenv("TRS_CIN_B_STAGES")) : 0;   // experiments
      if (cap_b >= 2 && cap_b < b_stages) b_stages = cap_b;
      const size_t smem = b_stages * b_stage + fixed;
      TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "cin: tensor-core tile does not fit shared memory");
      ap.b_stages = b_stages;
      if (getenv("TRS_CIN_VERBOSE")) fprintf(stderr, "cin layer %d pass %d: npad %d, a_tmem %d, b_stages %d, smem %zu\n", l, pass, np, (int)a_tmem, b_stages, smem);
      if (a_tmem) cin_tc_layer_kernel<true><<<grid, kThreads, smem, s>>>(ap);
      else cin_tc_layer_kernel<false><<<grid, kThreads, smem, s>>>(ap);
      rc = check_launch("cin_tc_layer_kernel");
      if (rc != TRS_OK) return rc;
    }
    h = a.h_next;
    hp = a.hp_next;
    h_prev = hl;
    pool_off += hl;
    wp += w_items;
  }
  return cin_fc_launch(pooled, p.pooled_width, fc_w, fc_b, out_features, batch, out, accumulate, s);
}


// ---- a plain dense layer out = act(x W^T + b) on the same kernel: a CIN layer with ONE field and x0 = 1 ----------------
// x (rows, K) row-major with K % 4 == 0, W (C, K), out (rows, C).  C <= 128 runs the TMEM-operand form; wider layers are
// cut into equal channel blocks of at most 256 (one block per blockIdx.y, the SMs divided between the blocks), each CTA
// streaming only its block of the weights.  Scratch for the pre-split weights comes from the stream-ordered pool.
int dense_tc_supported(int k_dim, int c_dim, const void* x, const void* out) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  return !disabled && k_dim % 4 == 0 && k_dim >= 16 && c_dim >= 16 && aligned16(x) && aligned16(out);
}

int dense_tc_run(const float* x, int64_t rows, int k_dim, const float* w, const float* bias, int c_dim, int activation,
                 float* out, cudaStream_t s) {
  if (rows == 0) return TRS_OK;
  const int kp = round_up(k_dim, 16);
  const int passes = (c_dim + 255) / 256;
  const int block = round_up((c_dim + passes - 1) / passes, 32);
  const bool a_tmem = block <= 128;
  const size_t w_floats = (size_t)(kp / 16) * 2 * 4 * block * 4;   // upper bound per channel block
  float* wp = nullptr;
  TRS_CUDA(scratch_alloc(reinterpret_cast<void**>(&wp), w_floats * sizeof(float) * passes, s));
  int rc = TRS_OK;
  for (int pass = 0; pass < passes && rc == TRS_OK; ++pass) {
    const int c0 = pass * block;
    const int c_cnt = c_dim - c0 < block ? c_dim - c0 : block;
    const int npad = round_up(c_cnt, 32);
    const int64_t w_items = (int64_t)(kp / 16) * 2 * 4 * npad * 4;
    cin_tc_prep_weights_kernel<<<grid_for(w_items / 2, 256, 8), 256, 0, s>>>(w, c0, c_cnt, 1, k_dim, kp, npad, 0,
                                                                             wp + (size_t)pass * w_floats);
    rc = check_launch("cin_tc_prep_weights_kernel");
  }
  if (rc == TRS_OK) {
    CinTcArgs a{};
    a.xt = nullptr; a.h = x; a.wp = wp; a.scale = nullptr; a.shift = bias;
    a.h_next = out; a.pooled = nullptr;
    a.m_rows = rows; a.fields = 1; a.embed = 1; a.hp0 = 0; a.hp = kp; a.h_pitch = k_dim; a.k_valid = k_dim;
    a.c_begin = 0; a.c_eff = c_dim < block ? c_dim : block; a.npad = round_up(a.c_eff, 32);
    a.c_total = c_dim; a.c_block = block; a.wp_pass_stride = (int64_t)w_floats;
    a.n_direct = 0; a.hid_begin = 0; a.hid_count = c_dim; a.hp_next = c_dim;
    a.pool_off = 0; a.pooled_width = 0; a.act = activation;
    const size_t a_stage = 2 * 4 * kTileM * 16, b_stage = (size_t)2 * 4 * a.npad * 16;
    const size_t fixed = (a_tmem ? 0 : kAStages * a_stage) + (size_t)kTileM * 4 + 2 * a.npad * 4 +
                         (2 * kMaxAStages + 2 * kMaxBStages + 2) * 8 + 16 + 128;
    int b_stages = kMaxBStages;
    while (b_stages > 2 && b_stages * b_stage + fixed > (size_t)kMaxDynSmem) --b_stages;
    a.b_stages = b_stages;
    const size_t smem = b_stages * b_stage + fixed;
    const int64_t tiles = (rows + kTileM - 1) / kTileM;
    const int per_pass = kNumSMs / passes > 0 ? kNumSMs / passes : 1;
    const dim3 grid(static_cast<unsigned>(tiles < per_pass ? tiles : per_pass), passes);
    if (a_tmem) {
      TRS_SMEM_OPT_IN(cin_tc_layer_kernel<true>);
      cin_tc_layer_kernel<true><<<grid, kThreads, smem, s>>>(a);
    } else {
      TRS_SMEM_OPT_IN(cin_tc_layer_kernel<false>);
      cin_tc_layer_kernel<false><<<grid, kThreads, smem, s>>>(a);
    }
    rc = check_launch("cin_tc_layer_kernel(dense)");
  }
  cudaFreeAsync(wp, s);
  return rc;
}

}  // namespace trs
