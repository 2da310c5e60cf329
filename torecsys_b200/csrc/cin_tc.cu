// a8 on the 5th-generation tensor cores: Compress Interaction Network with tcgen05.mma (kind::tf32), accumulators in
// TMEM, FP32-accurate through the 3xTF32 split (A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate).
//
// GEMM view per layer: D[(b,e), c] = sum_{x,y} x0[(b,e), x] * h[(b,e), y] * W[c, x*H + y].
//   * rows (b,e) are independent, so activations are kept ROW-major between layers: h_l [(b,e)][Hp_l] -- which is exactly
//     the accumulator layout (TMEM lane = row, column = channel); the layer input x (B,N,E) is read in place (x0 of
//     every layer, h of layer 0: the E lanes of a sample read consecutive floats of one field's row);
//   * the K index is re-ordered y-chunk-major: k' = (yc*N + x)*16 + (y%16).  A producer thread owns one row, keeps
//     the 16 h values of the current y-chunk in registers across the N chunks that share it, and per chunk emits
//     z = x0[x] * h[y] split into TF32 hi/lo straight into the UMMA canonical (K-major, no swizzle) smem layout;
//     the A operand never exists in HBM.  The weights are pre-split and pre-permuted once per call into the same
//     chunked layout (cin_tc_prep_weights) and stream from L2 with one cp.async.bulk per chunk;
//   * CTA tile = 256 rows x Npad channels (two M=128 accumulators = up to 512 TMEM columns), BK = 16 per stage;
//     per chunk 2 halves x 2 k-steps x 3 split terms = 12 tcgen05.mma; tcgen05.commit frees the stage / signals
//     the epilogue; roles: warps 0-7 = A producers then epilogue (tcgen05.ld -> folded Conv-bias + eval-BN + act ->
//     hidden half stored row-major through a shared-memory transpose (whole lines per store), direct half summed over
//     e by a butterfly reduce-scatter -> pooled), warp 8 = MMA issuer, warp 9 = weight loader, warps 10-11 idle: they
//     and warps 8-9 hand their registers to the producers (setmaxnreg);
//   * the same kernel is the DENSE layer of wide MLPs (one field, x0 = 1; kFused instantiations: channel blocks over
//     blockIdx.y, rows gathered from an embedding table, FM / first-order row base, logit Linear in the epilogue);
//   * TRS_CIN_TRACE / TRS_DENSE_TRACE: per-role cycle counters (what each role waits for), see trace_report();
//   * the dead hidden half of the last layer (computed and discarded by the reference) is not computed.
// Shapes outside (E in {8,16,32}, channels <= 256 per layer) use the FFMA path in cin.cu.
#include <stdlib.h>

#include "tc5.cuh"
#include "tile_ops.cuh"

namespace trs {
namespace {

constexpr int kTileM = 256;          // rows per CTA tile (2 x UMMA_M 128)
// (K per stage = 16 = 4 sixteen-byte chunks = 2 UMMA k-steps)
constexpr int kProducerThreads = 256;
constexpr int kThreads = kProducerThreads + 128;  // + MMA warp + weight-loader warp + two idle warps (register donors)
constexpr int kAStages = 2;          // operand-A ring in smem (32 KB per stage; generation is cheap, two stages suffice)
constexpr int kMaxAStages = 4;       // (four 64-column stages when the A ring lives in tensor memory)
constexpr int kMaxBStages = 8;       // weight ring: deep, the L2 -> smem stream is latency-bound

using namespace tc5;

// ---- preparation kernels ------------------------------------------------------------------------------------------------
// W (C, N*H) -> wp[q = yc*N + x][hi|lo][kc 0..3][n 0..npad-1][4 floats]  with k' = q*16 + kc*4 + j,  y = yc*16 + kc*4 + j
// fold != 0 (layer 0, where h = x0 and z[x, y] = z[y, x]): the pair (x, y) is kept only for y >= x, with the weight
// W[c, x, y] + W[c, y, x] (W[c, x, x] on the diagonal); the kernel then skips the chunks that lie below the diagonal.
__global__ void __launch_bounds__(256) cin_tc_prep_weights_kernel(const float* __restrict__ w, int c_begin, int c_eff, int fields,
                                                                  int h_prev, int hp, int npad, int fold,
                                                                  float* __restrict__ wp, int c_real = 1 << 30,
                                                                  int sel_embed = 1, int c_block = 0, int c_total = 0,
                                                                  int64_t pass_stride = 0) {
  if (c_block > 0) {   // dense layer: one launch prepares every channel block (blockIdx.y), 16-column granularity
    c_begin = blockIdx.y * c_block;
    c_eff = c_total - c_begin < c_block ? c_total - c_begin : c_block;
    npad = (c_eff + 15) & ~15;
    wp += blockIdx.y * pass_stride;
  }
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
    if (n < c_eff && y < h_prev && c_begin + n >= c_real) {
      v = (y % sel_embed == c_begin + n - c_real) ? 1.f : 0.f;   // selector channel (dense mode: fields == 1, k = y)
    } else if (n < c_eff && y < h_prev) {
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
  const float* xb;       // CIN: the layer input x (B, N, E) as the lookup wrote it -- the x0 operand of every layer and,
                         // with h == null, layer 0's activations; rows (b, e) read it with stride E (the 16 lanes of a
                         // sample make one 64-byte request per field).  Null = plain dense layer (x0 = 1)
  const float* h;        // [(b,e)][hp]   input activations of this layer (null: layer 0 of a CIN, see xb)
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
  // ---- dense mode, fused ends (wide deep branch of DeepFM) ---------------------------------------------
  // gather: row m of the input is sample m's embedding rows, read straight from the table by the A producers
  // (k = field * g_embed + e), so the (B, fields * embed) matrix never exists in HBM
  const void* g_idx;        // (m_rows, g_fields) indices, null = plain dense input `h`
  const int64_t* g_offsets;
  const float* g_table;     // (g_rows, g_embed)
  const float* g_wfeat;     // (g_rows,) first-order table or null
  const float* g_bias;      // scalar added to the row base or null
  int32_t* g_status;
  float* row_base;          // (m_rows,) <- first-order + FM (+ bias): written by the CTA that owns the selector channels
  int64_t g_rows;
  int g_idx_bits, g_fields, g_embed;
  int g_pitch, g_wcol;      // floats between table rows; >= 0: first-order value in that column of the row's line
  // selector channels [c_real, c_real + sel_count): weight 1 where k % g_embed == channel - c_real, so their accumulators
  // are the FM field sums s[e] = sum_f v[f, e]; the epilogue turns them into 0.5 * (sum_e s[e]^2 - sum v^2)
  int c_real, sel_count;
  // dot: the layer feeds a one-output Linear -- instead of storing its activations the epilogue writes
  // dot_out[blockIdx.y][m] = sum_{c in this block} act[m][c] * dot_w[c]
  const float* dot_w;
  float* dot_out;
  // TRS_DENSE_TRACE=1: per CTA cycle counters of the dense instantiations, 8 per CTA (see dense_tc_run)
  long long* trace;
  int coop;   // dense, A operand in shared memory: cooperative producers (always with gather)
  int a_stages;   // dense, A operand in shared memory: depth of the A ring (2..kMaxAStages)
};

// kATmem = true (npad <= 128): the generated A operand goes to TENSOR MEMORY (columns 256..511, four 64-column stages:
//   per half hi[16] | lo[16]) with tcgen05.st, the accumulators use columns 0..255, and only the weights stream from
//   shared memory -- the SS form with both operands in shared memory is smem-port bound for N = 128
//   (profiles/r01_cin_tcgen05_notes.md).  kATmem = false (npad = 256): accumulators need all 512 columns, A in smem.
//
// kFused = true: the dense-mode instantiation: channel blocks in multiples of 16 and the fused ends described in CinTcArgs
//   (rows gathered from the embedding table by the producers, FM / first-order row base, one-output Linear folded into the
//   epilogue).
// Sum of 32 channel values over the E consecutive lanes (rows) of a sample as a butterfly REDUCE-SCATTER: at every step
// a lane hands half of its channels to its partner and sums the partner's copies of the half it keeps -- 30 shuffles
// (E = 16) instead of the 128 of an all-reduce per channel, the same pairing and order of additions (bit-identical sums);
// the lane ends up with the 32 / E channels starting at `chan` and writes them.
template <int E>
__device__ __forceinline__ void pool_rows(float (&w)[32], int lane, bool row_ok, float* prow, int lim) {
  int chan = 0;
#pragma unroll
  for (int o = E / 2, half = 16; o >= 1; o >>= 1, half >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = up ? w[j] : w[j + half];
      const float keep = up ? w[j + half] : w[j];
      w[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
    if (up) chan += half;
  }
  constexpr int kLeft = 32 / E;
#pragma unroll
  for (int j = 0; j < kLeft; ++j)
    if (row_ok && chan + j < lim) prow[chan + j] = w[j];
}

__host__ __device__ inline int ss_pitch(int npad) { return (npad + 31) & ~31; }   // floats per per-channel array in smem

// The kernel runs 12 warps: ptxas budgets 168 registers per thread, the two idle warps, the MMA warp and the weight loader
// hand theirs back (setmaxnreg) and the eight producer / epilogue warps grow to 232 -- the cooperative gather keeps
// 5 chunks x 16 loaded values plus the next chunk's indices in flight per thread, and a SPILLED in-flight load stalls the
// warp until the load lands (measured 6 400 cycles per chunk with 300 bytes of spills).
constexpr int kDenseThreads = kThreads;
template <bool kATmem, bool kFused, bool kTrace = false, int kGB = 0>   // kGB: gather with 32- / 64-bit indices (0 = none)
__global__ void __launch_bounds__(kThreads, 1) cin_tc_layer_kernel(CinTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (gridDim.y > 1) {   // wide dense layer: this CTA column owns one channel block
    a.c_begin = blockIdx.y * a.c_block;
    a.c_eff = a.c_total - a.c_begin < a.c_block ? a.c_total - a.c_begin : a.c_block;
    a.npad = (a.c_eff + 15) & ~15;
    a.wp += blockIdx.y * a.wp_pass_stride;
  }
  const int kAS = kATmem ? 4 : (kFused ? a.a_stages : kAStages);   // A ring depth
  const int npad = a.npad;
  const int dstride = kATmem ? 128 : npad;                     // TMEM columns between the two accumulator halves
  const uint32_t a_stage_bytes = kATmem ? 0 : 2 * 4 * kTileM * 16;   // hi/lo x 4 chunks x 256 rows x 16 B = 32 KB
  const uint32_t b_stage_bytes = 2 * 4 * npad * 16;            // hi/lo x 4 chunks x npad rows x 16 B
  unsigned char* a_smem = smem_raw;
  unsigned char* b_smem = a_smem + (size_t)kAS * a_stage_bytes;
  float* x0_s = reinterpret_cast<float*>(b_smem + (size_t)a.b_stages * b_stage_bytes);   // [fields][256]
  float* ss_s = x0_s + (size_t)a.fields * kTileM;                    // scale, shift, dot weight, selector flag: [np32] each
  const int np32 = ss_pitch(npad);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ss_s + 4 * np32);
  // barriers: full_a[kMaxAStages], empty_a[kMaxAStages], full_b[kMaxBStages], empty_b[kMaxBStages], acc_full, acc_empty
  constexpr int kNumBars = 2 * kMaxAStages + 2 * kMaxBStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kNumBars);
  // (dense) per-warp transpose buffers of the epilogue, 8 x 32 rows x 36 floats: over the A ring, which is idle once the
  // tile's accumulators are complete, or -- A operand in tensor memory -- in their own space behind the barriers
  float* stage_s = kATmem ? reinterpret_cast<float*>(tmem_slot + 4) : reinterpret_cast<float*>(a_smem);
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
  for (int i = threadIdx.x; i < np32; i += blockDim.x) {
    const int gc = a.c_begin + i;
    const bool real = i < a.c_eff && gc < a.c_real;      // (selector channels: scale 1, no bias)
    ss_s[i] = i < a.c_eff ? ((real && a.scale) ? __ldg(a.scale + gc) : 1.f) : 0.f;
    ss_s[np32 + i] = (real && a.shift) ? __ldg(a.shift + gc) : 0.f;
    ss_s[2 * np32 + i] = (kFused && real && a.dot_w) ? __ldg(a.dot_w + gc) : 0.f;
    ss_s[3 * np32 + i] = (kFused && i < a.c_eff && gc >= a.c_real) ? 1.f : 0.f;
  }
  // the CTA column that writes the per-row base (first-order + FM): the one whose channel block holds the selectors
  const bool do_base = kFused && a.row_base != nullptr &&
                       (a.sel_count > 0 ? (a.c_begin <= a.c_real && a.c_real + a.sel_count <= a.c_begin + a.c_eff)
                                        : blockIdx.y == 0);
  // (gather) the field offsets live in shared memory, in the x0 array a dense layer does not use: a global load of a
  // warp-uniform value gets an R2UR right behind it, which stalls the warp for the load's latency once per chunk
  int64_t* off_s = reinterpret_cast<int64_t*>(x0_s);
  if (kGB != 0)
    for (int i = threadIdx.x; i < a.g_fields; i += blockDim.x) off_s[i] = __ldg(a.g_offsets + i);
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

  if (warp < 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp < 8) {
    // =========================== A producers, then epilogue =====================================================
    const int r = threadIdx.x;   // row within the tile
    int sa = 0;                  // A ring position and phase
    uint32_t pa = 0;
    uint32_t tile_n = 0;
    long long tr_wait_a = 0, tr_wait_acc = 0, tr_epi = 0;   // (trace) producer warp 0: cycles blocked / in the epilogue
    const long long tr_t0 = (kTrace) ? clock64() : 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
      const int64_t m = tile * kTileM + r;
      const bool row_ok = m < a.m_rows;
      const int64_t xb_b = (a.xb != nullptr && row_ok) ? m / a.embed : 0;           // sample / embedding column of this row
      const float* xrow = a.xb != nullptr ? a.xb + xb_b * a.fields * a.embed + (row_ok ? m - xb_b * a.embed : 0) : nullptr;
      if (!(kFused && !kATmem && (kGB != 0 || a.coop)))   // (the cooperative dense producers have no x0; gather keeps offsets there)
      for (int xf = 0; xf < a.fields; ++xf)
        x0_s[xf * kTileM + r] = row_ok ? (xrow != nullptr ? __ldg(xrow + xf * a.embed) : 1.f) : 0.f;   // dense: x0 = 1
      // kPF y-chunks of this row's h are in flight in registers: with few fields (one, for a plain dense layer) a chunk
      // lasts less than a global-load latency, so the loads run that many chunks ahead of their use
      constexpr int kPF = kATmem ? 3 : 4;   // (register budget: the TMEM form also holds hi[16] / lo[16])
      float hbuf[kPF][16];
      float first = 0.f, sq = 0.f;        // (fused gather) this row's sum of first-order values / of v^2
      if (kFused && !kATmem && (kGB != 0 || a.coop)) {
        // ---- dense layer, A operand in shared memory: the warp produces its 32 rows COOPERATIVELY -- lane (g, kq) =
        // (lane / 4, lane % 4) handles the 16-byte piece kq of rows 8 i + g (i = 0..3), so the four lanes of a group read
        // the 64 bytes of one row with ONE request (a thread per row issues four 16-byte requests per row: a random
        // gather is bound by the request rate, ~35 G/s -- measured 745 us for the gathered layer that way), and every
        // piece goes to the same place in the UMMA layout as before.  Gather (g_idx != null): row m = sample m, chunk yc
        // = 16 columns of field f = 16 yc / embed (embed % 16 == 0); the lane also owns the first-order lookup of row
        // 8 kq + g and reports that row's out-of-range indices.
        const int g = lane >> 2, kq = lane & 3;
        constexpr bool kGather = kGB != 0;

        const int64_t row0 = tile * kTileM + warp * 32 + g;   // rows of this lane: row0 + 8 i
        bool ok[4];
        const unsigned char* rp[4];   // gather: address of the row's first index (of row 0 for rows beyond the batch:
                                      // no branch around the load); else: address of the row's piece kq of chunk 0
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ok[i] = row0 + 8 * i < a.m_rows;
          if (kGather)
            rp[i] = static_cast<const unsigned char*>(a.g_idx) + (ok[i] ? (row0 + 8 * i) * a.g_fields : 0) * (kGB / 8);
          else
            rp[i] = reinterpret_cast<const unsigned char*>(a.h + (ok[i] ? (row0 + 8 * i) : 0) * a.h_pitch + 4 * kq);
        }
        constexpr int kPFc = 5;   // chunks in flight per thread: random DRAM rows under load take several chunk times
        float hbc[kPFc][16];
        float hw[kPFc];
        float sq4[4] = {0.f, 0.f, 0.f, 0.f};
        // The producers are bound by their INSTRUCTION count (two warps per scheduler): the walk over the chunks keeps
        // its field / column counters incrementally (no division), addresses are one IMAD.WIDE each, the range check one
        // unsigned compare.
        // Gather: the raw indices (and the field offset) of a chunk are fetched one chunk ahead of the row loads that need
        // them and NOTHING touches them before that: a warp issues in order, so any instruction that consumes a load (an
        // offset add, a sign extension, a range check) right behind it stalls the warp for a full memory latency -- five
        // such stalls per chunk cost 6 400 cycles per chunk where the tile's MMAs need 1 250.
        int cy = yc_rot;                                           // chunk the next fetch / plain load is for
        int cf = kGather ? (yc_rot * 16) / a.g_embed : 0;          // ... its field
        int cw = kGather ? yc_rot * 16 - cf * a.g_embed : 0;       // ... its first column within the field's row
        int lf = 0, lw = 0;                                        // field / column of the chunk whose indices are held
        const uint32_t row_bytes = kGather ? static_cast<uint32_t>(a.g_pitch) * 4u : 0u;
        uint32_t rlo[4] = {0u, 0u, 0u, 0u}, rhi[4] = {0u, 0u, 0u, 0u};
        int64_t offn = 0;
        auto advance = [&]() {
          cw += 16;
          if (kGather && cw == a.g_embed) { cw = 0; ++cf; }
          if (++cy == ychunks) { cy = 0; cf = 0; cw = 0; }
        };
        auto fetch_idx = [&](bool live) {   // live (warp-uniform): the chunk exists
          lf = live ? cf : 0;
          lw = cw;
          offn = off_s[lf];
          const uint32_t fo = static_cast<uint32_t>(lf) * (kGB / 8);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (kGB == 64)
              asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(rlo[i]), "=r"(rhi[i]) : "l"(rp[i] + fo));
            else
              asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(rlo[i]) : "l"(rp[i] + fo));
          }
          advance();
        };
        if (kGather) fetch_idx(true);
        auto load_c = [&](float (&dst)[16], float& wv, int yci) {
          const bool live = yci < ychunks;
          wv = 0.f;
          if (kGather) {
            const unsigned char* tb = reinterpret_cast<const unsigned char*>(a.g_table + lw + 4 * kq);
            const bool starts = lw == 0;
            // (ONE first-order load per chunk, issued behind the loop: four predicated loads into the same register
            // wait for each other -- write-after-write on the scoreboard, a memory latency each)
            uint32_t rr_own = 0u;
            bool own_live = false, own_in = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              const int64_t rr = (kGB == 64 ? static_cast<int64_t>((static_cast<uint64_t>(rhi[i]) << 32) | rlo[i])
                                            : static_cast<int64_t>(static_cast<int32_t>(rlo[i]))) + offn;
              const bool inr = static_cast<uint64_t>(rr) < static_cast<uint64_t>(a.g_rows);   // 0 <= rr < rows
              const bool lv = ok[i] && live;
              if (lv && inr)
                v = ldg_stream_f4(reinterpret_cast<const float4*>(
                    tb + static_cast<uint64_t>(static_cast<uint32_t>(rr)) * row_bytes));
              if (i == kq) { rr_own = static_cast<uint32_t>(rr); own_live = lv; own_in = inr; }
              dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
            }
            if (do_base && starts && own_live) {
              if (own_in) {
                if (a.g_wfeat != nullptr) wv = ldg_stream_f1(a.g_wfeat + rr_own);
                else if (a.g_wcol >= 0)   // packed table: the value lies in the line the row load is fetching anyway
                  wv = ldg_stream_f1(a.g_table + static_cast<uint64_t>(rr_own) * a.g_pitch + a.g_wcol);
              } else {
                report_oob(a.g_status, (row0 + 8 * kq) * (int64_t)a.g_fields + lf);
              }
            }
            fetch_idx(yci + 1 < ychunks);
            return;
          }
          const bool kok = live && cy * 16 + 4 * kq < a.k_valid;
          const uint32_t ko = static_cast<uint32_t>(cy) * 64u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[i] && kok) v = __ldg(reinterpret_cast<const float4*>(rp[i] + ko));
            dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
          }
          advance();
        };
#pragma unroll
        for (int p = 0; p < kPFc; ++p) load_c(hbc[p], hw[p], p);
        for (int y0 = 0; y0 < ychunks; y0 += kPFc) {
#pragma unroll
          for (int p = 0; p < kPFc; ++p) {
            if (y0 + p >= ychunks) break;
            float hreg[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) hreg[j] = hbc[p][j];
            if (do_base) {
              first += hw[p];
#pragma unroll
              for (int j = 0; j < 16; ++j) sq4[j >> 2] = fmaf(hreg[j], hreg[j], sq4[j >> 2]);
            }
            load_c(hbc[p], hw[p], y0 + p + kPFc);
            const long long tw0 = (kTrace && threadIdx.x == 0) ? clock64() : 0;
            mbar_wait(empty_a(sa), pa ^ 1);
            if (kTrace && threadIdx.x == 0) tr_wait_a += clock64() - tw0;
            unsigned char* dst = a_smem + (size_t)sa * a_stage_bytes + kq * (kTileM * 16) + (warp * 32 + g) * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float z = hreg[4 * i + j];
                hi[j] = (__float_as_uint(z) + 0x1000u) & 0xffffe000u;
                lo[j] = (__float_as_uint(z - __uint_as_float(hi[j])) + 0x1000u) & 0xffffe000u;
              }
              *reinterpret_cast<uint4*>(dst + 8 * i * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dst + 4 * (kTileM * 16) + 8 * i * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(sa));
            if (++sa == kAS) { sa = 0; pa ^= 1; }
          }
        }
        if (do_base) {
          // hand every row's sums to the thread that runs its epilogue (lane = row within the warp = 8 i + g)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            sq4[i] += __shfl_xor_sync(0xffffffffu, sq4[i], 1);
            sq4[i] += __shfl_xor_sync(0xffffffffu, sq4[i], 2);
          }
          first = __shfl_sync(0xffffffffu, first, 4 * (lane & 7) + (lane >> 3));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float t = __shfl_sync(0xffffffffu, sq4[i], 4 * (lane & 7));
            if ((lane >> 3) == i) sq = t;
          }
        }
      } else {
      auto load_h = [&](float (&dst)[16], int yci) {
        const int yc = yci + yc_rot < ychunks ? yci + yc_rot : yci + yc_rot - ychunks;
        if (!kFused && a.h == nullptr) {   // layer 0 of a CIN: h = x0, straight from x (B, N, E)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int y = yc * 16 + j;
            dst[j] = (row_ok && yci < ychunks && y < a.k_valid) ? __ldg(xrow + y * a.embed) : 0.f;
          }
          return;
        }
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
          const long long tw0 = (kTrace && threadIdx.x == 0) ? clock64() : 0;
          mbar_wait(empty_a(sa), pa ^ 1);   // stage free (first pass: passes immediately)
          if (kTrace && threadIdx.x == 0) tr_wait_a += clock64() - tw0;
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
      }   // (thread-per-row producers)
      // ---- epilogue of this tile: warps 0-3 -> accumulator half 0 (rows 0..127), warps 4-7 -> half 1 ------------
      const long long te0 = (kTrace && threadIdx.x == 0) ? clock64() : 0;
      mbar_wait(acc_full, tile_n & 1);
      tc_fence_after();
      const long long te1 = (kTrace && threadIdx.x == 0) ? clock64() : 0;
      const int half = warp >> 2;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + half * dstride;
      const int64_t b = row_ok ? m / a.embed : 0;
      float dot = 0.f, ssq = 0.f;   // (fused ends) this row's share of the one-output Linear / sum_e s[e]^2
      for (int c0 = 0; c0 < npad; c0 += 32) {
        uint32_t raw[32];
        if (!kFused || c0 + 32 <= npad) {
          tmem_ld32(taddr + c0, raw);
        } else {   // 16-column tail of a dense channel block
          uint32_t r16[16];
          tmem_ld16(taddr + c0, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { raw[j] = r16[j]; raw[16 + j] = 0u; }
        }
        float v[32];
        if (kFused) {
          // dense layer: no per-channel scale, the bias comes as 128-bit broadcast loads, ReLU without the switch (the
          // epilogue runs with two warps per scheduler, every dependent shared-memory load is exposed)
          const float4* sh4 = reinterpret_cast<const float4*>(ss_s + np32 + c0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = sh4[j4];
            v[4 * j4 + 0] = __uint_as_float(raw[4 * j4 + 0]) + b4.x;
            v[4 * j4 + 1] = __uint_as_float(raw[4 * j4 + 1]) + b4.y;
            v[4 * j4 + 2] = __uint_as_float(raw[4 * j4 + 2]) + b4.z;
            v[4 * j4 + 3] = __uint_as_float(raw[4 * j4 + 3]) + b4.w;
          }
          if (a.act == TRS_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (a.act != TRS_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], a.act);
          }
          if (a.dot_out != nullptr) {
            const float4* dw4 = reinterpret_cast<const float4*>(ss_s + 2 * np32 + c0);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 w4 = dw4[j4];
              dot = fmaf(v[4 * j4 + 0], w4.x, dot);
              dot = fmaf(v[4 * j4 + 1], w4.y, dot);
              dot = fmaf(v[4 * j4 + 2], w4.z, dot);
              dot = fmaf(v[4 * j4 + 3], w4.w, dot);
            }
          }
        } else {
          // CIN layer: folded Conv-bias + eval-BN as scale / shift (128-bit broadcast loads), ReLU without the switch
          const float4* sc4 = reinterpret_cast<const float4*>(ss_s + c0);
          const float4* sh4 = reinterpret_cast<const float4*>(ss_s + np32 + c0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 s4 = sc4[j4], b4 = sh4[j4];
            v[4 * j4 + 0] = fmaf(__uint_as_float(raw[4 * j4 + 0]), s4.x, b4.x);
            v[4 * j4 + 1] = fmaf(__uint_as_float(raw[4 * j4 + 1]), s4.y, b4.y);
            v[4 * j4 + 2] = fmaf(__uint_as_float(raw[4 * j4 + 2]), s4.z, b4.z);
            v[4 * j4 + 3] = fmaf(__uint_as_float(raw[4 * j4 + 3]), s4.w, b4.w);
          }
          if (a.act == TRS_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (a.act != TRS_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], a.act);
          }
        }
        if (kFused && do_base && a.sel_count > 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float rv = __uint_as_float(raw[j]);
            ssq = fmaf(rv * ss_s[3 * np32 + c0 + j], rv, ssq);
          }
        }
        // hidden half -> next layer's activations, row-major
        if (a.h_next != nullptr && a.c_begin + c0 + 32 > a.hid_begin && a.c_begin + c0 < a.hid_begin + a.hid_count) {
          // the warp's 32 rows x 32 columns go through a shared-memory transpose so that every store instruction
          // writes four whole 128-byte lines (a thread per row writes 16 bytes into each of 32 lines: the epilogue
          // of a 400-wide dense layer took 44 K cycles per tile that way, more than the tile's MMAs)
          float* st = stage_s + warp * (32 * 36);
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(st + lane * 36 + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
          __syncwarp();
          const int cc = lane & 7;
          const int c = a.c_begin + c0 + 4 * cc;              // global channel of this lane's four columns
          const int cl = c0 + 4 * cc;                         // ... within the block
          const bool vec_ok = (a.hp_next & 3) == 0 && ((c - a.hid_begin) & 3) == 0;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int jr = it * 4 + (lane >> 3);
            const int64_t mr = tile * kTileM + warp * 32 + jr;
            if (mr >= a.m_rows) continue;
            const float4 o = *reinterpret_cast<const float4*>(st + jr * 36 + 4 * cc);
            float* orow = a.h_next + mr * a.hp_next - a.hid_begin;
            if (vec_ok && c >= a.hid_begin && c + 3 < a.hid_begin + a.hid_count && cl + 3 < a.c_eff) {
              *reinterpret_cast<float4*>(orow + c) = o;
            } else {
              const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (c + j >= a.hid_begin && c + j < a.hid_begin + a.hid_count && cl + j < a.c_eff) orow[c + j] = ov[j];
            }
          }
        }
        // direct half -> sum over the `embed` rows of each sample (consecutive lanes), one writer per (b, c)
        if (!kFused && a.c_begin + c0 < a.n_direct) {
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          float* prow = a.pooled + b * a.pooled_width + a.pool_off + a.c_begin + c0;
          const int lim = min(a.n_direct - a.c_begin - c0, a.c_eff - c0);   // channels of this group that exist
          if (a.embed == 16) pool_rows<16>(v, lane, row_ok, prow, lim);
          else if (a.embed == 32) pool_rows<32>(v, lane, row_ok, prow, lim);
          else pool_rows<8>(v, lane, row_ok, prow, lim);
        }
      }
      if (kFused && row_ok) {
        if (a.dot_out != nullptr) a.dot_out[(int64_t)blockIdx.y * a.m_rows + m] = dot;
        if (do_base)
          a.row_base[m] = first + (a.sel_count > 0 ? 0.5f * (ssq - sq) : 0.f) + (a.g_bias ? __ldg(a.g_bias) : 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      // the transpose buffers lie over the A ring: no warp may start producing the next tile before all have stored
      if (!kATmem) asm volatile("bar.sync 1, 256;" ::: "memory");
      if (kTrace && threadIdx.x == 0) {
        tr_wait_acc += te1 - te0;
        tr_epi += clock64() - te1;
      }
    }
    if (kTrace && threadIdx.x == 0) {
      long long* t = a.trace + 8 * (blockIdx.y * gridDim.x + blockIdx.x);
      t[0] = clock64() - tr_t0; t[1] = tr_wait_a; t[2] = tr_wait_acc; t[3] = tr_epi;
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
    long long tr_fa = 0, tr_fb = 0, tr_ae = 0;   // (trace) cycles the MMA warp waits for A, for B, for the epilogue
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_n) {
      const long long tm0 = (kTrace) ? clock64() : 0;
      mbar_wait(acc_empty, (tile_n & 1) ^ 1);   // epilogue of the previous tile has drained TMEM
      tc_fence_after();
      if (kTrace) tr_ae += clock64() - tm0;
      for (int q = 0; q < chunks; ++q) {
        const long long tm1 = (kTrace) ? clock64() : 0;
        mbar_wait(full_a(sa), pa);
        const long long tm2 = (kTrace) ? clock64() : 0;
        mbar_wait(full_b(sb), pb);
        tc_fence_after();
        if (kTrace) { tr_fa += tm2 - tm1; tr_fb += clock64() - tm2; }
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
    if (kTrace && lane == 0) {
      long long* t = a.trace + 8 * (blockIdx.y * gridDim.x + blockIdx.x);
      t[4] = tr_fa; t[5] = tr_fb; t[6] = tr_ae;
    }
  } else if (warp == 9) {
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

// TRS_DENSE_TRACE / TRS_CIN_TRACE: averages the per-CTA cycle counters of a traced launch and prints them
static void trace_report(long long* dev, int n, cudaStream_t s, const char* label) {
  long long* hbuf = static_cast<long long*>(malloc(sizeof(long long) * 8 * n));
  cudaStreamSynchronize(s);
  cudaMemcpy(hbuf, dev, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost);
  double sum[8] = {0};
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 8; ++j) sum[j] += (double)hbuf[8 * i + j] / n;
  fprintf(stderr,
          "%s: total %.0f cyc | producer: wait A-free %.0f, wait acc %.0f, epilogue %.0f | mma: wait A %.0f, wait B %.0f, "
          "wait epilogue %.0f\n", label, sum[0], sum[1], sum[2], sum[3], sum[4], sum[5], sum[6]);
  free(hbuf);
  cudaFree(dev);
}

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
  float* xt = reinterpret_cast<float*>(ws);   // (space of the former transposed copy of x: unused, kept in the layout)
  float* hbuf[2] = {xt + p.xt_floats, xt + p.xt_floats + p.h_floats};
  float* pooled = hbuf[1] + p.h_floats;
  float* wprep = pooled + p.pooled_floats;
  const int64_t m_rows = batch * embed;

  int rc = TRS_OK;   // (no transposed copy of x any more: the layer kernels read x (B, N, E) in place)
  TRS_SMEM_OPT_IN((cin_tc_layer_kernel<true, false>));
  TRS_SMEM_OPT_IN((cin_tc_layer_kernel<false, false>));

  const float* h = nullptr;   // layer 0 reads x itself
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
    a.xb = x; a.h = h; a.wp = wp; a.scale = scale[l]; a.shift = shift[l];
    a.h_next = last ? nullptr : hbuf[l & 1];
    a.pooled = pooled;
    a.m_rows = m_rows; a.fields = fields; a.embed = embed; a.hp0 = p.hp0; a.hp = hp; a.npad = npad; a.c_eff = c_eff;
    a.n_direct = hl;
    a.hid_begin = is_direct ? 0 : hl;
    a.hid_count = last ? 0 : hl;
    a.hp_next = round_up(hl, 16);
    a.pool_off = pool_off; a.pooled_width = p.pooled_width; a.act = activation;
    a.h_pitch = hp; a.k_valid = l == 0 ? fields : hp; a.fold = fold;
    a.c_real = 1 << 30;   // (no selector channels: every channel is a real one)
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
      const size_t fixed = (a_tmem ? (size_t)8 * 32 * 36 * 4 : kAStages * a_stage) + (size_t)fields * kTileM * 4 +
                           4 * ss_pitch(np) * 4 +
                           (2 * kMaxAStages + 2 * kMaxBStages + 2) * 8 + 16 + 128;
      int b_stages = kMaxBStages;
      while (b_stages > 2 && b_stages * b_stage + fixed > (size_t)kMaxDynSmem) --b_stages;
      static const int cap_b = getenv("TRS_CIN_B_STAGES") ? atoi(getenv("TRS_CIN_B_STAGES")) : 0;   // experiments
      if (cap_b >= 2 && cap_b < b_stages) b_stages = cap_b;
      const size_t smem = b_stages * b_stage + fixed;
      TRS_UNSUPPORTED(smem > (size_t)kMaxDynSmem, "cin: tensor-core tile does not fit shared memory");
      ap.b_stages = b_stages;
      if (getenv("TRS_CIN_VERBOSE")) fprintf(stderr, "cin layer %d pass %d: npad %d, a_tmem %d, b_stages %d, smem %zu\n", l, pass, np, (int)a_tmem, b_stages, smem);
      static const bool cin_trace = getenv("TRS_CIN_TRACE") != nullptr;
      if (cin_trace) {
        TRS_CUDA(cudaMalloc(reinterpret_cast<void**>(&ap.trace), sizeof(long long) * 8 * grid));
        TRS_CUDA(cudaMemsetAsync(ap.trace, 0, sizeof(long long) * 8 * grid, s));
        if (a_tmem) {
          TRS_SMEM_OPT_IN((cin_tc_layer_kernel<true, false, true>));
          cin_tc_layer_kernel<true, false, true><<<grid, kThreads, smem, s>>>(ap);
        } else {
          TRS_SMEM_OPT_IN((cin_tc_layer_kernel<false, false, true>));
          cin_tc_layer_kernel<false, false, true><<<grid, kThreads, smem, s>>>(ap);
        }
      } else if (a_tmem) cin_tc_layer_kernel<true, false><<<grid, kThreads, smem, s>>>(ap);
      else cin_tc_layer_kernel<false, false><<<grid, kThreads, smem, s>>>(ap);
      rc = check_launch("cin_tc_layer_kernel");
      if (rc != TRS_OK) return rc;
      if (cin_trace) {
        char label[160];
        snprintf(label, sizeof label, "cin trace layer %d pass %d npad %d %s chunks/tile fold %d tiles/cta %.2f", l, pass, np,
                 a_tmem ? "TS" : "SS", fold, (double)tiles / grid);
        trace_report(ap.trace, grid, s, label);
      }
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
// cut into equal channel blocks of at most 256 (multiples of 16; one block per blockIdx.y, the SMs divided between the
// blocks), each CTA streaming only its block of the weights.  Scratch for the pre-split weights comes from the
// stream-ordered pool.  TRS_DENSE_BLOCK=n caps the channel block (A/B measurements: <= 128 selects the TMEM-operand form).
//
// DenseFuse (tile_ops.cuh) adds the fused ends of the wide deep branch (multilayer_perceptron.py:63-84 fed by
// multi_indices_emb.py:92-112): `idx` != null -> the input rows are gathered from the embedding table by the producers and
// the per-row base first-order + FM (+ bias) is written to row_base; `dot_w` != null -> the layer's activations are not
// stored, dot_out[block][row] receives their product with the one-output Linear that follows.
int dense_tc_supported(int k_dim, int c_dim, const void* x, const void* out) {
  static const bool disabled = getenv("TRS_DISABLE_TC") != nullptr;
  return !disabled && k_dim % 4 == 0 && k_dim >= 16 && c_dim >= 16 && aligned16(x) && aligned16(out);
}

// channel blocks of a dense layer with c_real outputs (+ sel selector channels, kept inside ONE block); 0 = no split found
// Plain layers take blocks of at most 128 channels -- the TMEM-operand form, and 7 waves of 37 CTAs instead of 3.46 of
// 74 for 400 channels at 65 536 rows (measured 558 -> 518 us for the 624-400-400-400-1 chain); a gathering layer keeps two
// wide blocks, every block's CTA column gathers the rows again.
static int dense_plan(int c_real, int sel, int gather, int* block_out) {
  static const int cap_env = getenv("TRS_DENSE_BLOCK") ? atoi(getenv("TRS_DENSE_BLOCK")) : 0;
  const int cap = (cap_env >= 16 && cap_env <= 256) ? cap_env / 16 * 16 : (gather ? 256 : 128);
  const int c_total = c_real + sel;
  const int passes0 = (c_total + cap - 1) / cap;
  for (int block = round_up((c_total + passes0 - 1) / passes0, 16); block <= 256; block += 16) {
    if (sel > 0 && c_real / block != (c_real + sel - 1) / block) continue;   // selectors would straddle two blocks
    *block_out = block;
    return (c_total + block - 1) / block;
  }
  return 0;
}

int dense_tc_passes(int c_dim, int sel, int gather) {
  int block = 0;
  return dense_plan(c_dim, sel, gather, &block);
}

__global__ void __launch_bounds__(256) dense_dot_finish_kernel(const float* __restrict__ dot, int passes, int64_t rows,
                                                               const float* __restrict__ bias, float* __restrict__ out,
                                                               int accumulate) {
  const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (m >= rows) return;
  float v = accumulate ? out[m] : 0.f;
  for (int p = 0; p < passes; ++p) v += __ldg(dot + (int64_t)p * rows + m);   // fixed order: deterministic
  out[m] = v + (bias ? __ldg(bias) : 0.f);
}

// out[m] (+)= sum_p dot[p][m] + bias: closes a layer that ran with DenseFuse::dot_w
int dense_dot_finish(const float* dot, int passes, int64_t rows, const float* bias, float* out, int accumulate,
                     cudaStream_t s) {
  if (rows == 0) return TRS_OK;
  dense_dot_finish_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, s>>>(dot, passes, rows, bias, out,
                                                                                    accumulate);
  return check_launch("dense_dot_finish_kernel");
}

int dense_tc_run(const float* x, int64_t rows, int k_dim, const float* w, const float* bias, int c_dim, int activation,
                 float* out, cudaStream_t s, const DenseFuse* fz) {
  if (rows == 0) return TRS_OK;
  const bool gather = fz != nullptr && fz->idx != nullptr;
  const int sel = (gather && fz->use_fm) ? fz->embed : 0;
  if (gather)
    TRS_REQUIRE(fz->embed % 16 == 0 && fz->fields * fz->embed == k_dim && fz->fields <= kTileM / 2 &&
                    fz->table_rows <= (int64_t)1 << 32 && (fz->idx_bits == 32 || fz->idx_bits == 64) &&
                    (fz->row_pitch == 0 || (fz->row_pitch % 4 == 0 && fz->row_pitch >= fz->embed)),
                "dense: bad gather description");
  const int kp = round_up(k_dim, 16);
  int block = 0;
  const int passes = dense_plan(c_dim, sel, gather ? 1 : 0, &block);
  TRS_UNSUPPORTED(passes == 0, "dense: no channel split keeps the selector channels in one block");
  const int c_total = c_dim + sel;
  const bool a_tmem = block <= 128 && !gather;   // (the gathering producers write the shared-memory operand form)
  const size_t w_floats = (size_t)(kp / 16) * 2 * 4 * block * 4;   // per channel block
  float* wp = nullptr;
  TRS_CUDA(scratch_alloc(reinterpret_cast<void**>(&wp), w_floats * sizeof(float) * passes, s));
  int rc = TRS_OK;
  {
    const int64_t w_items = (int64_t)(kp / 16) * 2 * 4 * block * 4;   // of the widest block
    const dim3 pgrid(grid_for(w_items / 2, 256, 2), passes);
    cin_tc_prep_weights_kernel<<<pgrid, 256, 0, s>>>(w, 0, 0, 1, k_dim, kp, 0, 0, wp, c_dim, sel > 0 ? sel : 1, block,
                                                     c_total, (int64_t)w_floats);
    rc = check_launch("cin_tc_prep_weights_kernel");
  }
  if (rc == TRS_OK) {
    CinTcArgs a{};
    a.h = x; a.wp = wp; a.scale = nullptr; a.shift = bias;
    a.h_next = (fz != nullptr && fz->dot_w != nullptr) ? nullptr : out; a.pooled = nullptr;
    a.m_rows = rows; a.fields = 1; a.embed = 1; a.hp0 = 0; a.hp = kp; a.h_pitch = k_dim; a.k_valid = k_dim;
    a.c_begin = 0; a.c_eff = c_total < block ? c_total : block; a.npad = round_up(a.c_eff, 16);
    a.c_total = c_total; a.c_block = block; a.wp_pass_stride = (int64_t)w_floats;
    a.n_direct = 0; a.hid_begin = 0; a.hid_count = c_dim; a.hp_next = c_dim;
    a.pool_off = 0; a.pooled_width = 0; a.act = activation;
    a.c_real = c_dim; a.sel_count = sel;
    static const bool coop_env = getenv("TRS_DENSE_COOP") ? atoi(getenv("TRS_DENSE_COOP")) != 0 : false;
    a.coop = (gather || coop_env) ? 1 : 0;
    static const int a_stages_env = getenv("TRS_DENSE_A_STAGES") ? atoi(getenv("TRS_DENSE_A_STAGES")) : 0;
    // (the gathering producers run a stage further ahead of the MMAs: 798 -> 777 us for DeepFM-[400,400,400])
    a.a_stages = (a_stages_env >= 2 && a_stages_env <= kMaxAStages) ? a_stages_env : (gather ? 3 : kAStages);
    if (fz != nullptr) {
      a.g_idx = fz->idx; a.g_idx_bits = fz->idx_bits; a.g_offsets = fz->offsets; a.g_table = fz->table;
      a.g_wfeat = fz->w_feat; a.g_bias = fz->bias; a.g_status = fz->status; a.row_base = gather ? fz->row_base : nullptr;
      a.g_rows = fz->table_rows; a.g_fields = fz->fields; a.g_embed = fz->embed;
      a.g_pitch = fz->row_pitch > 0 ? fz->row_pitch : fz->embed; a.g_wcol = fz->w_col;
      a.dot_w = fz->dot_w; a.dot_out = fz->dot_w != nullptr ? fz->dot_out : nullptr;
    }
    const size_t a_stage = 2 * 4 * kTileM * 16, b_stage = (size_t)2 * 4 * a.npad * 16;
    const size_t fixed = (a_tmem ? (size_t)8 * 32 * 36 * 4 : a.a_stages * a_stage) + (size_t)kTileM * 4 +
                         4 * ss_pitch(a.npad) * 4 + (2 * kMaxAStages + 2 * kMaxBStages + 2) * 8 + 16 + 128;
    int b_stages = kMaxBStages;
    while (b_stages > 2 && b_stages * b_stage + fixed > (size_t)kMaxDynSmem) --b_stages;
    a.b_stages = b_stages;
    const size_t smem = b_stages * b_stage + fixed;
    const int64_t tiles = (rows + kTileM - 1) / kTileM;
    const int per_pass = kNumSMs / passes > 0 ? kNumSMs / passes : 1;
    const dim3 grid(static_cast<unsigned>(tiles < per_pass ? tiles : per_pass), passes);
    static const bool trace_on = getenv("TRS_DENSE_TRACE") != nullptr;
    if (trace_on) {
      TRS_CUDA(cudaMalloc(reinterpret_cast<void**>(&a.trace), sizeof(long long) * 8 * grid.x * grid.y));
      TRS_CUDA(cudaMemsetAsync(a.trace, 0, sizeof(long long) * 8 * grid.x * grid.y, s));
    }
    // every dense launch takes the kFused instantiation (16-column channel-block tails); <*, false> is the CIN's
#define TRS_DENSE_LAUNCH(...)                                                    \
  do {                                                                           \
    TRS_SMEM_OPT_IN((cin_tc_layer_kernel<__VA_ARGS__>));                         \
    cin_tc_layer_kernel<__VA_ARGS__><<<grid, kDenseThreads, smem, s>>>(a);       \
  } while (0)
    const int gb = gather ? fz->idx_bits : 0;
    if (trace_on) {
      if (a_tmem) TRS_DENSE_LAUNCH(true, true, true, 0);
      else if (gb == 64) TRS_DENSE_LAUNCH(false, true, true, 64);
      else if (gb == 32) TRS_DENSE_LAUNCH(false, true, true, 32);
      else TRS_DENSE_LAUNCH(false, true, true, 0);
    } else {
      if (a_tmem) TRS_DENSE_LAUNCH(true, true, false, 0);
      else if (gb == 64) TRS_DENSE_LAUNCH(false, true, false, 64);
      else if (gb == 32) TRS_DENSE_LAUNCH(false, true, false, 32);
      else TRS_DENSE_LAUNCH(false, true, false, 0);
    }
#undef TRS_DENSE_LAUNCH
    rc = check_launch("cin_tc_layer_kernel(dense)");
    if (trace_on && rc == TRS_OK) {
      char label[200];
      snprintf(label, sizeof label, "dense trace K %d C %d(+%d) block %d x%d %s gather %d dot %d tiles/cta %.2f", k_dim, c_dim,
               sel, block, passes, a_tmem ? "TS" : "SS", (int)gather, (int)(a.dot_out != nullptr), (double)tiles / grid.x);
      trace_report(a.trace, grid.x * grid.y, s, label);
    }
  }
  cudaFreeAsync(wp, s);
  return rc;
}

}  // namespace trs
