// DeepFM forward on a PACKED table (embed = 16, hidden widths = 16): the B200-first layout of the hot path.
//
// Measured on B200 (tools/gather_microbench.cu, profiles/r01_gather_microbench_ncu.csv): a random row read costs one
// 128-byte DRAM transaction whatever its size <= 128 B, and the chip sustains ~38 G such transactions/s.  The
// reference's two tables (emb (R,16) and first-order (R,1), same row ids) therefore cost 78 transactions per sample;
// packing them as one 128-byte-aligned row  [ v0..v15 | w | pad ]  (trs_fm_pack_table) makes it 39, the floor for
// this model.  The packed table is a shadow of the registered parameters (rebuilt by the host layer when their
// version changes), 25.6 GB for 200 M rows.
//
// Kernel shape (one persistent CTA of 8 warps per SM, grid = 148):
//   * a tile = 16 samples; warp w owns fields [w*FPW, (w+1)*FPW) of EVERY tile, so its slice of W1 (B fragments of
//     mma.sync.m16n8k8 TF32, pre-split hi/lo) lives in its registers -- no shared memory, no L2 traffic for W1;
//   * rows travel global -> shared with cp.async (LDGSTS) into a per-warp ring of kStages tiles, so in-flight loads
//     cost no registers (8 warps x 2 stages x 16 x FPW rows = 1 248 rows in flight per SM); eight lanes per row copy
//     [v|w] as ONE 80-byte request; the (16 x fields) index tiles have their own, deeper cp.async ring;
//     lane (g,t) then reads chunk t of the rows of samples g and g+8 = exactly its A-fragment elements;
//   * per tile each warp produces partial layer-1 accumulators, partial FM sums and first-order sums for its fields;
//     one bar.sync per tile, then warp (tile % 8) reduces the 8 partials and runs the tiny 16x16 layers + output
//     while the other warps already work on the next tile.
// FP32 accuracy on the tensor pipe through the 3xTF32 split; the roofline is the HBM transaction rate.
#include <stdlib.h>

#include "common.cuh"

namespace trs {
namespace {

constexpr int kWarps = 8;
constexpr int kTile = 16;
constexpr int kStages = 3;      // row ring depth (kStages-1 tiles of rows in flight)
constexpr int kIdxAhead = 2 * kStages - 1;  // index tiles are requested this many tiles ahead of their use ...
constexpr int kIdxSlots = kIdxAhead + 1;    // ... into a ring of this many slots
constexpr int kMaxHidden = 4;
constexpr int kRowFloats = 32;   // packed row pitch: 128 B
constexpr int kPartial = 18;     // floats per lane per warp per tile: 8 acc + 8 S + 2 c

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_rna(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// cheap split for streamed activations: hi = x rounded to 10 mantissa bits by integer add + mask (full-rate ALU
// instead of the quarter-rate cvt), lo = x - hi exactly; the tensor core ignores lo's low 13 bits (error 2^-21 |x|)
__device__ __forceinline__ void split_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                       uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al, bh0, bh1);
  mma_tf32(d, ah, bl0, bl1);
  mma_tf32(d, ah, bh0, bh1);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct PackedArgs {
  const void* idx;
  const int64_t* offsets;
  const float* packed;          // (rows, 32): v[16], w, pad
  const float* w1;              // (16, 16 N)
  const float* b1;
  const float* wh[kMaxHidden];  // (16, 16)
  const float* bh[kMaxHidden];
  const float* w_out;           // (1, 16)
  const float* b_out;
  float* logits;
  int32_t* status;
  int64_t batch, rows;
  int fields, hidden_layers;
};

// Row staging layout of one (warp, stage): [field f][sample pair k = s/2][40 floats]:
//   floats  0..15 = v of sample 2k, 16..31 = v of sample 2k+1, 32..35 / 36..39 = the 16-byte chunk holding w of each.
// The two rows of a pair fill the 32 banks exactly once, so the consumers' LDS.128 (quarter-warp = two samples x four
// chunks) are conflict-free, and every cp.async destination is 16-byte aligned.
constexpr int kPairFloats = 40;
constexpr int kFieldFloats = 8 * kPairFloats;   // 16 samples of one field

template <int FPW>
struct Smem {
  static constexpr size_t v_bytes = (size_t)kWarps * kStages * FPW * kFieldFloats * sizeof(float);
  static constexpr size_t w_bytes = 0;
  static constexpr size_t p_bytes = (size_t)2 * kWarps * kPartial * 32 * sizeof(float);
  static constexpr size_t h_bytes = (size_t)kMaxHidden * 8 * 32 * sizeof(float2);
  static constexpr size_t b_bytes = ((1 + kMaxHidden) * 16 + 16 + 4) * sizeof(float);
  static constexpr size_t fixed = v_bytes + w_bytes + p_bytes + h_bytes + b_bytes;
  // + index ring: kIdxSlots x (16 samples x fields x idx bytes), sized at launch
  static size_t total(int fields, int idx_bits) { return fixed + (size_t)kIdxSlots * kTile * fields * (idx_bits / 8); }
};

template <int IdxBits, int FPW>
__global__ void __launch_bounds__(kWarps * 32, 1) deepfm_packed_kernel(PackedArgs a) {
  using S = Smem<FPW>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* vbuf = reinterpret_cast<float*>(smem_raw);
  float* pbuf = reinterpret_cast<float*>(smem_raw + S::v_bytes + S::w_bytes);
  float2* whs = reinterpret_cast<float2*>(smem_raw + S::v_bytes + S::w_bytes + S::p_bytes);
  float* bias_s = reinterpret_cast<float*>(smem_raw + S::v_bytes + S::w_bytes + S::p_bytes + S::h_bytes);
  unsigned char* idx_ring = smem_raw + S::fixed;   // [kIdxSlots][16 * fields] indices, 16-byte aligned slots

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_fields = a.fields;
  const int kdim = 16 * n_fields;
  const int f0 = warp * FPW;  // first field of this warp

  // ---- one-time: this warp's W1 B-fragments into registers (hi/lo), hidden-layer fragments + biases to smem ----
  float4 w1h[FPW][2], w1l[FPW][2];
#pragma unroll
  for (int f = 0; f < FPW; ++f) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f0 + f < n_fields)
        w = __ldg(reinterpret_cast<const float4*>(a.w1 + (size_t)(8 * j + g) * kdim + 16 * (f0 + f) + 4 * t));
      uint32_t h[4], l[4];
      split_rna(w.x, h[0], l[0]);
      split_rna(w.y, h[1], l[1]);
      split_rna(w.z, h[2], l[2]);
      split_rna(w.w, h[3], l[3]);
      w1h[f][j] = make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
      w1l[f][j] = make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
    }
  }
  for (int i = threadIdx.x; i < a.hidden_layers * 4 * 32; i += blockDim.x) {
    const int l = i & 31, jk = (i >> 5) & 1, jn = (i >> 6) & 1, layer = i >> 7;
    const float* w = a.wh[layer] + (8 * jn + (l >> 2)) * 16 + 8 * jk + 2 * (l & 3);
    uint32_t h0, l0, h1, l1;
    split_rna(__ldg(w), h0, l0);
    split_rna(__ldg(w + 1), h1, l1);
    whs[(((layer * 2 + jn) * 2 + jk) * 2 + 0) * 32 + l] = make_float2(__uint_as_float(h0), __uint_as_float(h1));
    whs[(((layer * 2 + jn) * 2 + jk) * 2 + 1) * 32 + l] = make_float2(__uint_as_float(l0), __uint_as_float(l1));
  }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) {
    bias_s[i] = __ldg(a.b1 + i);
    for (int l = 0; l < a.hidden_layers; ++l) bias_s[(1 + l) * 16 + i] = __ldg(a.bh[l] + i);
    bias_s[(1 + kMaxHidden) * 16 + i] = __ldg(a.w_out + i);
  }
  if (threadIdx.x == 0) bias_s[(1 + kMaxHidden) * 16 + 16] = __ldg(a.b_out);
  __syncthreads();

  const int64_t tiles = (a.batch + kTile - 1) / kTile;
  const int64_t my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  float* my_v = vbuf + (size_t)warp * kStages * FPW * kFieldFloats;   // + stage * FPW * kFieldFloats

  // Index tiles: the (16 x fields) indices of a tile are contiguous in global memory; the CTA copies them with
  // cp.async into slot (tile % kIdxSlots) kIdxAhead tiles before `resolve` reads them.  (With ~1 300 rows in flight
  // per SM the loaded memory latency is several microseconds -- longer than one tile iteration -- so a register
  // prefetch one iteration ahead still stalls; profiles/r01_deepfm_packed_notes.md.)
  constexpr int kIdxBytes = IdxBits / 8;
  const int tile_idx_bytes = kTile * n_fields * kIdxBytes;   // multiple of 16
  const int64_t total_idx_bytes = a.batch * n_fields * kIdxBytes;
  auto issue_idx = [&](int64_t it) {
    if (it >= my_tiles) return;
    const int64_t tile = blockIdx.x + it * (int64_t)gridDim.x;
    const int64_t g0 = tile * tile_idx_bytes;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(idx_ring + (size_t)(it % kIdxSlots) * tile_idx_bytes));
    for (int c = threadIdx.x * 16; c < tile_idx_bytes; c += blockDim.x * 16) {
      const int64_t remain = total_idx_bytes - (g0 + c);
      const int sz = remain >= 16 ? 16 : (remain > 0 ? static_cast<int>(remain) : 0);   // zero-fill past the batch
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + c),
                   "l"(static_cast<const unsigned char*>(a.idx) + (sz > 0 ? g0 + c : 0)), "r"(sz) : "memory");
    }
  };
  int64_t foff[FPW];
#pragma unroll
  for (int f = 0; f < FPW; ++f) foff[f] = (f0 + f < n_fields) ? __ldg(a.offsets + f0 + f) : 0;
  // Row copies of tile `it` into `stage`.  EIGHT lanes per row, five active: chunks 0..3 = v, chunk 4 = the 16 bytes
  // holding w, so that [v|w] of a row is ONE 80-byte request of one warp instruction.  (A separate 4-byte read of w
  // from the same 128-byte line costs almost a full extra transaction: tools/gather_microbench.cu, 69 -> 121 us.)
  // Instruction i of the warp covers rows rho = 4i + (lane>>3), rho = f*16 + s  =>  f = i>>2, s = 4(i&3) + (lane>>3).
  const int sub = lane & 7, rsel = lane >> 3;
  auto issue = [&](int64_t it, int stage) {
    const int64_t b0 = (blockIdx.x + it * (int64_t)gridDim.x) * kTile;
    const bool tile_ok = it < my_tiles;
    const unsigned char* slot = idx_ring + (size_t)(it % kIdxSlots) * tile_idx_bytes;
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(my_v + (size_t)stage * FPW * kFieldFloats));
#pragma unroll
    for (int i = 0; i < FPW * 4; ++i) {
      const int f = i >> 2, s = 4 * (i & 3) + rsel;
      const bool live = tile_ok && f0 + f < n_fields && b0 + s < a.batch;
      int64_t ix = 0;
      if (live) {
        if (IdxBits == 64) ix = reinterpret_cast<const long long*>(slot)[s * n_fields + f0 + f];
        else ix = reinterpret_cast<const int*>(slot)[s * n_fields + f0 + f];
      }
      const int64_t r = ix + foff[f];
      const bool in = live && r >= 0 && r < a.rows;
      if (live && !in && sub == 0) report_oob(a.status, (b0 + s) * n_fields + f0 + f);
      const float* src = a.packed + (in ? r * kRowFloats : 0) + 4 * sub;
      const int dst_f = f * kFieldFloats + (s >> 1) * kPairFloats + (sub < 4 ? (s & 1) * 16 + 4 * sub : 32 + (s & 1) * 4);
      if (sub < 5) cp_async16(base + dst_f * 4, src, in);
    }
  };

  // ---- prologue: index tiles 0..kIdxAhead-1, then rows of tiles 0..kStages-2 ---------------------------------------
  for (int s = 0; s < kIdxAhead; ++s) issue_idx(s);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    issue(s, s);
    cp_async_commit();
  }

  for (int64_t it = 0; it < my_tiles; ++it) {
    const int stage = static_cast<int>(it % kStages);
    // rows of tile it + kStages - 1 (its index tile landed >= 1 barrier ago; zeros past the end) + a new index tile
    issue(it + kStages - 1, static_cast<int>((it + kStages - 1) % kStages));
    issue_idx(it + kIdxAhead);
    cp_async_commit();
    cp_async_wait<kStages - 1>();   // this lane's copies of tile `it` have landed ...
    __syncwarp();                   // ... and so have the other lanes' (rows are copied and consumed by different lanes)

    // ---- this warp's fields of tile `it`: FM sums, first-order, layer-1 partial accumulators ----------------------
    const float* sv = my_v + (size_t)stage * FPW * kFieldFloats + (g >> 1) * kPairFloats + (g & 1) * 16 + 4 * t;
    const float* sw = my_v + (size_t)stage * FPW * kFieldFloats + (g >> 1) * kPairFloats + 32 + (g & 1) * 4;
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    float qsa = 0.f, qsb = 0.f;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
      const float4 va = *reinterpret_cast<const float4*>(sv + f * kFieldFloats);                    // sample g
      const float4 vb = *reinterpret_cast<const float4*>(sv + f * kFieldFloats + 4 * kPairFloats);  // sample g + 8
      sa.x += va.x; sa.y += va.y; sa.z += va.z; sa.w += va.w;
      sb.x += vb.x; sb.y += vb.y; sb.z += vb.z; sb.w += vb.w;
      qsa = fmaf(va.x, va.x, qsa); qsa = fmaf(va.y, va.y, qsa); qsa = fmaf(va.z, va.z, qsa); qsa = fmaf(va.w, va.w, qsa);
      qsb = fmaf(vb.x, vb.x, qsb); qsb = fmaf(vb.y, vb.y, qsb); qsb = fmaf(vb.z, vb.z, qsb); qsb = fmaf(vb.w, vb.w, qsb);
      uint32_t ah0[4], al0[4], ah1[4], al1[4];
      split_fast(va.x, ah0[0], al0[0]); split_fast(vb.x, ah0[1], al0[1]);
      split_fast(va.y, ah0[2], al0[2]); split_fast(vb.y, ah0[3], al0[3]);
      split_fast(va.z, ah1[0], al1[0]); split_fast(vb.z, ah1[1], al1[1]);
      split_fast(va.w, ah1[2], al1[2]); split_fast(vb.w, ah1[3], al1[3]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 wh = w1h[f][j], wl = w1l[f][j];
        mma_3x(acc[j], ah0, al0, __float_as_uint(wh.x), __float_as_uint(wh.y), __float_as_uint(wl.x),
               __float_as_uint(wl.y));
        mma_3x(acc[j], ah1, al1, __float_as_uint(wh.z), __float_as_uint(wh.w), __float_as_uint(wl.z),
               __float_as_uint(wl.w));
      }
    }
    float ca = -0.5f * qsa, cb = -0.5f * qsb;
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
      if ((f & 3) == t) {   // lane t adds the first-order weights of fields f = t, t+4 of its two samples
        ca += sw[f * kFieldFloats];
        cb += sw[f * kFieldFloats + 4 * kPairFloats];
      }
    }
    // ---- publish partials -------------------------------------------------------------------------------------------
    float* pw = pbuf + ((size_t)(it & 1) * kWarps + warp) * kPartial * 32 + lane;
    pw[0 * 32] = acc[0][0]; pw[1 * 32] = acc[0][1]; pw[2 * 32] = acc[0][2]; pw[3 * 32] = acc[0][3];
    pw[4 * 32] = acc[1][0]; pw[5 * 32] = acc[1][1]; pw[6 * 32] = acc[1][2]; pw[7 * 32] = acc[1][3];
    pw[8 * 32] = sa.x; pw[9 * 32] = sa.y; pw[10 * 32] = sa.z; pw[11 * 32] = sa.w;
    pw[12 * 32] = sb.x; pw[13 * 32] = sb.y; pw[14 * 32] = sb.z; pw[15 * 32] = sb.w;
    pw[16 * 32] = ca; pw[17 * 32] = cb;
    __syncthreads();

    if (warp == static_cast<int>(it % kWarps)) {
      // ---- finisher: reduce the 8 partials, FM, MLP tail, store 16 logits -------------------------------------------
      float r[kPartial];
#pragma unroll
      for (int k = 0; k < kPartial; ++k) r[k] = 0.f;
      const float* pr = pbuf + (size_t)(it & 1) * kWarps * kPartial * 32 + lane;
#pragma unroll
      for (int w = 0; w < kWarps; ++w)
#pragma unroll
        for (int k = 0; k < kPartial; ++k) r[k] += pr[(w * kPartial + k) * 32];
      float side_a = 0.5f * (r[8] * r[8] + r[9] * r[9] + r[10] * r[10] + r[11] * r[11]) + r[16];
      float side_b = 0.5f * (r[12] * r[12] + r[13] * r[13] + r[14] * r[14] + r[15] * r[15]) + r[17];
      float h[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float b0v = bias_s[8 * j + 2 * t], b1v = bias_s[8 * j + 2 * t + 1];
        h[j][0] = fmaxf(r[4 * j + 0] + b0v, 0.f);
        h[j][1] = fmaxf(r[4 * j + 1] + b1v, 0.f);
        h[j][2] = fmaxf(r[4 * j + 2] + b0v, 0.f);
        h[j][3] = fmaxf(r[4 * j + 3] + b1v, 0.f);
      }
      for (int layer = 0; layer < a.hidden_layers; ++layer) {
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int jk = 0; jk < 2; ++jk) {
          uint32_t ah[4], al[4];
          split_rna(h[jk][0], ah[0], al[0]);
          split_rna(h[jk][2], ah[1], al[1]);
          split_rna(h[jk][1], ah[2], al[2]);
          split_rna(h[jk][3], ah[3], al[3]);
#pragma unroll
          for (int jn = 0; jn < 2; ++jn) {
            const float2 wh = whs[(((layer * 2 + jn) * 2 + jk) * 2 + 0) * 32 + lane];
            const float2 wl = whs[(((layer * 2 + jn) * 2 + jk) * 2 + 1) * 32 + lane];
            mma_3x(o[jn], ah, al, __float_as_uint(wh.x), __float_as_uint(wh.y), __float_as_uint(wl.x),
                   __float_as_uint(wl.y));
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float b0v = bias_s[(1 + layer) * 16 + 8 * j + 2 * t];
          const float b1v = bias_s[(1 + layer) * 16 + 8 * j + 2 * t + 1];
          h[j][0] = fmaxf(o[j][0] + b0v, 0.f);
          h[j][1] = fmaxf(o[j][1] + b1v, 0.f);
          h[j][2] = fmaxf(o[j][2] + b0v, 0.f);
          h[j][3] = fmaxf(o[j][3] + b1v, 0.f);
        }
      }
      const float* wo = bias_s + (1 + kMaxHidden) * 16;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float w0 = wo[8 * j + 2 * t], w1v = wo[8 * j + 2 * t + 1];
        side_a = fmaf(h[j][0], w0, side_a);
        side_a = fmaf(h[j][1], w1v, side_a);
        side_b = fmaf(h[j][2], w0, side_b);
        side_b = fmaf(h[j][3], w1v, side_b);
      }
      side_a += __shfl_xor_sync(0xffffffffu, side_a, 1);
      side_b += __shfl_xor_sync(0xffffffffu, side_b, 1);
      side_a += __shfl_xor_sync(0xffffffffu, side_a, 2);
      side_b += __shfl_xor_sync(0xffffffffu, side_b, 2);
      if (t == 0) {
        const int64_t b0 = (blockIdx.x + it * (int64_t)gridDim.x) * kTile;
        const float bo = wo[16];
        if (b0 + g < a.batch) a.logits[b0 + g] = side_a + bo;
        if (b0 + g + 8 < a.batch) a.logits[b0 + g + 8] = side_b + bo;
      }
    }
  }
  cp_async_wait<0>();
}

// ---- packing kernel: packed[r] = [ w_emb[r][0..15], w_feat[r], 0 x 15 ] ---------------------------------------------
__global__ void __launch_bounds__(256) pack_table_kernel(const float4* __restrict__ w_emb,
                                                         const float* __restrict__ w_feat, int64_t rows,
                                                         float4* __restrict__ packed) {
  const int64_t items = rows * 8;  // 8 x 16-byte chunks per packed row
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < items; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 3;
    const int c = static_cast<int>(i & 7);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < 4) v = __ldg(w_emb + r * 4 + c);
    else if (c == 4) v.x = __ldg(w_feat + r);
    packed[i] = v;
  }
}

}  // namespace

int deepfm_packed_supported(int fields, int embed, const int* mlp_dims, int mlp_layers, int activation,
                            int64_t rows) {
  if (embed != 16 || activation != TRS_ACT_RELU || rows >= (int64_t(1) << 31)) return 0;
  if (mlp_layers < 2 || mlp_layers - 2 > kMaxHidden) return 0;
  if (mlp_dims[0] != fields * 16 || mlp_dims[mlp_layers] != 1) return 0;
  for (int l = 1; l < mlp_layers; ++l)
    if (mlp_dims[l] != 16) return 0;
  return fields >= 1 && fields <= 5 * kWarps;
}

template <int IdxBits, int FPW>
static int launch_packed(const PackedArgs& a, cudaStream_t s) {
  const size_t smem = Smem<FPW>::total(a.fields, IdxBits);
  TRS_SMEM_OPT_IN((deepfm_packed_kernel<IdxBits, FPW>));
  const int64_t tiles = (a.batch + kTile - 1) / kTile;
  const int grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  deepfm_packed_kernel<IdxBits, FPW><<<grid, kWarps * 32, smem, s>>>(a);
  return check_launch("deepfm_packed_kernel");
}

}  // namespace trs

using namespace trs;

extern "C" int trs_fm_pack_table(const float* w_emb, const float* w_feat, int64_t rows, int embed, float* packed,
                                 void* stream) {
  TRS_REQUIRE(w_emb && w_feat && packed, "trs_fm_pack_table: null pointer");
  TRS_REQUIRE(rows > 0, "trs_fm_pack_table: bad sizes");
  TRS_UNSUPPORTED(embed != 16, "trs_fm_pack_table: the packed layout is defined for embed_size 16 (got %d)", embed);
  TRS_REQUIRE(aligned16(w_emb) && (reinterpret_cast<uintptr_t>(packed) & 127u) == 0,
              "trs_fm_pack_table: w_emb must be 16-byte and packed 128-byte aligned");
  const int grid = grid_for(rows * 8, 256, 8);
  pack_table_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(w_emb),
                                                                         w_feat, rows,
                                                                         reinterpret_cast<float4*>(packed));
  return check_launch("pack_table_kernel");
}

extern "C" int trs_deepfm_forward_packed(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                         int fields, const float* packed, int64_t rows, const int* mlp_dims,
                                         int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                                         int activation, float* logits, int32_t* status, void* stream) {
  TRS_REQUIRE(idx && offsets && packed && logits && mlp_dims && mlp_w && mlp_b,
              "trs_deepfm_forward_packed: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_deepfm_forward_packed: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0 && mlp_layers >= 1, "trs_deepfm_forward_packed: bad sizes");
  TRS_UNSUPPORTED(!deepfm_packed_supported(fields, 16, mlp_dims, mlp_layers, activation, rows),
                  "trs_deepfm_forward_packed: needs embed 16, hidden widths 16, ReLU, <= 40 fields, < 2^31 rows");
  TRS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127u) == 0 && aligned16(mlp_w[0]) && aligned16(idx),
              "trs_deepfm_forward_packed: packed table must be 128-byte aligned, W1 and idx 16-byte aligned");
  if (batch == 0) return TRS_OK;
  PackedArgs a{};
  a.idx = idx; a.offsets = offsets; a.packed = packed; a.logits = logits; a.status = status;
  a.batch = batch; a.rows = rows; a.fields = fields;
  a.hidden_layers = mlp_layers - 2;
  for (int l = 0; l < mlp_layers; ++l) TRS_REQUIRE(mlp_w[l] && mlp_b[l], "trs_deepfm_forward_packed: null MLP parameter");
  a.w1 = mlp_w[0];
  a.b1 = mlp_b[0];
  for (int l = 0; l < a.hidden_layers; ++l) {
    a.wh[l] = mlp_w[1 + l];
    a.bh[l] = mlp_b[1 + l];
  }
  a.w_out = mlp_w[mlp_layers - 1];
  a.b_out = mlp_b[mlp_layers - 1];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int fpw = (fields + kWarps - 1) / kWarps;
#define DISPATCH(FPW)                                                        \
  case FPW:                                                                  \
    return idx_bits == 64 ? launch_packed<64, FPW>(a, s) : launch_packed<32, FPW>(a, s);
  switch (fpw) {
    DISPATCH(1)
    DISPATCH(2)
    DISPATCH(3)
    DISPATCH(4)
    DISPATCH(5)
  }
#undef DISPATCH
  set_error("trs_deepfm_forward_packed: unsupported field count %d", fields);
  return TRS_ERR_UNSUPPORTED;
}
