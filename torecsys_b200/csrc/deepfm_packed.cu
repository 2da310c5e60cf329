// DeepFM forward on a PACKED table (embed = 16, hidden widths = 16): the B200-first layout of the hot path.
//
// Measured on B200 (tools/gather_microbench.cu, profiles/r01_gather_microbench_ncu.csv): random row reads are bound
// by the number of memory REQUESTS, not by bytes -- the chip sustains ~32-38 G random requests/s whatever the row
// size up to 128 B, DRAM always moves the whole 128-byte line, and a second request to the same line from another
// instruction costs almost as much as the first (64 B + a separate 4 B read of the same line: 69 -> 121 us).
// The reference's two tables (emb (R,16) and first-order (R,1), same row ids) therefore cost 78 requests per sample;
// packing them as one 128-byte-aligned row  [ v0..v15 | w | pad ]  (trs_fm_pack_table) and reading [v|w] with ONE
// warp instruction makes it 39, the floor for this model.  The packed table is a shadow of the registered
// parameters (rebuilt by the host layer when their version changes), 25.6 GB for 200 M rows.
//
// Kernel shape (one persistent CTA of 8 warps per SM, grid = 148):
//   * a tile = 16 samples; warp w owns fields [w*FPW, (w+1)*FPW) of EVERY tile, so its slice of W1 (B fragments of
//     mma.sync.m16n8k8 TF32, pre-split hi/lo) lives in its registers -- no shared memory, no L2 traffic for W1;
//   * rows travel global -> shared with cp.async (LDGSTS) into a per-warp ring of kStages tiles, so in-flight loads
//     cost no registers (8 warps x 2 stages x 16 x FPW rows = 1 248 rows in flight per SM); eight lanes per row copy
//     [v|w] as ONE 80-byte request; each lane resolves (index + offset, range check) at most three rows of the
//     warp's 16 x FPW and the copy loop gets the row ids by shuffle; the (16 x fields) index tiles have their own,
//     deeper cp.async ring because the loaded memory latency (several microseconds) exceeds one tile iteration;
//   * lane (g,t) reads chunk t of the rows of samples g and g+8 = exactly its mma A-fragment elements; per tile each
//     warp publishes partial layer-1 accumulators, partial FM sums and first-order sums for its fields; after ONE
//     bar.sync per tile every warp finishes two of the 16 samples (reduce the 8 partials, bias + ReLU, the 16x16
//     hidden layers and the output layer in plain FP32 FFMA with shuffles, FM, logit).
// Layer 1 (9 984 MACs/sample) runs on the tensor pipe, FP32-accurate through the 3xTF32 split; the roofline of the
// kernel is the HBM request rate, not the tensor pipe.
#include <stdlib.h>

#include "tile_ops.cuh"

namespace trs {
int mlp_chain_supported(const int* dims, int layers, int64_t rows, const void* x, const void* out, int accumulate);
int mlp_chain_gather_supported(const int* dims, int layers, int fields, int embed, int use_fm);
int mlp_chain_run(const float* x, int64_t rows, const MlpParams& mp, float* out, int accumulate, cudaStream_t s,
                  const DenseFuse* gather);
namespace {

constexpr int kWarps = 8;
constexpr int kTile = 16;
constexpr int kStages = 3;                  // row ring depth (kStages-1 tiles of rows in flight)
constexpr int kIdxAhead = 2 * kStages - 1;  // index tiles are requested this many tiles ahead of their use ...
constexpr int kIdxSlots = kIdxAhead;        // ... into a ring of this many slots
constexpr int kMaxHidden = 4;
constexpr int kRowFloats = 32;              // packed row pitch: 128 B

// Row staging layout of one (warp, stage): [field f][sample pair k = s/2][40 floats]:
//   floats  0..15 = v of sample 2k, 16..31 = v of sample 2k+1, 32..35 / 36..39 = the 16-byte chunk holding w of each.
// The two rows of a pair fill the 32 banks exactly once, so the consumers' LDS.128 (quarter-warp = two samples x four
// chunks) are conflict-free, and every cp.async destination is 16-byte aligned.
constexpr int kPairFloats = 40;
constexpr int kFieldFloats = 8 * kPairFloats;  // 16 samples of one field

// Partial exchange buffer of one (parity, warp): H[16 samples][24] layer-1 partials, S[16][16] FM sums,
// C[16][4] = first-order - 0.5 * sum of squares (per t lane).  Pitches chosen for conflict-free publishing stores.
constexpr int kHPitch = 24;
constexpr int kPartialFloats = 16 * kHPitch + 16 * 16 + 16 * 4;  // 704
constexpr int kHidPitch = 20;   // 16-byte aligned rows; 8 consecutive rows cover the 32 banks once (LDS.128)

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
// d0 += A_lo*B_hi, d1 += A_hi*B_lo, d2 += A_hi*B_hi: three independent accumulator chains (summed once per tile)
// instead of one 3x longer dependent chain -- with two warps per scheduler the MMA latency is otherwise exposed
__device__ __forceinline__ void mma_3x(float (&d0)[4], float (&d1)[4], float (&d2)[4], const uint32_t (&ah)[4],
                                       const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0,
                                       uint32_t bl1) {
  mma_tf32(d0, al, bh0, bh1);
  mma_tf32(d1, ah, bl0, bl1);
  mma_tf32(d2, ah, bh0, bh1);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  // src_bytes < 16 => the remaining destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Programmatic dependent launch (TRS_LAUNCH_OVERLAP_PREVIOUS): every CTA lets the NEXT grid on the stream be scheduled
// as soon as SMs free up (launch_dependents at entry), and a grid launched that way only READS its inputs until the
// previous grid has completed and flushed (wait before the first global write).  Without the launch attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// mbarriers (arrival count = one elected lane per warp) that order the partial-exchange buffers between warps
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TRS_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TRS_MBAR_DONE;\n"
      "bra TRS_MBAR_WAIT;\n"
      "TRS_MBAR_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
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

template <int FPW>
struct Smem {
  static constexpr size_t v_bytes = (size_t)kWarps * kStages * FPW * kFieldFloats * sizeof(float);
  static constexpr size_t p_bytes = (size_t)2 * kWarps * kPartialFloats * sizeof(float);
  static constexpr size_t h_bytes = (size_t)kMaxHidden * 16 * kHidPitch * sizeof(float);
  static constexpr size_t b_bytes = ((1 + kMaxHidden) * 16 + 16 + 4) * sizeof(float);
  static constexpr size_t m_bytes = 4 * sizeof(unsigned long long);   // full[2], empty[2] mbarriers
  static constexpr size_t fixed = v_bytes + p_bytes + h_bytes + b_bytes + m_bytes;
  // + field offsets (int64 per field, padded to 16 B)
  // + per-warp index rings (kWarps x kIdxSlots x [16 samples][FPW fields] indices)
  __host__ __device__ static size_t off_bytes(int fields) { return (((size_t)fields * 8 + 15) / 16) * 16; }
  static size_t total(int fields, int idx_bits) {
    return fixed + off_bytes(fields) + (size_t)kWarps * kIdxSlots * kTile * FPW * (idx_bits / 8);
  }
};

template <int IdxBits, int FPW>
__global__ void __launch_bounds__(kWarps * 32, 1) deepfm_packed_kernel(PackedArgs a) {
  using S = Smem<FPW>;
  constexpr int kRowsPerWarp = 16 * FPW;                 // rows of one tile copied by one warp
  constexpr int kResolve = (kRowsPerWarp + 31) / 32;     // rows resolved per lane
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* vbuf = reinterpret_cast<float*>(smem_raw);
  float* pbuf = reinterpret_cast<float*>(smem_raw + S::v_bytes);
  float* hid_s = reinterpret_cast<float*>(smem_raw + S::v_bytes + S::p_bytes);   // [layer][16][17]
  float* bias_s = reinterpret_cast<float*>(smem_raw + S::v_bytes + S::p_bytes + S::h_bytes);
  const uint32_t mbar_s =
      static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw + S::v_bytes + S::p_bytes + S::h_bytes + S::b_bytes));
  long long* off_s = reinterpret_cast<long long*>(smem_raw + S::fixed);
  unsigned char* idx_rings = smem_raw + S::fixed + S::off_bytes(a.fields);   // [warp][kIdxSlots][16][FPW] indices

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n_fields = a.fields;
  const int kdim = 16 * n_fields;
  const int f0 = warp * FPW;  // first field of this warp

  const int64_t tiles = (a.batch + kTile - 1) / kTile;
  const int my_tiles = blockIdx.x < tiles ? static_cast<int>((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  // Index slices: every warp keeps its OWN ring of the (16 samples x FPW fields) indices it resolves, copied with
  // 8-/4-byte cp.async kIdxAhead tiles before the row copies read them -- no other warp is involved, so the ring
  // needs no CTA-wide barrier (lane e copies element (s, f) = (e / FPW, e % FPW): a sample's fields are contiguous).
  constexpr int kIdxBytes = IdxBits / 8;
  constexpr int kSliceBytes = kTile * FPW * kIdxBytes;
  unsigned char* my_idx = idx_rings + (size_t)warp * kIdxSlots * kSliceBytes;
  const uint32_t my_idx_s = static_cast<uint32_t>(__cvta_generic_to_shared(my_idx));
  auto issue_idx = [&](int it) {
    const int64_t b0 = (blockIdx.x + it * (int64_t)gridDim.x) * kTile;
    const uint32_t dst = my_idx_s + (it % kIdxSlots) * kSliceBytes;
#pragma unroll
    for (int k = 0; k < kResolve; ++k) {
      const int e = lane + 32 * k;
      if (e < kRowsPerWarp) {
        const int s_ = e / FPW, f = e - s_ * FPW;
        const bool live = it < my_tiles && f0 + f < n_fields && b0 + s_ < a.batch;
        const unsigned char* src = static_cast<const unsigned char*>(a.idx) +
                                   (live ? ((b0 + s_) * n_fields + f0 + f) * kIdxBytes : 0);
        if (IdxBits == 64) cp_async8(dst + e * 8, src, live ? 8 : 0);
        else cp_async4(dst + e * 4, src, live ? 4 : 0);
      }
    }
  };

  // the first index slices are requested before anything else so that they fly during the weight set-up below
  for (int s = 0; s < kIdxAhead; ++s) issue_idx(s);
  cp_async_commit();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(mbar_s + 8 * i, kWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- one-time: this warp's W1 B-fragments into registers (hi/lo); hidden layers, biases, offsets to smem ----------
  float4 w1h[FPW][2], w1l[FPW][2];
#pragma unroll
  for (int f = 0; f < FPW; ++f) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f0 + f < n_fields && a.w1 != nullptr)   // w1 == null: no deep part (FactorizationMachineModel)
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
  for (int i = threadIdx.x; i < a.hidden_layers * 256; i += blockDim.x) {
    const int layer = i >> 8, o = (i >> 4) & 15, k = i & 15;
    hid_s[(layer * 16 + o) * kHidPitch + k] = __ldg(a.wh[layer] + o * 16 + k);
  }
  for (int i = threadIdx.x; i < 16; i += blockDim.x) {
    bias_s[i] = a.b1 ? __ldg(a.b1 + i) : 0.f;
    for (int l = 0; l < a.hidden_layers; ++l) bias_s[(1 + l) * 16 + i] = __ldg(a.bh[l] + i);
    bias_s[(1 + kMaxHidden) * 16 + i] = a.w_out ? __ldg(a.w_out + i) : 0.f;
  }
  if (threadIdx.x == 0) bias_s[(1 + kMaxHidden) * 16 + 16] = a.b_out ? __ldg(a.b_out) : 0.f;
  for (int i = threadIdx.x; i < n_fields; i += blockDim.x) off_s[i] = __ldg(a.offsets + i);

  float* my_v = vbuf + (size_t)warp * kStages * FPW * kFieldFloats;   // + stage * FPW * kFieldFloats
  const uint32_t my_v_s = static_cast<uint32_t>(__cvta_generic_to_shared(my_v));

  // Row copies of tile `it` into `stage`.
  //   phase 1: lane resolves rows rho = lane + 32k (rho = f*16 + s): index + field offset, range check -> row id / -1;
  //   phase 2: EIGHT lanes per row, five active (chunks 0..3 = v, chunk 4 = the 16 bytes holding w): instruction i of
  //            the warp covers rows rho = 4i + (lane>>3); the row id comes from lane (rho & 31), register rho >> 5.
  const int sub = lane & 7, rsel = lane >> 3;
  const int lane_dst = (rsel >> 1) * kPairFloats + (sub < 4 ? (rsel & 1) * 16 + 4 * sub : 32 + (rsel & 1) * 4);
  const unsigned long long src_lane = reinterpret_cast<unsigned long long>(a.packed + 4 * sub);
  auto issue = [&](int it, int stage) {
    const int64_t b0 = (blockIdx.x + it * (int64_t)gridDim.x) * kTile;
    const bool tile_ok = it < my_tiles;
    const unsigned char* slot = my_idx + (it % kIdxSlots) * kSliceBytes;
    int rid[kResolve];
#pragma unroll
    for (int k = 0; k < kResolve; ++k) {
      const int rho = lane + 32 * k;
      const int f = rho >> 4, s = rho & 15;
      const bool live = tile_ok && rho < kRowsPerWarp && f0 + f < n_fields && b0 + s < a.batch;
      rid[k] = -1;
      if (live) {
        int64_t ix;
        if (IdxBits == 64) ix = reinterpret_cast<const long long*>(slot)[s * FPW + f];
        else ix = reinterpret_cast<const int*>(slot)[s * FPW + f];
        const int64_t r = ix + off_s[f0 + f];
        if (r >= 0 && r < a.rows) rid[k] = static_cast<int>(r);
        else {
          pdl_wait();   // status may still be written (or cleared) by the work this launch overlaps
          report_oob(a.status, (b0 + s) * n_fields + f0 + f);
        }
      }
    }
    const uint32_t base = my_v_s + (stage * FPW * kFieldFloats + lane_dst) * 4;
#pragma unroll
    for (int i = 0; i < FPW * 4; ++i) {
      const int r = __shfl_sync(0xffffffffu, rid[i >> 3], 4 * (i & 7) + rsel);
      const int dst_f = (i >> 2) * kFieldFloats + 2 * (i & 3) * kPairFloats;   // + lane_dst (folded into base)
      // address = lane base + max(r, 0) * 128 in ONE mad.wide.u32; an invalid row (r < 0) copies nothing, zero-fills
      unsigned long long src;
      asm("mad.wide.u32 %0, %1, 128, %2;" : "=l"(src) : "r"(static_cast<unsigned>(max(r, 0))), "l"(src_lane));
      if (sub < 5) cp_async16(base + dst_f * 4, reinterpret_cast<const void*>(src), r >= 0 ? 16 : 0);
    }
  };

  // ---- prologue: index tiles 0..kIdxAhead-1 (requested at kernel entry), then rows of tiles 0..kStages-2 ----------
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    issue(s, s);
    cp_async_commit();
  }

  // finisher role of this lane: sample 2*warp + (lane>>4) of every tile, output / FM component (lane & 15)
  const int fs = 2 * warp + (lane >> 4), fo = lane & 15;

  // Finisher: lane = (sample fs, output / component fo) reduces the 8 warps' partials of tile `jt`, runs the 16x16
  // hidden layers, the output layer and FM, and stores two logits per warp.
  auto finish = [&](int jt) {
    const int p = jt & 1;
    mbar_wait(mbar_s + 8 * p, (jt >> 1) & 1);          // every warp has published its partials of tile jt
    const float* pr = pbuf + (size_t)p * kWarps * kPartialFloats;
    float h = bias_s[fo], sx = 0.f, c = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      h += pr[w * kPartialFloats + fs * kHPitch + fo];
      sx += pr[w * kPartialFloats + 16 * kHPitch + fs * 16 + fo];
    }
    // the 32 C partials of the sample (8 warps x 4 t), two per lane
    c = pr[(fo >> 1) * kPartialFloats + 16 * kHPitch + 256 + fs * 4 + 2 * (fo & 1)] +
        pr[(fo >> 1) * kPartialFloats + 16 * kHPitch + 256 + fs * 4 + 2 * (fo & 1) + 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(mbar_s + 16 + 8 * p);    // this warp no longer reads buffer p
    h = fmaxf(h, 0.f);
    const int src_base = lane & 16;
    for (int layer = 0; layer < a.hidden_layers; ++layer) {
      float o = bias_s[(1 + layer) * 16 + fo];
      const float4* wr = reinterpret_cast<const float4*>(hid_s + (layer * 16 + fo) * kHidPitch);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w4 = wr[k];
        o = fmaf(w4.x, __shfl_sync(0xffffffffu, h, src_base + 4 * k), o);
        o = fmaf(w4.y, __shfl_sync(0xffffffffu, h, src_base + 4 * k + 1), o);
        o = fmaf(w4.z, __shfl_sync(0xffffffffu, h, src_base + 4 * k + 2), o);
        o = fmaf(w4.w, __shfl_sync(0xffffffffu, h, src_base + 4 * k + 3), o);
      }
      h = fmaxf(o, 0.f);
    }
    float side = fmaf(0.5f * sx, sx, c);
    side = fmaf(h, bias_s[(1 + kMaxHidden) * 16 + fo], side);
    side += __shfl_xor_sync(0xffffffffu, side, 8);
    side += __shfl_xor_sync(0xffffffffu, side, 4);
    side += __shfl_xor_sync(0xffffffffu, side, 2);
    side += __shfl_xor_sync(0xffffffffu, side, 1);
    if (fo == 0) {
      if (jt == 0) pdl_wait();   // first global write: the overlapped previous grid must be complete
      const int64_t b = (blockIdx.x + jt * (int64_t)gridDim.x) * kTile + fs;
      if (b < a.batch) a.logits[b] = side + bias_s[(1 + kMaxHidden) * 16 + 16];
    }
  };

  // Main loop.  The warps are NOT barrier-synchronised per tile: warp w publishes its partials of tile `it` into
  // buffer it & 1, arrives on full[it & 1] and goes on; it finishes tile it - 1 (whose partials the other warps
  // published about one tile ago) afterwards, so a warp only ever waits for a warp that is a whole tile behind.
  // empty[p] keeps a fast warp from overwriting buffer p before every warp has read tile it - 2 out of it.
  int stage = 0, fill_stage = kStages - 1;
  for (int it = 0; it < my_tiles; ++it) {
    // rows of tile it + kStages - 1 (its index slice landed two iterations ago; zeros past the end) + a new index slice
    issue(it + kStages - 1, fill_stage);
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
    float acl[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // A_lo * B_hi terms
    float acm[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};   // A_hi * B_lo terms
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
        mma_3x(acl[j], acm[j], acc[j], ah0, al0, __float_as_uint(wh.x), __float_as_uint(wh.y),
               __float_as_uint(wl.x), __float_as_uint(wl.y));
        mma_3x(acl[j], acm[j], acc[j], ah1, al1, __float_as_uint(wh.z), __float_as_uint(wh.w),
               __float_as_uint(wl.z), __float_as_uint(wl.w));
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j][k] += acl[j][k] + acm[j][k];
    float ca = -0.5f * qsa, cb = -0.5f * qsb;
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
      if ((f & 3) == t) {   // lane t adds the first-order weights of fields f = t, t+4 of its two samples
        ca += sw[f * kFieldFloats];
        cb += sw[f * kFieldFloats + 4 * kPairFloats];
      }
    }
    // ---- publish partials: acc[j] = {(g, 8j+2t), (g, 8j+2t+1), (g+8, 8j+2t), (g+8, 8j+2t+1)} ---------------------------
    const int p = it & 1;
    if (it >= 2) mbar_wait(mbar_s + 16 + 8 * p, ((it >> 1) - 1) & 1);   // tile it - 2 has been read out of buffer p
    float* pw = pbuf + ((size_t)p * kWarps + warp) * kPartialFloats;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      *reinterpret_cast<float2*>(pw + g * kHPitch + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
      *reinterpret_cast<float2*>(pw + (g + 8) * kHPitch + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    *reinterpret_cast<float4*>(pw + 16 * kHPitch + g * 16 + 4 * t) = sa;
    *reinterpret_cast<float4*>(pw + 16 * kHPitch + (g + 8) * 16 + 4 * t) = sb;
    pw[16 * kHPitch + 256 + g * 4 + t] = ca;
    pw[16 * kHPitch + 256 + (g + 8) * 4 + t] = cb;
    __syncwarp();
    if (lane == 0) mbar_arrive(mbar_s + 8 * p);
    if (it > 0) finish(it - 1);
    stage = stage + 1 == kStages ? 0 : stage + 1;
    fill_stage = fill_stage + 1 == kStages ? 0 : fill_stage + 1;
  }
  if (my_tiles > 0) finish(my_tiles - 1);
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
static int launch_packed(const PackedArgs& a, cudaStream_t s, unsigned flags = 0) {
  const size_t smem = Smem<FPW>::total(a.fields, IdxBits);
  TRS_REQUIRE(smem <= (size_t)kMaxDynSmem, "trs_deepfm_forward_packed: shared memory budget exceeded (%zu B)", smem);
  TRS_SMEM_OPT_IN((deepfm_packed_kernel<IdxBits, FPW>));
  const int64_t tiles = (a.batch + kTile - 1) / kTile;
  const int grid = static_cast<int>(tiles < kNumSMs ? tiles : kNumSMs);
  if (flags & TRS_LAUNCH_OVERLAP_PREVIOUS) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, deepfm_packed_kernel<IdxBits, FPW>, a);
    if (e != cudaSuccess) {
      set_error("launch of deepfm_packed_kernel (overlapped) failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return TRS_ERR_CUDA;
    }
    return TRS_OK;
  }
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

// FactorizationMachineModel on the packed table = the DeepFM kernel with no deep part: logit = FM + first-order + bias
extern "C" int trs_fm_model_forward_packed(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                           int fields, const float* packed, int64_t rows, const float* bias,
                                           float* logits, int32_t* status, void* stream) {
  TRS_REQUIRE(idx && offsets && packed && logits, "trs_fm_model_forward_packed: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_fm_model_forward_packed: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0, "trs_fm_model_forward_packed: bad sizes");
  TRS_UNSUPPORTED(fields > 5 * kWarps || rows >= (int64_t(1) << 31),
                  "trs_fm_model_forward_packed: needs <= 40 fields and < 2^31 rows");
  TRS_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127u) == 0 && aligned16(idx),
              "trs_fm_model_forward_packed: packed table must be 128-byte aligned, idx 16-byte aligned");
  if (batch == 0) return TRS_OK;
  PackedArgs a{};
  a.idx = idx; a.offsets = offsets; a.packed = packed; a.logits = logits; a.status = status;
  a.batch = batch; a.rows = rows; a.fields = fields;
  a.hidden_layers = 0;
  a.b_out = bias;   // w1, b1, w_out stay null: zero deep part
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch ((fields + kWarps - 1) / kWarps) {
    case 1: return idx_bits == 64 ? launch_packed<64, 1>(a, s) : launch_packed<32, 1>(a, s);
    case 2: return idx_bits == 64 ? launch_packed<64, 2>(a, s) : launch_packed<32, 2>(a, s);
    case 3: return idx_bits == 64 ? launch_packed<64, 3>(a, s) : launch_packed<32, 3>(a, s);
    case 4: return idx_bits == 64 ? launch_packed<64, 4>(a, s) : launch_packed<32, 4>(a, s);
    case 5: return idx_bits == 64 ? launch_packed<64, 5>(a, s) : launch_packed<32, 5>(a, s);
  }
  set_error("trs_fm_model_forward_packed: unsupported field count %d", fields);
  return TRS_ERR_UNSUPPORTED;
}

extern "C" int trs_deepfm_forward_packed(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                         int fields, const float* packed, int64_t rows, const int* mlp_dims,
                                         int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                                         int activation, float* logits, int32_t* status, void* stream) {
  return trs_deepfm_forward_packed_ex(idx, idx_bits, offsets, batch, fields, packed, rows, mlp_dims, mlp_layers, mlp_w,
                                      mlp_b, activation, logits, status, 0u, stream);
}

// Does trs_deepfm_forward_packed[_ex] take this deep branch on the packed table through the gathering tcgen05 layer (a wide
// MLP: some layer >= 64 x 64, one output, batch >= 1 024)?  Narrow 16-wide branches take the packed-table kernels.
extern "C" int trs_deepfm_packed_wide_supported(int fields, const int* mlp_dims, int mlp_layers, int64_t batch) {
  if (mlp_dims == nullptr || mlp_layers < 2 || fields < 1) return 0;
  alignas(16) static const float probe[4] = {0.f, 0.f, 0.f, 0.f};   // (the chain test wants 16-byte aligned x / out)
  return mlp_chain_supported(mlp_dims, mlp_layers, batch, probe, probe, 1) &&
                 mlp_chain_gather_supported(mlp_dims, mlp_layers, fields, 16, 1)
             ? 1 : 0;
}

extern "C" int trs_deepfm_forward_packed_ex(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch,
                                            int fields, const float* packed, int64_t rows, const int* mlp_dims,
                                            int mlp_layers, const float* const* mlp_w, const float* const* mlp_b,
                                            int activation, float* logits, int32_t* status, unsigned flags,
                                            void* stream) {
  TRS_REQUIRE((flags & ~TRS_LAUNCH_OVERLAP_PREVIOUS) == 0, "trs_deepfm_forward_packed_ex: unknown flags 0x%x", flags);
  TRS_REQUIRE(idx && offsets && packed && logits && mlp_dims && mlp_w && mlp_b,
              "trs_deepfm_forward_packed: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_deepfm_forward_packed: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0 && mlp_layers >= 1, "trs_deepfm_forward_packed: bad sizes");
  // A paper-size deep branch on the packed table: the gathering tcgen05 layer of cin_tc.cu reads the 64-byte
  // row AND its first-order value out of the same 128-byte line -- the 4-byte lookups into a separate first-order table
  // cost as much DRAM traffic as all the rows.
  if (trs_deepfm_packed_wide_supported(fields, mlp_dims, mlp_layers, batch) && aligned16(logits)) {
    MlpParams mp;
    TRS_REQUIRE(fill_mlp_params(mp, mlp_dims, mlp_layers, mlp_w, mlp_b, activation) == 0,
                "trs_deepfm_forward_packed: bad MLP description");
    DenseFuse g;
    g.idx = idx; g.idx_bits = idx_bits; g.offsets = offsets; g.table = packed; g.w_feat = nullptr; g.bias = nullptr;
    g.status = status; g.table_rows = rows; g.fields = fields; g.embed = 16; g.use_fm = 1; g.row_pitch = 32; g.w_col = 16;
    return mlp_chain_run(nullptr, batch, mp, logits, 1, static_cast<cudaStream_t>(stream), &g);
  }
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
    return idx_bits == 64 ? launch_packed<64, FPW>(a, s, flags) : launch_packed<32, FPW>(a, s, flags);
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
