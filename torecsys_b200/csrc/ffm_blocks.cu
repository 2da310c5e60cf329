// FieldAwareFactorizationMachineModel forward on field-aware tables SHARDED over the GPUs of one NVSwitch box:
// the exchange of looked-up vectors of BASELINE.json configs[4], at >= 256-byte granularity and at half the naive
// volume (torecsys/models/ctr/field_aware_factorization_machine.py:39-81 over
// torecsys/inputs/base/multi_indices_field_aware_emb.py:49-54,90-111; the reference itself is single-device).
//
// Partition.  Rank m owns the tables {t : t % world == m} (S_m of them) and stores them INTERLEAVED per row id:
//     shard_m[r] = [ T_{m}[r] | T_{m+world}[r] | ... ]        pitch = ceil(N / world) * E floats
// so the vectors of one row id that live on one rank are ONE contiguous chunk of S_m * E * 4 bytes (320 B at N = 39,
// world = 8, E = 16).  Measured on the B200 box (tools/peer_probe.cu): random 64-byte peer reads move 425 GB/s over
// NVLink, one bulk copy per 320-byte chunk 685 GB/s (the copy engine: 822 GB/s).
//
// Work split ("rank blocks").  logit = sum_{i<j} <T_j[r_i], T_i[r_j]>.  The pairs whose two tables live on ranks k and m
// form the block (k, m): for the fields i_a of k and j_b of m it needs R[a] = chunk_m[r_{i_a}] (S_k whole chunks of
// rank m) and L[b] = chunk_k[r_{j_b}] (S_m whole chunks of rank k) and is sum_{a,b} <R[a][b], L[b][a]> -- every byte of
// every chunk is used.  Block (k, m) is reduced on k for the samples of one parity and on m for the others, so every
// dot product moves exactly ONE of its two vectors over NVLink (the minimum), every rank does the same amount of work
// and every rank's HBM serves the same amount.  The diagonal block (k, k) is all local.  Every rank therefore walks ALL
// samples of the global batch (row ids all-gathered, 4 bytes per lookup) and produces a partial logit per sample; the
// partial logits are summed and distributed by one reduce-scatter (NCCL) outside this file.
//
// Kernel: persistent CTA per SM; stage = the chunks of one sample (diagonal + its assigned partner blocks).  Warp 0 is
// the producer: it reads the sample's row ids and issues one bulk copy (cp.async.bulk global -> shared, the global
// address being local HBM or NVLink peer memory) per chunk, following a per-parity copy list built on the host
// (trs_ffm_shard_plan).  Warps 1.. consume: one warp per sample walks a per-parity item list (offsets of the two
// 16-byte pieces of each dot product), sums in a fixed order and writes the sample's partial logit.  The fetch of
// sample n + STAGES - 1 overlaps the reduction of sample n.
#include "tc5.cuh"

#include <vector>

namespace trs {
namespace {

using tc5::bulk_g2s;
using tc5::mbar_arrive;
using tc5::mbar_expect_tx;
using tc5::mbar_init;
using tc5::mbar_wait;
using tc5::smem_u32;

constexpr int kProducers = 4;     // a warp-wide bulk copy is issued lane by lane (~60 cycles each): the chunk copies
                                  // of a sample are spread over several warps (measured: 1 warp 2.9 ms)
constexpr int kConsumers = 6;
constexpr int kThreads = (kProducers + kConsumers) * 32;
constexpr int kMaxWorld = 8;
constexpr int kMaxStages = 16;
constexpr int kRowRing = 8;      // samples whose row ids are in flight / in shared memory

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct BlockArgs {
  const int32_t* rows_all;   // (batch_all, fields) global row ids
  int64_t batch_all;
  int fields;
  const unsigned char* shard[kMaxWorld];
  int64_t pitch_bytes;
  const int2* copy_tab;      // [2][max_copies]: .x = src rank | field << 8 | (bytes / 16) << 16, .y = dst byte offset
  const uint32_t* item_tab;  // [2][max_items]: float4 index of piece A | float4 index of piece B << 16
  int n_copies[2], n_items[2], tx_bytes[2];
  int max_copies, max_items;
  int stage_bytes, stages;
  const float* first;        // first-order term + bias of the samples [own_lo, own_hi) (may be null)
  int64_t own_lo, own_hi;
  float* partial;            // (batch_all,)
};

__global__ void __launch_bounds__(kThreads, 1) ffm_blocks_kernel(const BlockArgs a) {
  extern __shared__ __align__(128) unsigned char fb_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* stage0 = fb_smem;
  unsigned char* p = fb_smem + (size_t)a.stages * a.stage_bytes;
  int2* copy_s = reinterpret_cast<int2*>(p);            p += (size_t)2 * a.max_copies * sizeof(int2);
  uint32_t* item_s = reinterpret_cast<uint32_t*>(p);    p += (size_t)2 * a.max_items * sizeof(uint32_t);
  p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  int* rows_s = reinterpret_cast<int*>(p);              p += kProducers * kRowRing * 64 * sizeof(int);
  unsigned long long* shard_s = reinterpret_cast<unsigned long long*>(p);  p += kMaxWorld * 8;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p);      // full[stages], empty[stages]
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + a.stages);
  float* part = reinterpret_cast<float*>(bars + 2 * a.stages);   // [stages][kConsumers]

  for (int i = threadIdx.x; i < 2 * a.max_copies; i += kThreads) copy_s[i] = a.copy_tab[i];
  for (int i = threadIdx.x; i < 2 * a.max_items; i += kThreads) item_s[i] = a.item_tab[i];
  if (threadIdx.x < kMaxWorld) shard_s[threadIdx.x] = reinterpret_cast<unsigned long long>(a.shard[threadIdx.x]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(full0 + 8 * s, kProducers);
      mbar_init(empty0 + 8 * s, kConsumers);
    }
    tc5::fence_barrier_init();
  }
  __syncthreads();

  // sample of this CTA's n-th turn: row n of a (turns x grid) tiling, rotated by n so that the parities -- and with them
  // the 3-or-4-partner workloads -- alternate within every CTA (grid is even: a plain stride would pin one parity).
  // Only the last row can run past batch_all.  Walked incrementally (no division in the loops).
  const int grid = static_cast<int>(gridDim.x);
  const int turns = static_cast<int>((a.batch_all + grid - 1) / grid);
  struct Walk {
    int n, rot, grid;
    __device__ Walk(int n0, int cta, int g) : n(n0), rot((cta + n0) % g), grid(g) {}
    __device__ int64_t sample() const { return static_cast<int64_t>(n) * grid + rot; }
    __device__ void next() { ++n; if (++rot == grid) rot = 0; }
  };
  const int fields = a.fields;

  if (warp < kProducers) {
    // ------------------------------------------------------------------ producers: warp j issues the copies c = j, j + P, ..
    // (lane l: copy l * P + j) of every sample, with its own ring of row ids and its own share of the stage's byte count
    int* my_rows = rows_s + warp * kRowRing * 64;
    int2 mine[2];          // this lane's copy of a sample of parity 0 / 1 (.x = 0: none)
    uint32_t my_tx[2];     // bytes this warp fetches per sample of parity 0 / 1
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const int c = lane * kProducers + warp;
      mine[par] = c < a.n_copies[par] ? copy_s[par * a.max_copies + c] : make_int2(0, 0);
      uint32_t b = static_cast<uint32_t>(mine[par].x >> 16) << 4;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
      my_tx[par] = b;
    }
    // row ids: a ring of kRowRing samples in shared memory, requested kRowRing - 1 samples ahead with 4-byte cp.async
    Walk ahead(0, blockIdx.x, grid);
    auto request_rows = [&]() {
      if (ahead.n < turns) {
        const int64_t s = ahead.sample();
        if (s < a.batch_all) {
          const uint32_t dst = smem_u32(my_rows + (ahead.n % kRowRing) * 64);
          const int32_t* src = a.rows_all + s * fields;
          if (lane < fields) cp_async4(dst + lane * 4, src + lane);
          if (lane + 32 < fields) cp_async4(dst + (lane + 32) * 4, src + lane + 32);
        }
      }
      cp_async_commit();
      ahead.next();
    };
    // the consumers have left their kConsumers partial sums of the sample in this stage: one value, fixed order
    auto finish = [&](int64_t s, int stage, uint32_t wrap) {
      mbar_wait(empty0 + 8 * stage, wrap & 1);
      if (warp == 0 && lane == 0) {
        float v = part[stage * kConsumers];
#pragma unroll
        for (int w = 1; w < kConsumers; ++w) v += part[stage * kConsumers + w];
        a.partial[s] = v;
      }
      __syncwarp();
    };
#pragma unroll 1
    for (int i = 0; i < kRowRing - 1; ++i) request_rows();
    Walk cur(0, blockIdx.x, grid), done(0, blockIdx.x, grid);
    int stage = 0;
    uint32_t wrap = 0;   // how many times the ring has been filled
#pragma unroll 1
    for (; cur.n < turns; cur.next()) {
      const int64_t s = cur.sample();
      if (s >= a.batch_all) break;
      const int par = static_cast<int>(s & 1);
      request_rows();
      cp_async_wait<kRowRing - 1>();   // the row ids of sample cur.n have landed
      __syncwarp();
      if (wrap > 0) {
        finish(done.sample(), stage, wrap - 1);
        done.next();
      }
      const uint32_t bar = full0 + 8 * stage;
      if (lane == 0) mbar_expect_tx(bar, my_tx[par]);
      __syncwarp();
      const int2 e = mine[par];
      if (e.x != 0) {
        const int src = e.x & 0xff, f = (e.x >> 8) & 0xff;
        const uint32_t bytes = static_cast<uint32_t>(e.x >> 16) << 4;
        const unsigned long long g = shard_s[src] +
            static_cast<unsigned long long>(my_rows[(cur.n % kRowRing) * 64 + f]) * a.pitch_bytes;
        bulk_g2s(smem_u32(stage0 + (size_t)stage * a.stage_bytes) + e.y, reinterpret_cast<const void*>(g), bytes, bar);
      }
      if (++stage == a.stages) { stage = 0; ++wrap; }
    }
    cp_async_wait<0>();
    for (; done.n < cur.n; done.next())
      finish(done.sample(), done.n % a.stages, static_cast<uint32_t>(done.n / a.stages));
  } else {
    // ------------------------------------------------------------------ consumers: every warp takes a slice of the
    // sample's items; all of them walk the stages in order, so a parity wait always refers to the fill it means
    const int c = warp - kProducers;
    int stage = 0;
    uint32_t wrap = 0;
#pragma unroll 1
    for (Walk cur(0, blockIdx.x, grid); cur.n < turns; cur.next()) {
      const int64_t s = cur.sample();
      if (s >= a.batch_all) break;
      const int par = static_cast<int>(s & 1);
      // the part of the logit that needed no exchange joins warp 0's slice (requested before the wait)
      float own = 0.f;
      if (c == 0 && lane == 0 && a.first != nullptr && s >= a.own_lo && s < a.own_hi) own = __ldg(a.first + (s - a.own_lo));
      mbar_wait(full0 + 8 * stage, wrap & 1);
      const float4* A = reinterpret_cast<const float4*>(stage0 + (size_t)stage * a.stage_bytes);
      const uint32_t* items = item_s + par * a.max_items;
      const int ni = a.n_items[par];
      float acc0 = 0.f, acc1 = 0.f;
      int q = c * 32 + lane;
      for (; q + kConsumers * 32 < ni; q += 2 * kConsumers * 32) {
        const uint32_t e0 = items[q], e1 = items[q + kConsumers * 32];
        const float4 x0 = A[e0 & 0xffff], y0 = A[e0 >> 16];
        const float4 x1 = A[e1 & 0xffff], y1 = A[e1 >> 16];
        acc0 = fmaf(x0.x, y0.x, acc0); acc0 = fmaf(x0.y, y0.y, acc0);
        acc0 = fmaf(x0.z, y0.z, acc0); acc0 = fmaf(x0.w, y0.w, acc0);
        acc1 = fmaf(x1.x, y1.x, acc1); acc1 = fmaf(x1.y, y1.y, acc1);
        acc1 = fmaf(x1.z, y1.z, acc1); acc1 = fmaf(x1.w, y1.w, acc1);
      }
      if (q < ni) {
        const uint32_t e0 = items[q];
        const float4 x0 = A[e0 & 0xffff], y0 = A[e0 >> 16];
        acc0 = fmaf(x0.x, y0.x, acc0); acc0 = fmaf(x0.y, y0.y, acc0);
        acc0 = fmaf(x0.z, y0.z, acc0); acc0 = fmaf(x0.w, y0.w, acc0);
      }
      const float acc = warp_sum(acc0 + acc1);
      if (lane == 0) part[stage * kConsumers + c] = acc + own;
      __syncwarp();                                     // every lane has read its pieces
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);   // (release: the partial sum is visible to the producer)
      if (++stage == a.stages) { stage = 0; ++wrap; }
    }
  }
}

// idx + offsets -> int32 global row ids (bounds-checked like every lookup of this library), and the part of the logit
// that needs no exchange: bias + sum_f w_feat[r_f] (w_feat is 4 bytes per row and replicated).  One warp per sample.
template <int IdxBits>
__global__ void __launch_bounds__(256) ffm_shard_resolve_kernel(const void* __restrict__ idx,
                                                                const int64_t* __restrict__ offsets, int64_t batch,
                                                                int fields, int64_t rows,
                                                                const float* __restrict__ w_feat,
                                                                const float* __restrict__ bias,
                                                                int32_t* __restrict__ rows_out,
                                                                float* __restrict__ first_out, int32_t* status) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < batch; s += warps) {
    float acc = 0.f;
    for (int f = lane; f < fields; f += 32) {
      const int64_t pos = s * fields + f;
      int64_t r = load_index<IdxBits>(idx, pos) + __ldg(offsets + f);
      if (r < 0 || r >= rows) {
        report_oob(status, pos);
        r = 0;
      } else if (w_feat != nullptr) {
        acc += __ldg(w_feat + r);
      }
      rows_out[pos] = static_cast<int32_t>(r);
    }
    acc = warp_sum(acc);
    if (lane == 0 && first_out != nullptr) first_out[s] = acc + (bias != nullptr ? __ldg(bias) : 0.f);
  }
}

// shard[r][slot][e] = tables[slot][r][e]; a warp writes 128 contiguous bytes
__global__ void __launch_bounds__(256) ffm_shard_pack_kernel(const float* const* __restrict__ tables, int slots,
                                                             int slots_pitch, int64_t rows, int embed,
                                                             float* __restrict__ shard) {
  const int pitch = slots_pitch * embed;
  const int64_t items = rows * pitch;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < items;
       item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = item / pitch;
    const int c = static_cast<int>(item - r * pitch);
    const int slot = c / embed;
    shard[item] = slot < slots ? __ldg(tables[slot] + r * embed + (c - slot * embed)) : 0.f;
  }
}

// ---- the plan (host) ------------------------------------------------------------------------------------------------
inline int tables_on(int fields, int world, int rank) { return rank < fields ? (fields - rank + world - 1) / world : 0; }

// which of the two ranks reduces block {k, m} for the samples of parity `par`
inline int block_owner(int par, int k, int m) {
  const int lo = k < m ? k : m, hi = k < m ? m : k;
  return ((par + lo + hi) & 1) == 0 ? lo : hi;
}

struct Plan {
  std::vector<int2> copies[2];
  std::vector<uint32_t> items[2];
  int tx_bytes[2];
  int stage_bytes;
};

bool build_plan(int fields, int world, int rank, int embed, Plan& pl) {
  const int sk = tables_on(fields, world, rank);
  const int vec = embed * 4;          // bytes of one looked-up vector
  const int pieces = embed / 4;       // 16-byte pieces per vector
  pl.stage_bytes = 128;
  // chunks sit in shared memory at a pitch of 64 mod 128 bytes: the consumers read pieces [a][b] of one operand and
  // [b][a] of the other, 8 pairs per warp instruction -- with that pitch both streams are bank-conflict free
  auto pitch_of = [](int bytes) { return bytes % 128 == 64 ? bytes : bytes + 64 - bytes % 128 + (bytes % 128 > 64 ? 128 : 0); };
  for (int par = 0; par < 2; ++par) {
    auto& cp = pl.copies[par];
    auto& it = pl.items[par];
    cp.clear();
    it.clear();
    int off = 0, tx = 0;              // bytes from the start of the stage; bytes fetched
    auto add_copy = [&](int src, int field, int bytes) {
      cp.push_back(make_int2(src | (field << 8) | ((bytes / 16) << 16), off));
      off += pitch_of(bytes);
      tx += bytes;
    };
    auto add_item = [&](int byte_a, int byte_b) {
      for (int c = 0; c < pieces; ++c)
        it.push_back(static_cast<uint32_t>(byte_a / 16 + c) | (static_cast<uint32_t>(byte_b / 16 + c) << 16));
    };
    // diagonal block: D[a] = chunk_k[r_{i_a}], pairs a < a'
    const int d0 = off;
    for (int a = 0; a < sk; ++a) add_copy(rank, rank + a * world, sk * vec);
    const int dp = pitch_of(sk * vec);
    for (int a = 0; a < sk; ++a)
      for (int b = a + 1; b < sk; ++b) add_item(d0 + a * dp + b * vec, d0 + b * dp + a * vec);
    for (int m = 0; m < world; ++m) {
      if (m == rank || block_owner(par, rank, m) != rank) continue;
      const int sm = tables_on(fields, world, m);
      if (sm == 0 || sk == 0) continue;
      const int r0 = off;             // R[a] = chunk_m[r_{i_a}]  (sm vectors each)
      for (int a = 0; a < sk; ++a) add_copy(m, rank + a * world, sm * vec);
      const int l0 = off;             // L[b] = chunk_k[r_{j_b}]  (sk vectors each)
      for (int b = 0; b < sm; ++b) add_copy(rank, m + b * world, sk * vec);
      const int rp = pitch_of(sm * vec), lp = pitch_of(sk * vec);
      for (int a = 0; a < sk; ++a)
        for (int b = 0; b < sm; ++b) add_item(r0 + a * rp + b * vec, l0 + b * lp + a * vec);
    }
    pl.tx_bytes[par] = tx;
    if (off > pl.stage_bytes) pl.stage_bytes = off;
  }
  pl.stage_bytes = (pl.stage_bytes + 127) / 128 * 128;
  return pl.stage_bytes / 16 <= 0xffff;
}

}  // namespace
}  // namespace trs

using namespace trs;

extern "C" int trs_ffm_shard_plan(int fields, int world, int rank, int embed, int32_t* copy_tab, int copy_capacity,
                                  uint32_t* item_tab, int item_capacity, int* n_copies, int* n_items, int* tx_bytes,
                                  int* stage_bytes) {
  TRS_REQUIRE(fields > 1 && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
              "trs_ffm_shard_plan: need fields > 1, 1 <= world <= %d, 0 <= rank < world", kMaxWorld);
  TRS_UNSUPPORTED(fields > 64, "trs_ffm_shard_plan: at most 64 fields (got %d)", fields);
  TRS_UNSUPPORTED(embed < 4 || embed > 128 || (embed & 3) != 0, "trs_ffm_shard_plan: embed must be a multiple of 4 in [4, 128]");
  TRS_REQUIRE(n_copies && n_items && tx_bytes && stage_bytes, "trs_ffm_shard_plan: null output");
  Plan pl;
  TRS_UNSUPPORTED(!build_plan(fields, world, rank, embed, pl), "trs_ffm_shard_plan: a sample of this shape exceeds 1 MB");
  for (int par = 0; par < 2; ++par) {
    n_copies[par] = static_cast<int>(pl.copies[par].size());
    n_items[par] = static_cast<int>(pl.items[par].size());
    tx_bytes[par] = pl.tx_bytes[par];
  }
  *stage_bytes = pl.stage_bytes;
  if (copy_tab != nullptr) {   // [2][copy_capacity][2]
    for (int par = 0; par < 2; ++par) {
      TRS_REQUIRE(n_copies[par] <= copy_capacity, "trs_ffm_shard_plan: copy_capacity too small");
      for (int c = 0; c < copy_capacity; ++c) {
        const bool on = c < n_copies[par];
        copy_tab[(par * copy_capacity + c) * 2] = on ? pl.copies[par][c].x : 0;
        copy_tab[(par * copy_capacity + c) * 2 + 1] = on ? pl.copies[par][c].y : 0;
      }
    }
  }
  if (item_tab != nullptr) {   // [2][item_capacity]
    for (int par = 0; par < 2; ++par) {
      TRS_REQUIRE(n_items[par] <= item_capacity, "trs_ffm_shard_plan: item_capacity too small");
      for (int q = 0; q < item_capacity; ++q) item_tab[par * item_capacity + q] = q < n_items[par] ? pl.items[par][q] : 0u;
    }
  }
  return TRS_OK;
}

extern "C" int trs_ffm_shard_pack(const float* const* tables, int slots, int slots_pitch, int64_t rows, int embed,
                                  float* shard, void* stream) {
  TRS_REQUIRE(tables && shard, "trs_ffm_shard_pack: null pointer");
  TRS_REQUIRE(slots >= 0 && slots <= slots_pitch && rows >= 0 && embed > 0, "trs_ffm_shard_pack: bad sizes");
  if (rows == 0 || slots_pitch == 0) return TRS_OK;
  ffm_shard_pack_kernel<<<grid_for(rows * slots_pitch * embed, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      tables, slots, slots_pitch, rows, embed, shard);
  return check_launch("ffm_shard_pack_kernel");
}

extern "C" int trs_ffm_shard_resolve(const void* idx, int idx_bits, const int64_t* offsets, int64_t batch, int fields,
                                     int64_t rows, const float* w_feat, const float* bias, int32_t* rows_out,
                                     float* first_out, int32_t* status, void* stream) {
  TRS_REQUIRE(idx && offsets && rows_out && status, "trs_ffm_shard_resolve: null pointer");
  TRS_REQUIRE(idx_bits == 32 || idx_bits == 64, "trs_ffm_shard_resolve: idx_bits must be 32 or 64");
  TRS_REQUIRE(batch >= 0 && fields > 0 && rows > 0, "trs_ffm_shard_resolve: bad sizes");
  TRS_UNSUPPORTED(rows >= (int64_t(1) << 31), "trs_ffm_shard_resolve: row ids must fit 31 bits");
  if (batch == 0) return TRS_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(batch * 32, 256, 8);
  if (idx_bits == 64)
    ffm_shard_resolve_kernel<64><<<grid, 256, 0, s>>>(idx, offsets, batch, fields, rows, w_feat, bias, rows_out,
                                                      first_out, status);
  else
    ffm_shard_resolve_kernel<32><<<grid, 256, 0, s>>>(idx, offsets, batch, fields, rows, w_feat, bias, rows_out,
                                                      first_out, status);
  return check_launch("ffm_shard_resolve_kernel");
}

extern "C" int trs_ffm_shard_blocks(const int32_t* rows_all, int64_t batch_all, int fields, int embed,
                                    const float* const* shards, int world, int rank, const int32_t* copy_tab,
                                    int copy_capacity, const uint32_t* item_tab, int item_capacity, const float* first,
                                    int64_t own_lo, int64_t own_hi, float* partial, void* stream) {
  TRS_REQUIRE(rows_all && shards && copy_tab && item_tab && partial, "trs_ffm_shard_blocks: null pointer");
  TRS_REQUIRE(batch_all >= 0 && fields > 1 && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
              "trs_ffm_shard_blocks: bad sizes");
  TRS_REQUIRE(own_lo >= 0 && own_lo <= own_hi && own_hi <= batch_all, "trs_ffm_shard_blocks: bad sample range");
  TRS_UNSUPPORTED(fields > 64, "trs_ffm_shard_blocks: at most 64 fields (got %d)", fields);
  TRS_UNSUPPORTED(embed < 4 || embed > 128 || (embed & 3) != 0, "trs_ffm_shard_blocks: embed must be a multiple of 4 in [4, 128]");
  Plan pl;   // the sizes are re-derived here: the device tables must come from trs_ffm_shard_plan with the same arguments
  TRS_UNSUPPORTED(!build_plan(fields, world, rank, embed, pl), "trs_ffm_shard_blocks: a sample of this shape exceeds 1 MB");
  BlockArgs a{};
  a.rows_all = rows_all; a.batch_all = batch_all; a.fields = fields;
  for (int r = 0; r < kMaxWorld; ++r) {
    const float* sp = shards[r < world ? r : 0];
    TRS_REQUIRE(sp != nullptr && aligned16(sp), "trs_ffm_shard_blocks: shard %d is null or not 16-byte aligned", r);
    a.shard[r] = reinterpret_cast<const unsigned char*>(sp);
  }
  a.pitch_bytes = static_cast<int64_t>((fields + world - 1) / world) * embed * 4;
  a.copy_tab = reinterpret_cast<const int2*>(copy_tab);
  a.item_tab = item_tab;
  for (int par = 0; par < 2; ++par) {
    a.n_copies[par] = static_cast<int>(pl.copies[par].size());
    a.n_items[par] = static_cast<int>(pl.items[par].size());
    a.tx_bytes[par] = pl.tx_bytes[par];
    TRS_UNSUPPORTED(a.n_copies[par] > 32 * kProducers, "trs_ffm_shard_blocks: more than %d chunks per sample", 32 * kProducers);
    TRS_REQUIRE(a.n_copies[par] <= copy_capacity && a.n_items[par] <= item_capacity,
                "trs_ffm_shard_blocks: table capacities smaller than the plan");
  }
  a.max_copies = copy_capacity; a.max_items = item_capacity;
  a.stage_bytes = pl.stage_bytes;
  const size_t fixed = (size_t)2 * copy_capacity * 8 + (size_t)2 * item_capacity * 4 + 16 + kProducers * kRowRing * 64 * 4 + kMaxWorld * 8 +
                       2 * kMaxStages * 8 + kMaxStages * kConsumers * 4;
  TRS_UNSUPPORTED(fixed + 2 * (size_t)pl.stage_bytes > (size_t)kMaxDynSmem,
                  "trs_ffm_shard_blocks: two samples of %d fields x %d over %d ranks do not fit shared memory", fields,
                  embed, world);
  int stages = static_cast<int>(((size_t)kMaxDynSmem - fixed) / pl.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  a.stages = stages;
  a.first = first; a.own_lo = own_lo; a.own_hi = own_hi; a.partial = partial;
  if (batch_all == 0) return TRS_OK;
  const size_t smem = fixed + (size_t)stages * pl.stage_bytes;
  TRS_SMEM_OPT_IN(ffm_blocks_kernel);
  const int grid = static_cast<int>(batch_all < kNumSMs ? batch_all : kNumSMs);
  ffm_blocks_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("ffm_blocks_kernel");
}
